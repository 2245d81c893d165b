"""ctypes binding of include/navsim.h and include/navppo.h.

The shared library is the product: if it is missing or a symbol is absent this module
raises — there is no Python/CPU fallback for the hot path.
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libnavbot_b200.so")

MAX_RECTS = 8
MAP_CLOSED_BOXES = 1
OBS_DIM = 16
ACT_DIM = 2

# enum navsim_field
F_X, F_Y, F_THETA, F_GOAL_X, F_GOAL_Y, F_PAST_DIST, F_PREV_A0, F_PREV_A1, F_STEPS, F_DRAWS, F_EP_RETURN, \
    F_EP_PATH, F_LAST_MOVE, F_WHEEL_L, F_WHEEL_R = range(15)

FIELD_DTYPES = {
    F_X: "float64", F_Y: "float64", F_THETA: "float64", F_GOAL_X: "float64", F_GOAL_Y: "float64",
    F_PAST_DIST: "float64", F_PREV_A0: "float32", F_PREV_A1: "float32", F_STEPS: "int32", F_DRAWS: "uint32",
    F_EP_RETURN: "float32", F_EP_PATH: "float32", F_LAST_MOVE: "float32", F_WHEEL_L: "float64", F_WHEEL_R: "float64",
}


class NavsimCfg(ctypes.Structure):
    _fields_ = [
        ("num_agents", ctypes.c_int32), ("num_beams", ctypes.c_int32), ("max_episode_steps", ctypes.c_int32),
        ("auto_reset", ctypes.c_int32), ("device", ctypes.c_int32), ("n_reset_rects", ctypes.c_int32),
        ("n_respawn_rects", ctypes.c_int32), ("lanes_per_agent", ctypes.c_int32),
        ("seed", ctypes.c_uint64), ("agent_id_offset", ctypes.c_int64),
        ("dt", ctypes.c_double), ("lidar_offset_x", ctypes.c_double),
        ("lidar_min", ctypes.c_double), ("lidar_max", ctypes.c_double),
        ("fov_min", ctypes.c_double), ("fov_max", ctypes.c_double),
        ("collision_range", ctypes.c_double), ("arrive_threshold", ctypes.c_double),
        ("reward_scale", ctypes.c_double), ("reward_collide", ctypes.c_double), ("reward_arrive", ctypes.c_double),
        ("diag_norm", ctypes.c_double), ("goal_lo", ctypes.c_double), ("goal_hi", ctypes.c_double),
        ("start_x", ctypes.c_double), ("start_y", ctypes.c_double), ("start_theta", ctypes.c_double),
        ("reset_rects", ctypes.c_double * (MAX_RECTS * 4)), ("respawn_rects", ctypes.c_double * (MAX_RECTS * 4)),
        ("lidar_noise_sigma", ctypes.c_double), ("wheel_accel", ctypes.c_double), ("wheel_separation", ctypes.c_double),
        ("sampler_min_dist", ctypes.c_double), ("sampler_max_dist", ctypes.c_double),
        ("sampler_mode", ctypes.c_int32), ("reserved0", ctypes.c_int32),
    ]


class NavsimStepOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("obs", "rew", "done", "arrive", "trunc", "ep_return", "ep_path", "ep_len")]


class NavsimStats(ctypes.Structure):
    _fields_ = [
        ("episodes", ctypes.c_uint64), ("successes", ctypes.c_uint64), ("collisions", ctypes.c_uint64),
        ("timeouts", ctypes.c_uint64), ("steps", ctypes.c_uint64),
        ("return_sum", ctypes.c_double), ("length_sum", ctypes.c_double), ("path_sum", ctypes.c_double),
    ]


class NavError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


_vp, _i32, _i64, _u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64


# name -> (restype, argtypes); every symbol include/navsim.h declares
NAVSIM_SYMBOLS = {
    "nav_last_error": (ctypes.c_char_p, []),
    "navsim_abi_version": (ctypes.c_int, []),
    "navsim_default_cfg": (ctypes.c_int, [ctypes.POINTER(NavsimCfg), _i32]),
    "navsim_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.POINTER(NavsimCfg)]),
    "navsim_destroy": (ctypes.c_int, [_vp]),
    "navsim_set_map": (ctypes.c_int, [_vp, _vp, _i32, _i32]),
    "navsim_set_sampler": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i32]),
    "navsim_reset": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "navsim_step": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "navsim_step_ex": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "navsim_reset_host": (ctypes.c_int, [_vp, _vp, _vp]),
    "navsim_step_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "navsim_step_host_async": (_i64, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "navsim_wait": (ctypes.c_int, [_vp, _i64]),
    "navsim_step_host_pipelined": (_i64, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "navsim_step_host_ex": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "navsim_reset_host_ex": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "navsim_step_scripted": (ctypes.c_int, [_vp, _i32, _u64, _vp, _vp, _vp, _vp, _vp]),
    "navsim_rollout_scripted": (ctypes.c_int, [_vp, _i32, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "navsim_scan": (ctypes.c_int, [_vp, _vp, _vp]),
    "navsim_get_state": (ctypes.c_int, [_vp, _i32, _vp]),
    "navsim_set_state": (ctypes.c_int, [_vp, _i32, _vp]),
    "navsim_get_stats": (ctypes.c_int, [_vp, ctypes.POINTER(NavsimStats), _i32]),
    "navsim_clear_stats": (ctypes.c_int, [_vp, _vp]),
    "navsim_num_agents": (ctypes.c_int, [_vp]),
    "navsim_launch_count": (_i64, [_vp]),
    "navsim_lanes_per_agent": (ctypes.c_int, [_vp]),
}

class NavppoCfg(ctypes.Structure):
    _fields_ = [
        ("device", ctypes.c_int32), ("max_samples", ctypes.c_int32), ("precision", ctypes.c_int32),
        ("reserved0", ctypes.c_int32),
        ("lr", ctypes.c_double), ("beta1", ctypes.c_double), ("beta2", ctypes.c_double), ("adam_eps", ctypes.c_double),
        ("clip", ctypes.c_double),
    ]


# include/navppo.h constants
PPO_ACTOR_PARAMS, PPO_CRITIC_PARAMS, PPO_CRITIC_OFFSET, PPO_FLAT, PPO_NUM_METRICS = 50290, 50257, 50304, 100608, 8
M_ACTOR_LOSS, M_CRITIC_LOSS, M_APPROX_KL, M_CLIP_FRAC, M_ACTOR_GRAD_SQ, M_CRITIC_GRAD_SQ = range(6)
PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2

_f64 = ctypes.c_double
_u32 = ctypes.c_uint32

# name -> (restype, argtypes); every symbol include/navppo.h declares
NAVPPO_SYMBOLS = {
    "navppo_default_cfg": (ctypes.c_int, [ctypes.POINTER(NavppoCfg)]),
    "navppo_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.POINTER(NavppoCfg)]),
    "navppo_destroy": (ctypes.c_int, [_vp]),
    "navppo_launch_count": (_i64, [_vp]),
    "navppo_rtg_scan": (ctypes.c_int, [_vp, _vp, _vp, _vp, _f64, _f64, _vp, _i32, _i32, _vp]),
    "navppo_forward": (ctypes.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "navppo_act": (ctypes.c_int, [_vp, _vp, _vp, _i32, _f64, _u64, _i64, _u32, _vp, _vp, _vp, _vp, _vp]),
    "navppo_evaluate": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _f64, _vp, _vp, _vp]),
    "navppo_rollout": (ctypes.c_int, [_vp, _vp, _vp, _i32, _f64, _u64, _i64, _u32] + [_vp] * 11),
    "navppo_rollout_ex": (ctypes.c_int, [_vp, _vp, _vp, _i32, _f64, _u64, _i64, _u32] + [_vp] * 13),
    "navppo_adv_stats": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp]),
    "navppo_adv_normalize": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "navppo_grad": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _f64, _vp, _vp, _vp]),
    "navppo_adam": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "navppo_tc_selftest": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "navppo_tc_selftest_bf16": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "navppo_tc_profile": (ctypes.c_int, [_vp]),
    "navppo_peer_setup": (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "navppo_adam_peer": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _u32, _i32, _vp, _vp]),
    "navppo_update": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _f64, _i32, _vp, _vp, _vp, _vp]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libnavbot_b200.so once; raise loudly if it (or any declared symbol) is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m navbot_ppo_b200.build` "
                "(there is no CPU fallback for the simulator or the PPO kernels)")
        L = ctypes.CDLL(LIB_PATH)
        for table in (NAVSIM_SYMBOLS, NAVPPO_SYMBOLS):
            for name, (res, args) in table.items():
                fn = getattr(L, name)  # AttributeError if the symbol is not exported
                fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().nav_last_error()
        raise NavError(rc, msg.decode() if msg else "unknown error")


def default_ppo_cfg() -> NavppoCfg:
    cfg = NavppoCfg()
    check(lib().navppo_default_cfg(ctypes.byref(cfg)))
    return cfg


def default_cfg(num_agents: int) -> NavsimCfg:
    cfg = NavsimCfg()
    check(lib().navsim_default_cfg(ctypes.byref(cfg), num_agents))
    return cfg
