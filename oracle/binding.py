"""binding.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy-facing wrapper of oracle/liboracle.so (navsim_oracle.c): N reference environments
stepped on the CPU.  Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "navsim_oracle.c")
    deps = [src, os.path.join(HERE, "..", "include", "navsim.h"),
            os.path.join(HERE, "..", "navbot_ppo_b200", "csrc", "navsim_math.h")]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if stale:
        subprocess.run(["make", "-C", HERE, "-B", "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


class _State(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("x", "y", "th", "gx", "gy", "past", "pa0", "pa1", "steps", "draws", "ep_ret", "ep_path", "last_move", "vl", "vr")]


class _Map(ctypes.Structure):
    _fields_ = [("seg", ctypes.c_void_p), ("S", ctypes.c_int32), ("closed_boxes", ctypes.c_int32),
                ("bc", ctypes.c_void_p), ("bs", ctypes.c_void_p),
                ("starts", ctypes.c_void_p), ("n_starts", ctypes.c_int32),
                ("goals", ctypes.c_void_p), ("n_goals", ctypes.c_int32)]


class OracleCfg(ctypes.Structure):
    """include/navsim.h `navsim_cfg`, restated so that the CPU arm of bench.py never imports the product
    package (tests/test_oracle_env.py checks the two definitions field by field)."""
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
        ("reset_rects", ctypes.c_double * 32), ("respawn_rects", ctypes.c_double * 32),
        ("lidar_noise_sigma", ctypes.c_double), ("wheel_accel", ctypes.c_double), ("wheel_separation", ctypes.c_double),
        ("sampler_min_dist", ctypes.c_double), ("sampler_max_dist", ctypes.c_double),
        ("sampler_mode", ctypes.c_int32), ("reserved0", ctypes.c_int32),
    ]


# stage_1 = worlds/train_world1.world:85-252 (turtlebot3_stage_1.launch:8): four walls as
# (centre x, centre y, size x, size y, yaw), yaw as printed in the world file
STAGE_1_BOXES = [(4.0, 0.0, 8.1, 0.1, -1.5708), (0.0, -4.0, 8.10002, 0.1, 3.14159), (-4.0, 0.0, 8.1, 0.1, 1.5708),
                 (0.0, 4.0, 8.1, 0.1, 0.0)]


def stage_1_segments() -> np.ndarray:
    """The four counter-clockwise edges of each stage_1 wall, float64 [16, 4]."""
    import math
    segs = []
    for cx, cy, sx, sy, yaw in STAGE_1_BOXES:
        c, s = math.cos(yaw), math.sin(yaw)
        hx, hy = sx / 2.0, sy / 2.0
        pts = [(cx + c * px - s * py, cy + s * px + c * py) for px, py in ((-hx, -hy), (hx, -hy), (hx, hy), (-hx, hy))]
        segs += [(*pts[k], *pts[(k + 1) % 4]) for k in range(4)]
    return np.asarray(segs, dtype=np.float64).reshape(-1, 4)


def default_cfg(num_agents: int) -> OracleCfg:
    cfg = OracleCfg()
    lib().oracle_default_cfg(ctypes.byref(cfg), int(num_agents))
    return cfg


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        vp, d, i32, u32, u64 = ctypes.c_void_p, ctypes.c_double, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64
        L.oracle_reset.argtypes = [vp, vp, vp, vp, vp]
        L.oracle_reset.restype = None
        L.oracle_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int]
        L.oracle_step.restype = ctypes.c_int
        L.oracle_run_scripted.argtypes = [vp, vp, vp, u64, i32, i32, vp, vp, vp, vp, vp, vp, ctypes.c_int]
        L.oracle_run_scripted.restype = ctypes.c_int
        L.oracle_table_indices.argtypes = [u64, u64, u32, ctypes.c_int, ctypes.c_int, vp, vp]
        L.oracle_table_indices.restype = None
        L.oracle_drive_ramped.argtypes = [vp, vp, vp, vp, vp, d, d, d, d, d, ctypes.c_int]
        L.oracle_drive_ramped.restype = None
        L.oracle_default_cfg.argtypes = [vp, i32]
        L.oracle_default_cfg.restype = ctypes.c_int
        L.oracle_scripted_actions.argtypes = [u64, ctypes.c_int64, i32, i32, vp]
        L.oracle_scripted_actions.restype = None
        L.shim_beam_table.argtypes = [i32, d, d, vp, vp]
        L.shim_beam_table.restype = None
        L.shim_scan.argtypes = [d, d, d, vp, i32, i32, i32, vp, vp, d, d, d, vp]
        L.shim_scan.restype = None
        L.shim_pack_segments.argtypes = [vp, i32, d, vp]
        L.shim_pack_segments.restype = None
        for name in ("oracle_nv_sin", "oracle_nv_cos", "oracle_nv_atan"):
            getattr(L, name).argtypes = [d]
            getattr(L, name).restype = d
        L.oracle_nv_pyround.argtypes = [d, ctypes.c_int]
        L.oracle_nv_pyround.restype = d
        L.ref_pyround.argtypes = [d, ctypes.c_int]
        L.ref_pyround.restype = d
        L.ref_odometry.argtypes = [d, d, d, d, d, d, vp, vp, vp]
        L.ref_odometry.restype = None
        L.oracle_philox.argtypes = [u32, u32, u32, u32, u32, u32, vp]
        L.oracle_philox.restype = None
        _lib = L
    return _lib


_FIELDS = [("x", np.float64), ("y", np.float64), ("th", np.float64), ("gx", np.float64), ("gy", np.float64),
           ("past", np.float64), ("pa0", np.float32), ("pa1", np.float32), ("steps", np.int32),
           ("draws", np.uint32), ("ep_ret", np.float32), ("ep_path", np.float32), ("last_move", np.float32),
           ("vl", np.float64), ("vr", np.float64)]


class OracleSim:
    """N independent reference environments on the CPU (state as numpy SoA arrays)."""

    def __init__(self, cfg, segments, nthreads: int = 1, closed_boxes: bool = True, sampler_tables=None):
        self.cfg = cfg  # a navbot_ppo_b200._capi.NavsimCfg (the public C struct)
        self.n = int(cfg.num_agents)
        self.nthreads = nthreads
        self.seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 4)
        nb = int(cfg.num_beams)
        self.bc, self.bs = np.zeros(nb, np.float32), np.zeros(nb, np.float32)
        lib().shim_beam_table(nb, cfg.fov_min, cfg.fov_max, self.bc.ctypes.data, self.bs.ctypes.data)
        self.segf = np.zeros((len(self.seg), 8), np.float32)
        lib().shim_pack_segments(self.seg.ctypes.data, len(self.seg), cfg.lidar_max, self.segf.ctypes.data)
        self.closed_boxes = 1 if closed_boxes else 0
        if sampler_tables is not None:   # (starts [n,3], goals [n,2]) of spawn_goal_sampler.py:5-35
            self.starts = np.ascontiguousarray(sampler_tables[0], dtype=np.float64).reshape(-1, 3)
            self.goals = np.ascontiguousarray(sampler_tables[1], dtype=np.float64).reshape(-1, 2)
            self.map = _Map(self.segf.ctypes.data, len(self.seg), self.closed_boxes, self.bc.ctypes.data, self.bs.ctypes.data,
                            self.starts.ctypes.data, len(self.starts), self.goals.ctypes.data, len(self.goals))
        else:
            self.map = _Map(self.segf.ctypes.data, len(self.seg), self.closed_boxes, self.bc.ctypes.data,
                            self.bs.ctypes.data, None, 0, None, 0)
        self.arr = {name: np.zeros(self.n, dtype=dt) for name, dt in _FIELDS}
        self.state = _State(*[self.arr[name].ctypes.data for name, _ in _FIELDS])
        self.stats = None

    def reset(self, mask=None):
        obs = np.zeros((self.n, 16))
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().oracle_reset(ctypes.byref(self.cfg), ctypes.byref(self.map), ctypes.byref(self.state),
                           None if m is None else m.ctypes.data, obs.ctypes.data)
        return obs

    def step(self, act, stats=None):
        act = np.ascontiguousarray(act, dtype=np.float32).reshape(self.n, 2)
        obs = np.zeros((self.n, 16))
        rew = np.zeros(self.n)
        done = np.zeros(self.n, np.uint8)
        arrive = np.zeros(self.n, np.uint8)
        trunc = np.zeros(self.n, np.uint8)
        lib().oracle_step(ctypes.byref(self.cfg), ctypes.byref(self.map), ctypes.byref(self.state), act.ctypes.data,
                          obs.ctypes.data, rew.ctypes.data, done.ctypes.data, arrive.ctypes.data, trunc.ctypes.data,
                          None if stats is None else ctypes.byref(stats), self.nthreads)
        return obs, rew, done, arrive, trunc

    def run_scripted(self, nsteps: int, action_seed: int = 0, step0: int = 0):
        """`nsteps` Env.step calls per agent with the benchmark's scripted actions, entirely in C
        (liboracle.so: oracle_run_scripted); returns the last step's (obs, rew, done, arrive)."""
        if not hasattr(self, "_run_bufs"):
            self._run_bufs = (np.zeros((self.n, 2), np.float32), np.zeros((self.n, 16)), np.zeros(self.n),
                              np.zeros(self.n, np.uint8), np.zeros(self.n, np.uint8))
        act, obs, rew, done, arrive = self._run_bufs
        lib().oracle_run_scripted(ctypes.byref(self.cfg), ctypes.byref(self.map), ctypes.byref(self.state), action_seed,
                                  step0, nsteps, act.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data,
                                  arrive.ctypes.data, None, self.nthreads)
        return obs, rew, done, arrive

    def scan(self):
        nb = int(self.cfg.num_beams)
        out = np.zeros((self.n, nb))
        c = self.cfg
        for i in range(self.n):
            lib().shim_scan(self.arr["x"][i], self.arr["y"][i], self.arr["th"][i], self.segf.ctypes.data, len(self.seg),
                            self.closed_boxes, nb, self.bc.ctypes.data, self.bs.ctypes.data, c.lidar_offset_x, c.lidar_min, c.lidar_max,
                            out[i].ctypes.data)
        return out


def scripted_actions(action_seed: int, agent_id_offset: int, step: int, n: int) -> np.ndarray:
    act = np.zeros((n, 2), np.float32)
    lib().oracle_scripted_actions(action_seed, agent_id_offset, step, n, act.ctypes.data)
    return act
