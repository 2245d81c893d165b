"""Host-side mirror of the reference's environment interface.

  Env     — same constructor / reset() / step(action, past_action) surface as
            project_ppo/src/environment_new.py:26-382 (one robot, numpy in/out), so the
            reference's main.py / ppo.py drop in by re-pointing one import (main.py:433).
  VecEnv  — N robots per GPU; torch tensors in/out on the device, auto-reset with the
            rollout's episode protocol (ppo.py:549-593) folded into the step kernel.

Both are thin: every number comes out of libnavbot_b200.so (csrc/navsim_kernels.cu).
"""
from __future__ import annotations

import ctypes
from types import SimpleNamespace

import numpy as np
import torch

from . import _capi, maps


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class VecEnv:
    """N agents of the reference Env stepped by one kernel launch per step."""

    obs_dim = _capi.OBS_DIM
    act_dim = _capi.ACT_DIM

    def __init__(self, num_envs: int, map: str | np.ndarray = "stage_1", device: int | str | torch.device = 0,
                 seed: int = 0, max_episode_steps: int = 500, auto_reset: bool = True, is_training: bool = True,
                 num_beams: int = 10, agent_id_offset: int = 0, cfg: _capi.NavsimCfg | None = None,
                 closed_boxes: bool = True, lanes_per_agent: int = 0, lidar: str | None = None,
                 lidar_noise_sigma: float = 0.0, wheel_accel: float = 0.0, use_external_sampler: bool | str = False,
                 sampler_min_dist: float = 1.5, sampler_max_dist: float = 6.0):
        """Options beyond the reference Env (all off by default):
        lidar="waffle"            the 360-beam, full-circle scan of turtlebot3_waffle.gazebo.xacro:118-125
        lidar_noise_sigma         Gaussian range noise of the ray-sensor plugin (burger.gazebo.xacro:122-126: 0.01)
        wheel_accel               the diff-drive plugin's wheel acceleration limit (burger.gazebo.xacro:67: 1)
        use_external_sampler      start pose + goal from the GoalSpawnSampler tables (spawn_goal_sampler.py:37-72,
                                  arguments.py:41): True = the table set of the map, or a world_type name"""
        self._h = ctypes.c_void_p()
        L = _capi.lib()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise ValueError("VecEnv runs on a CUDA device; there is no CPU path")
        self.device = dev
        if cfg is None:
            cfg = _capi.default_cfg(num_envs)
            cfg.seed = seed
            cfg.max_episode_steps = max_episode_steps
            cfg.auto_reset = 1 if auto_reset else 0
            cfg.num_beams = num_beams
            if lidar is not None:
                if lidar != "waffle":
                    raise ValueError("lidar presets: 'waffle' (default: the burger's 10-beam front scan)")
                cfg.num_beams, cfg.fov_min, cfg.fov_max = 360, 0.0, 6.28319   # waffle.gazebo.xacro:118-125
            cfg.lidar_noise_sigma = float(lidar_noise_sigma)
            cfg.wheel_accel = float(wheel_accel)
            cfg.sampler_mode = 1 if use_external_sampler else 0
            cfg.sampler_min_dist, cfg.sampler_max_dist = float(sampler_min_dist), float(sampler_max_dist)
            cfg.agent_id_offset = agent_id_offset
            cfg.lanes_per_agent = lanes_per_agent   # 0: chosen from N by the library
            # environment_new.py:44-47
            cfg.arrive_threshold = 0.2 if is_training else 0.4
            if isinstance(map, str) and map in maps.SPAWN:
                cfg.start_x, cfg.start_y, cfg.start_theta = maps.SPAWN[map]
                cfg.goal_lo, cfg.goal_hi, rects = maps.GOAL_RANGE[map]
                if not rects:
                    cfg.n_reset_rects = cfg.n_respawn_rects = 0
        cfg.device = dev.index if dev.index is not None else torch.cuda.current_device()
        self.cfg = cfg
        self.num_envs = int(cfg.num_agents)
        _capi.check(L.navsim_create(ctypes.byref(self._h), ctypes.byref(cfg)))
        seg = maps.get_map(map) if isinstance(map, str) else np.asarray(map)
        self.segments = np.ascontiguousarray(seg, dtype=np.float64).reshape(-1, 4)
        # maps.get_map()/boxes_to_segments() emit counter-clockwise box edges; pass closed_boxes=False
        # for hand-made free-standing walls
        self.closed_boxes = bool(closed_boxes)
        _capi.check(L.navsim_set_map(self._h, self.segments.ctypes.data, len(self.segments),
                                     _capi.MAP_CLOSED_BOXES if closed_boxes else 0))
        self.sampler_tables = None
        if use_external_sampler:      # (a caller-built cfg with sampler_mode = 1 sets its tables itself: set_sampler)
            world = use_external_sampler if isinstance(use_external_sampler, str) else \
                maps.SAMPLER_FOR_MAP.get(map if isinstance(map, str) else "", None)
            if world is None:
                raise ValueError("use_external_sampler=True needs a named map; pass the world_type ('small_house', 'stage1')")
            self.set_sampler(*maps.sampler_tables(world))
        n = self.num_envs
        self.obs = torch.zeros((n, self.obs_dim), dtype=torch.float32, device=dev)
        self.rew = torch.zeros(n, dtype=torch.float32, device=dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.arrive = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.trunc = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._host_ptr_key, self._host_ptrs, self._host_keepalive = None, None, None
        self._step_host_fn = L.navsim_step_host
        self._step_async_fn, self._async_ptrs, self._async_keepalive = L.navsim_step_host_async, {}, None
        self._step_pipelined_fn = L.navsim_step_host_pipelined

    def set_sampler(self, starts, goals):
        """GoalSpawnSampler tables: starts [n, 3] (x, y, yaw), goals [n, 2] (navsim_set_sampler)."""
        st = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 3)
        go = np.ascontiguousarray(goals, dtype=np.float64).reshape(-1, 2)
        _capi.check(_capi.lib().navsim_set_sampler(self._h, st.ctypes.data, len(st), go.ctypes.data, len(go)))
        self.sampler_tables = (st, go)

    # -- lifecycle ----------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _capi.lib().navsim_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device API ---------------------------------------------------------------------
    def reset(self, mask: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        obs = self.obs if out is None else out
        mptr = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mptr = mask.data_ptr()
        _capi.check(_capi.lib().navsim_reset(self._h, mptr, obs.data_ptr(), _stream_ptr(self.device)))
        return obs

    def step(self, actions: torch.Tensor, out_obs: torch.Tensor | None = None, out_rew: torch.Tensor | None = None,
             out_done: torch.Tensor | None = None, out_arrive: torch.Tensor | None = None,
             out_trunc: torch.Tensor | None = None):
        """actions[N,2] float32 on the device -> (obs[N,16], rew[N], done[N], arrive[N]) (+ self.trunc)."""
        if actions.dtype != torch.float32 or not actions.is_contiguous() or actions.device != self.device:
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if actions.numel() != self.num_envs * 2:
            raise ValueError(f"actions must have shape [{self.num_envs}, 2]")
        obs = self.obs if out_obs is None else out_obs
        rew = self.rew if out_rew is None else out_rew
        done = self.done if out_done is None else out_done
        arrive = self.arrive if out_arrive is None else out_arrive
        trunc = self.trunc if out_trunc is None else out_trunc
        _capi.check(_capi.lib().navsim_step(self._h, actions.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                            arrive.data_ptr(), trunc.data_ptr(), _stream_ptr(self.device)))
        return obs, rew, done, arrive

    def step_scripted(self, num_steps: int, action_seed: int = 0):
        _capi.check(_capi.lib().navsim_step_scripted(self._h, num_steps, action_seed, self.obs.data_ptr(),
                                                     self.rew.data_ptr(), self.done.data_ptr(), self.arrive.data_ptr(),
                                                     _stream_ptr(self.device)))
        return self.obs, self.rew, self.done, self.arrive

    def rollout_scripted(self, num_steps: int, action_seed: int = 0, out: dict | None = None):
        """`num_steps` scripted steps in ONE launch, outputs in the rollout layout [H, N, .]
        (navsim_rollout_scripted).  Returns dict(obs, rew, done, arrive, trunc)."""
        n, d = self.num_envs, self.device
        if out is None:
            out = dict(obs=torch.empty((num_steps, n, self.obs_dim), dtype=torch.float32, device=d),
                       rew=torch.empty((num_steps, n), dtype=torch.float32, device=d),
                       done=torch.empty((num_steps, n), dtype=torch.uint8, device=d),
                       arrive=torch.empty((num_steps, n), dtype=torch.uint8, device=d),
                       trunc=torch.empty((num_steps, n), dtype=torch.uint8, device=d))
        assert out["obs"].shape[0] >= num_steps
        _capi.check(_capi.lib().navsim_rollout_scripted(self._h, num_steps, action_seed, out["obs"].data_ptr(),
                                                        out["rew"].data_ptr(), out["done"].data_ptr(),
                                                        out["arrive"].data_ptr(), out["trunc"].data_ptr(),
                                                        _stream_ptr(self.device)))
        return out

    def scan(self) -> torch.Tensor:
        out = torch.empty((self.num_envs, int(self.cfg.num_beams)), dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib().navsim_scan(self._h, out.data_ptr(), _stream_ptr(self.device)))
        return out

    # -- host API (pinned staging inside the library) -------------------------------------
    def reset_host(self, mask: np.ndarray | None = None) -> np.ndarray:
        obs = np.zeros((self.num_envs, self.obs_dim), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        _capi.check(_capi.lib().navsim_reset_host(self._h, None if m is None else m.ctypes.data, obs.ctypes.data))
        return obs

    def alloc_host_buffers(self):
        """Page-locked host arrays for step_host(): dict(act, obs, rew, done, arrive, trunc).  With
        these the library DMAs straight from / into the caller's memory (no staging copy)."""
        n = self.num_envs
        # the outputs are views of ONE page-locked block laid out obs | rew | done | arrive | trunc: the asynchronous
        # step then returns a whole step's results with a single device-to-host copy
        block = torch.empty(n * (4 * self.obs_dim + 4 + 3), dtype=torch.uint8).pin_memory().numpy()
        o0, r0, f0 = 0, 4 * self.obs_dim * n, (4 * self.obs_dim + 4) * n
        act = torch.empty((n, 2), dtype=torch.float32).pin_memory().numpy()
        return dict(act=act, obs=block[o0:r0].view(np.float32).reshape(n, self.obs_dim), rew=block[r0:f0].view(np.float32),
                    done=block[f0:f0 + n], arrive=block[f0 + n:f0 + 2 * n], trunc=block[f0 + 2 * n:f0 + 3 * n], _block=block)

    def step_host(self, actions: np.ndarray, out: dict | None = None):
        """Env.step with HOST buffers: actions[N,2] float32 in, (obs, rew, done, arrive, trunc) numpy
        arrays out (written into `out` when given, e.g. the arrays of alloc_host_buffers())."""
        n = self.num_envs
        a = actions if (isinstance(actions, np.ndarray) and actions.dtype == np.float32 and actions.flags.c_contiguous
                        and actions.size == 2 * n) else np.ascontiguousarray(actions, np.float32).reshape(n, 2)
        if out is None:
            out = dict(obs=np.empty((n, self.obs_dim), np.float32), rew=np.empty(n, np.float32), done=np.empty(n, np.uint8),
                       arrive=np.empty(n, np.uint8), trunc=np.empty(n, np.uint8))
        obs, rew, done, arrive, trunc = out["obs"], out["rew"], out["done"], out["arrive"], out["trunc"]
        # raw addresses are cached per buffer set: ndarray.ctypes builds a helper object on every access
        key = (id(a), id(obs), id(rew), id(done), id(arrive), id(trunc))
        if self._host_ptr_key != key:
            self._host_ptr_key = key
            self._host_ptrs = tuple(int(x.__array_interface__["data"][0]) for x in (a, obs, rew, done, arrive, trunc))
            self._host_keepalive = (a, obs, rew, done, arrive, trunc)
        rc = self._step_host_fn(self._h, *self._host_ptrs)
        if rc:
            _capi.check(rc)
        return obs, rew, done, arrive, trunc

    def step_host_async(self, actions: np.ndarray, out: dict) -> int:
        """Enqueue Env.step with page-locked HOST buffers (alloc_host_buffers()) and return a ticket at once;
        wait(ticket) blocks until `out` holds that step's results.  Up to four steps may be in flight: cycle through
        as many buffer sets and prepare later steps while earlier observations cross PCIe (navsim_step_host_async)."""
        key = (id(actions), id(out))
        ptrs = self._async_ptrs.get(key)
        if ptrs is None:
            ptrs = tuple(int(x.__array_interface__["data"][0]) for x in (actions, out["obs"], out["rew"], out["done"],
                                                                          out["arrive"], out["trunc"]))
            if len(self._async_ptrs) > 8:
                self._async_ptrs.clear()
            self._async_ptrs[key] = ptrs
            self._async_keepalive = (actions, out)
        t = self._step_async_fn(self._h, *ptrs)
        if t < 0:
            _capi.check(int(t))
        return int(t)

    def step_host_pipelined(self, actions: np.ndarray, out: dict) -> int:
        """step_host_async + the wait of a steady pipeline in ONE library call: enqueues the step, then blocks until the
        step issued three calls earlier has delivered its results — with four buffer sets used in turn, those are in
        the set the next call will overwrite (navsim_step_host_pipelined).  Returns the new step's ticket."""
        key = (id(actions), id(out))
        ptrs = self._async_ptrs.get(key)
        if ptrs is None:
            ptrs = tuple(int(x.__array_interface__["data"][0]) for x in (actions, out["obs"], out["rew"], out["done"],
                                                                          out["arrive"], out["trunc"]))
            if len(self._async_ptrs) > 8:
                self._async_ptrs.clear()
            self._async_ptrs[key] = ptrs
            self._async_keepalive = (actions, out)
        t = self._step_pipelined_fn(self._h, *ptrs)
        if t < 0:
            _capi.check(int(t))
        return int(t)

    def wait(self, ticket: int = 0) -> None:
        """Block until the asynchronous step `ticket` (0: every step issued so far) has delivered its outputs."""
        _capi.check(_capi.lib().navsim_wait(self._h, int(ticket)))

    # -- state inspection / injection ------------------------------------------------------
    def get_state(self, field: int) -> np.ndarray:
        out = np.empty(self.num_envs, dtype=_capi.FIELD_DTYPES[field])
        _capi.check(_capi.lib().navsim_get_state(self._h, field, out.ctypes.data))
        return out

    def set_state(self, field: int, values) -> None:
        v = np.ascontiguousarray(values, dtype=_capi.FIELD_DTYPES[field]).reshape(self.num_envs)
        _capi.check(_capi.lib().navsim_set_state(self._h, field, v.ctypes.data))

    def stats(self, clear: bool = False) -> _capi.NavsimStats:
        s = _capi.NavsimStats()
        _capi.check(_capi.lib().navsim_get_stats(self._h, ctypes.byref(s), 1 if clear else 0))
        return s

    def clear_stats(self) -> None:
        """Zero the episode statistics in stream order (no read-back, no synchronisation)."""
        _capi.check(_capi.lib().navsim_clear_stats(self._h, _stream_ptr(self.device)))

    @property
    def lanes_per_agent(self) -> int:
        """GPU lanes cooperating on one agent's step (chosen by the library from N unless requested)."""
        return int(_capi.lib().navsim_lanes_per_agent(self._h))

    @property
    def launch_count(self) -> int:
        return int(_capi.lib().navsim_launch_count(self._h))


class Env:
    """Drop-in for environment_new.Env: one robot, numpy in / numpy out.

    reset() -> ndarray(16,); step(action, past_action) -> (ndarray(16,), float, bool, bool)
    (environment_new.py:272-310, 312-382).  `past_action` is what the caller passes — the
    action *before* the one being executed (ppo.py:541-543) — and is copied into the
    observation exactly as the reference does; the simulator's own copy is overwritten
    with it first so both bookkeeping schemes agree.
    """

    def __init__(self, is_training, use_vision=False, vision_dim=64, map="stage_1", device=0, seed=0):
        if use_vision:
            raise NotImplementedError("the camera path is outside the LiDAR hot path (SURVEY.md section 2, row 9)")
        self.use_vision = False
        self.vision_dim = vision_dim
        self._vec = VecEnv(1, map=map, device=device, seed=seed, auto_reset=False, is_training=is_training,
                           max_episode_steps=1 << 30)
        self.threshold_arrive = 0.2 if is_training else 0.4
        # one C call per reset / step (navsim_*_host_ex): action, caller-owned previous action in; observation,
        # reward, flags and the pose block out
        self._L = _capi.lib()
        self._step_ex = self._L.navsim_step_host_ex
        self._act, self._past = np.zeros((1, 2), np.float32), np.zeros((1, 2), np.float32)
        self._obs, self._rew = np.zeros((1, 16), np.float32), np.zeros(1, np.float32)
        self._done, self._arrive, self._trunc = np.zeros(1, np.uint8), np.zeros(1, np.uint8), np.zeros(1, np.uint8)
        self._pose = np.zeros((1, 6), np.float64)
        self._ptrs = tuple(int(x.__array_interface__["data"][0]) for x in
                           (self._act, self._past, self._obs, self._rew, self._done, self._arrive, self._trunc, self._pose))
        self.position = SimpleNamespace(x=0.0, y=0.0, z=0.0)
        self.goal_position = SimpleNamespace(position=SimpleNamespace(x=0.0, y=0.0, z=0.01))
        self.past_distance = 0.0

    def _publish(self):
        """env.position / env.goal_position / env.past_distance (read by ppo.py:535, main.py:202) from the
        pose block the last call returned."""
        x, y, _th, gx, gy, past = (float(v) for v in self._pose[0])
        self.position.x, self.position.y = x, y
        self.goal_position.position.x, self.goal_position.position.y = gx, gy
        self.past_distance = past

    def getLatestImage(self):
        return None

    def reset(self):
        _capi.check(self._L.navsim_reset_host_ex(self._vec._h, None, self._obs.ctypes.data, self._pose.ctypes.data))
        self._publish()
        return self._obs[0].astype(np.float64)

    def step(self, action, past_action):
        self._act[0] = np.asarray(action, dtype=np.float32).reshape(2)
        self._past[0] = np.asarray(past_action, dtype=np.float32).reshape(2)
        _capi.check(self._step_ex(self._vec._h, *self._ptrs))
        self._publish()
        return self._obs[0].astype(np.float64), float(self._rew[0]), bool(self._done[0]), bool(self._arrive[0])
