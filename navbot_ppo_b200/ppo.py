"""Host-side mirror of the reference's trainer interface (project_ppo/src/ppo.py).

    PPO(policy_class, value_func, env, state_dim, action_dim, **hyperparameters)
        learn(total_timesteps, past_action)        ppo.py:218
        rollout(past_action, t_so_far)             ppo.py:463   -> the reference's 7-tuple
        compute_rtgs(batch_rews)                   ppo.py:643
        get_action(obs, t_so_far, one_round)       ppo.py:673
        evaluate(batch_obs, batch_acts)            ppo.py:708
        update(batch)                              the inline update of ppo.py:275-397, factored out

`env` is either navbot_ppo_b200.Env (one robot, the reference's numpy protocol: every line of
the reference's rollout loop has its counterpart here) or navbot_ppo_b200.VecEnv (N robots on
the GPU: policy forward + sampling, environment step and the rollout buffers never leave the
device).  All arithmetic runs in libnavbot_b200.so (include/navppo.h); this file only owns
buffers, bookkeeping, logging and checkpoints.

Multi-GPU: one process per GPU, agents sharded by global id; per epoch ONE all-reduce (sum) of
the flat gradient over NCCL, plus one 3-double all-reduce for the advantage statistics per
iteration.  See dist.py.
"""
from __future__ import annotations

import csv
import ctypes
import math
import os
import time

import numpy as np
import torch

from . import _capi, dist as navdist, layout
from .env import VecEnv
from .nets import NetActor, NetCritic, _Handles, _stream


def _makepath(p):
    os.makedirs(p, exist_ok=True)
    return p


class _FlatAdam:
    """What `agent.actor_optim` / `agent.critic_optim` expose: the moment vectors of one
    network inside the trainer's flat Adam state (ppo.py:116-117; stepped by navppo_adam)."""

    def __init__(self, owner, kind):
        self._owner, self.kind = owner, kind
        off = 0 if kind == "actor" else _capi.PPO_CRITIC_OFFSET
        n = layout.ACTOR_PARAMS if kind == "actor" else layout.CRITIC_PARAMS
        self.exp_avg = owner._exp_avg[off:off + n]
        self.exp_avg_sq = owner._exp_avg_sq[off:off + n]
        self.defaults = dict(lr=owner.lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False)

    def state_dict(self):
        return {"step": self._owner._adam_step, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "defaults": dict(self.defaults)}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self._owner._adam_step = int(sd["step"])

    def zero_grad(self):
        pass


class PPO:
    def __init__(self, policy_class, value_func, env, state_dim, action_dim, **hyperparameters):
        self._init_hyperparameters(hyperparameters)
        self.env = env
        self.obs_dim, self.act_dim = state_dim, action_dim
        if state_dim != layout.OBS_DIM or action_dim != layout.ACT_DIM:
            raise ValueError("the kernels are built for the reference's 16-D state / 2-D action")
        self.use_vision = bool(getattr(env, "use_vision", False))
        if self.use_vision:
            raise NotImplementedError("the camera path is outside the LiDAR hot path")
        self.vectorized = isinstance(env, VecEnv)
        if self.vectorized and "max_timesteps_per_episode" in hyperparameters and \
                int(self.max_timesteps_per_episode) != int(env.cfg.max_episode_steps):
            # the episode cap lives in the simulator (ppo.py:552 is folded into the step kernel): a trainer that
            # logs one cap while the environment applies another would silently change the episode protocol
            raise ValueError(f"max_timesteps_per_episode={self.max_timesteps_per_episode} but the VecEnv was built with "
                             f"max_episode_steps={int(env.cfg.max_episode_steps)}; pass the same value to both")
        if self.vectorized:
            self.max_timesteps_per_episode = int(env.cfg.max_episode_steps)
        dev = env.device if self.vectorized else getattr(getattr(env, "_vec", None), "device", None)
        self.device = dev if dev is not None else torch.device("cuda", torch.cuda.current_device())

        # ---- output directories, config, episode csv (ppo.py:62-76,149-164)
        out = self.output_dir if self.output_dir is not None else "runs"
        self.method_run_dir = os.path.join(out, self.method_name)
        self.checkpoint_dir = _makepath(os.path.join(self.method_run_dir, "checkpoints"))
        self.log_dir_path = _makepath(os.path.join(self.method_run_dir, "logs"))
        self.tb_dir_path = _makepath(os.path.join(self.method_run_dir, "tb"))
        try:
            import yaml
            with open(os.path.join(self.method_run_dir, "config.yml"), "w") as f:
                yaml.dump(self.config, f, default_flow_style=False)
        except ImportError:
            pass
        self.episode_csv_path = os.path.join(self.log_dir_path, f"{self.method_name}_train_episodes.csv")
        with open(self.episode_csv_path, "w", newline="") as f:
            csv.writer(f).writerow(["episode", "timestep", "success", "collision", "timeout", "length", "return",
                                    "path_length", "time"])
        self.episode_count = 0
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(log_dir=self.tb_dir_path) if self.tensorboard else None
        except Exception:
            self.writer = None

        # ---- networks in one flat vector + flat Adam state
        self.flat = torch.zeros(_capi.PPO_FLAT, dtype=torch.float32, device=self.device)
        self._grad = torch.zeros_like(self.flat)
        self._exp_avg = torch.zeros_like(self.flat)
        self._exp_avg_sq = torch.zeros_like(self.flat)
        self._adam_step = 0
        kw = dict(use_vision=False, vision_feat_dim=1280, vision_proj_dim=64)
        self.actor = policy_class(self.obs_dim, self.act_dim, **kw).to(self.device)     # ppo.py:112
        self.critic = value_func(self.obs_dim, 1, **kw).to(self.device)                 # ppo.py:113
        if not isinstance(self.actor, NetActor) or not isinstance(self.critic, NetCritic):
            raise TypeError("policy_class / value_func must be navbot_ppo_b200.NetActor / NetCritic")
        self.actor.attach(self.flat)
        self.critic.attach(self.flat)
        self.actor_optim = _FlatAdam(self, "actor")        # ppo.py:116
        self.critic_optim = _FlatAdam(self, "critic")      # ppo.py:117
        self.world = navdist.world_size()
        self.rank = navdist.rank()
        self._peer = None                                  # gradient exchange over NVLink peer memory (else NCCL)
        if self.world > 1:
            navdist.broadcast_(self.flat)                  # every rank starts from rank 0's weights

        # ---- exploration covariance (ppo.py:123-124)
        self.cov_var = torch.full((self.act_dim,), 0.8, device=self.device)
        self.cov_mat = torch.diag(self.cov_var)

        n_env = env.num_envs if self.vectorized else 1
        self.horizon = max(1, math.ceil(self.timesteps_per_batch / n_env)) if self.vectorized else self.timesteps_per_batch
        self._max_T = self.horizon * n_env if self.vectorized else self.timesteps_per_batch
        self._h = self._handle_for(self._max_T)
        if self.world > 1:
            self._setup_peer_exchange()
        self._draw = 0                                     # Philox counter of the action noise
        self._stats = torch.zeros(3, dtype=torch.float64, device=self.device)
        self.logger = {"delta_t": time.time(), "t_so_far": 0, "i_so_far": 0, "batch_lens": [], "batch_rews": [],
                       "actor_losses": [], "critic_losses": [], "Episode_Rewards": []}
        self.logger_global = dict(self.logger, Iteration=0)
        self._alloc_rollout_buffers()

    # ------------------------------------------------------------------------------------
    def _init_hyperparameters(self, hyperparameters):
        """ppo.py:752-811, without exec(): unknown keys become attributes, like the reference."""
        self.timesteps_per_batch = 8000
        self.max_timesteps_per_episode = 800
        self.n_updates_per_iteration = 50
        self.lr = 3e-4
        self.gamma = 0.99
        self.clip = 0.2
        self.render = True
        self.render_every_i = 10
        self.save_freq = 2
        self.seed = None
        self.exp_id = "v02_simple_env_60_reward_proportion"
        self.method_name = "baseline"
        self.output_dir = None
        # additions (not in the reference): GEMM arithmetic, quiet logging for benchmarks.  (The reference has no
        # GAE: its advantage is reward-to-go minus V, ppo.py:277; navppo_rtg_scan's (gamma, lambda) form is reachable
        # through the C-ABI only.)
        self.precision = _capi.PREC_FP32
        self.tensorboard = False
        self.verbose = True
        # vectorised rollouts only, not in the reference (which resets at every rollout, ppo.py:486, and never
        # bootstraps, :601): keep episodes running across rollouts / bootstrap the cut-off tail from the critic
        self.continue_episodes = False
        self.bootstrap_value = False
        self.graph_rollout = True         # vectorised rollouts: replay the step loop as one CUDA graph
        self.gradient_exchange = "peer"   # multi-GPU: "peer" | "nvls" | "nccl" (see _setup_peer_exchange)
        self.log_episodes = True          # vectorised rollouts: one csv row per completed episode ...
        self.max_logged_episodes = 4096   # ... up to this many per iteration (None: all)
        for k, v in hyperparameters.items():
            setattr(self, k, v)
        self.config = {k: getattr(self, k) for k in
                       ("timesteps_per_batch", "max_timesteps_per_episode", "n_updates_per_iteration", "lr", "gamma",
                        "clip", "render", "render_every_i", "save_freq", "seed", "exp_id", "method_name", "output_dir")}
        if self.seed is not None:                          # ppo.py:805-811
            assert type(self.seed) == int
            torch.manual_seed(self.seed)

    def _setup_peer_exchange(self):
        """Per-epoch gradient exchange fused into the optimiser step: every rank's flat gradient lives in
        symmetric (peer-mapped) memory, and the Adam kernel adds the ranks' buffers itself, straight over
        NVLink (navppo_adam_peer).  gradient_exchange = "peer" (default: ordered peer loads, identical bits on
        every rank), "nvls" (multimem.ld_reduce on the multicast address) or "nccl" (all-reduce + Adam, the
        fallback and the correctness reference).  Anything the platform lacks falls back to NCCL."""
        mode = os.environ.get("NAVPPO_GRAD_EXCHANGE", self.gradient_exchange)
        if mode == "nccl" or self.device.type != "cuda":
            return
        try:
            import torch.distributed as td
            import torch.distributed._symmetric_memory as symm
            F = _capi.PPO_FLAT
            buf = symm.empty(2 * F + 16, dtype=torch.float32, device=self.device)
            hdl = symm.rendezvous(buf, group=td.group.WORLD)
            buf.zero_()
            torch.cuda.synchronize(self.device)
            hdl.barrier()
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            W, R = int(hdl.world_size), int(hdl.rank)
            grad_ptrs = (ctypes.c_uint64 * (2 * W))(*[ptrs[r] + b * F * 4 for b in range(2) for r in range(W)])
            flag_ptrs = (ctypes.c_uint64 * W)(*[ptrs[r] + 2 * F * 4 for r in range(W)])
            mc = int(hdl.multicast_ptr) if getattr(hdl, "has_multicast_support", lambda *a: False) and hdl.multicast_ptr else 0
            mc_ptrs = (ctypes.c_uint64 * 2)(mc, mc + F * 4) if mc else None
            if mode == "nvls" and not mc:
                mode = "peer"
            self._peer = dict(buf=buf, hdl=hdl, grads=[buf[:F], buf[F:2 * F]], epoch=0, nvls=(mode == "nvls"),
                              tables=(R, W, grad_ptrs, flag_ptrs, mc_ptrs), bound=None)
            self._bind_peer()
        except Exception as e:  # noqa: BLE001 - no peer access / no symmetric memory: NCCL does the exchange
            if self.verbose and self.rank == 0:
                print(f"[PPO] peer-memory gradient exchange unavailable ({type(e).__name__}: {e}); using NCCL", flush=True)
            self._peer = None

    def _bind_peer(self, force=False):
        """Hand the peer pointer tables to the navppo handle in use (handles are re-picked when the batch grows)."""
        if self._peer is not None and (force or self._peer["bound"] != self._h.value):
            R, W, grad_ptrs, flag_ptrs, mc_ptrs = self._peer["tables"]
            _capi.check(_capi.lib().navppo_peer_setup(self._h, R, W, grad_ptrs, flag_ptrs, mc_ptrs))
            self._peer["bound"] = self._h.value

    def _alloc_rollout_buffers(self):
        if not self.vectorized:
            return
        H, N, d = self.horizon, self.env.num_envs, self.device
        self._b_obs = torch.empty((H, N, layout.OBS_DIM), dtype=torch.float32, device=d)
        self._b_act = torch.empty((H, N, 2), dtype=torch.float32, device=d)
        self._b_logp = torch.empty((H, N), dtype=torch.float32, device=d)
        self._b_rew = torch.empty((H, N), dtype=torch.float32, device=d)
        self._b_flags = torch.empty((3, H, N), dtype=torch.uint8, device=d)   # done, arrive, timeout per step
        self._b_term = torch.empty((H, N), dtype=torch.uint8, device=d)
        self._b_rtg = torch.empty((H, N), dtype=torch.float32, device=d)
        self._b_epret = torch.zeros((H, N), dtype=torch.float32, device=d)    # return / path length of the episode
        self._b_eppath = torch.zeros((H, N), dtype=torch.float32, device=d)   # that ends at [t, n] (ppo.py:739-746)
        self._b_eplen = torch.zeros((H, N), dtype=torch.int32, device=d)      # and its length in steps (ppo.py:583)
        self._dyn = torch.zeros(2, dtype=torch.int32, device=d)               # {float bits of var, noise-counter increment}
        self._dyn_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self._rollout_graph = None
        self._next_obs = torch.empty((N, layout.OBS_DIM), dtype=torch.float32, device=d)

    def _handle_for(self, T: int):
        """navppo handle whose gradient workspace covers T samples."""
        self._max_T = max(self._max_T, int(T), 1)
        self._h = _Handles.get(self.device, self._max_T, self.clip, self.lr, self.precision)
        if getattr(self, "_peer", None) is not None:
            self._bind_peer()
        return self._h

    @property
    def var(self) -> float:
        return float(self.cov_var[0].item())

    def _decay_cov(self):
        self.cov_mat *= 0.995                              # ppo.py:695
        self.cov_var = torch.diagonal(self.cov_mat).clone()

    # ------------------------------------------------------------------------------------ P1
    def get_action(self, obs, t_so_far, one_round, vision_feat=None, noise=None):
        """ppo.py:673-706 for one observation -> (action np[2], log_prob np scalar)."""
        self.t_step = one_round
        if self.t_step == 0 and t_so_far > 50000 and self.cov_mat[0][0] >= 0.1:     # ppo.py:694-695
            self._decay_cov()
        o = torch.as_tensor(np.asarray(obs, dtype=np.float32)).reshape(1, -1).to(self.device)
        act, logp = self.act_batch(o, noise=None if noise is None else
                                   torch.as_tensor(np.asarray(noise, dtype=np.float32)).reshape(1, 2).to(self.device))
        return act.cpu().numpy()[0], logp.cpu().numpy()[0]

    def act_batch(self, obs: torch.Tensor, noise: torch.Tensor | None = None, out_act=None, out_logp=None,
                  agent_id_offset: int = 0):
        """navppo_act on a device batch: (actions [N,2], log-probs [N])."""
        n = obs.shape[0]
        act = out_act if out_act is not None else torch.empty((n, 2), dtype=torch.float32, device=self.device)
        logp = out_logp if out_logp is not None else torch.empty(n, dtype=torch.float32, device=self.device)
        seed = 0 if self.seed is None else int(self.seed)
        _capi.check(_capi.lib().navppo_act(self._h, self.flat.data_ptr(), obs.data_ptr(), n, self.var, seed,
                                           agent_id_offset, self._draw, None if noise is None else noise.data_ptr(),
                                           act.data_ptr(), logp.data_ptr(), None, _stream(self.device)))
        self._draw += 1
        return act, logp

    # ------------------------------------------------------------------------------------ P4
    def evaluate(self, batch_obs, batch_acts, batch_vision_feats=None):
        """ppo.py:708-737 -> (V [T], log_probs [T])."""
        obs = batch_obs.to(device=self.device, dtype=torch.float32).reshape(-1, layout.OBS_DIM).contiguous()
        act = batch_acts.to(device=self.device, dtype=torch.float32).reshape(-1, 2).contiguous()
        T = obs.shape[0]
        v = torch.empty(T, dtype=torch.float32, device=self.device)
        logp = torch.empty(T, dtype=torch.float32, device=self.device)
        _capi.check(_capi.lib().navppo_evaluate(self._h, self.flat.data_ptr(), obs.data_ptr(), act.data_ptr(), T, self.var,
                                                v.data_ptr(), logp.data_ptr(), _stream(self.device)))
        self.V = v
        return v, logp

    # ------------------------------------------------------------------------------------ P3
    def compute_rtgs(self, batch_rews):
        """ppo.py:643-671 on the reference's ragged list-of-episodes: laid out as one column with
        episode-end flags and scanned on the device."""
        lens = [len(e) for e in batch_rews]
        flat = np.asarray([r for e in batch_rews for r in e], dtype=np.float32)
        if flat.size == 0:
            return torch.zeros(0, dtype=torch.float)
        term = np.zeros(flat.size, np.uint8)
        ends = np.cumsum([l for l in lens if l > 0]) - 1
        term[ends] = 1
        rew = torch.from_numpy(flat).to(self.device)
        term_d = torch.from_numpy(term).to(self.device)
        out = torch.empty_like(rew)
        _capi.check(_capi.lib().navppo_rtg_scan(rew.data_ptr(), term_d.data_ptr(), None,
                                                None, float(self.gamma), 1.0, out.data_ptr(), flat.size, 1,
                                                _stream(self.device)))
        return out.cpu()

    # ------------------------------------------------------------------------------------ P2
    def rollout(self, past_action, t_so_far):
        if self.vectorized:
            return self._rollout_vec(t_so_far)
        return self._rollout_single(past_action, t_so_far)

    def _rollout_single(self, past_action, t_so_far):
        """The reference's loop, line for line (ppo.py:476-641), over the one-robot Env."""
        batch_obs, batch_acts, batch_log_probs, batch_rews, batch_lens = [], [], [], [], []
        obs = self.env.reset()                                                   # :486
        episode_reward, one_round, ep_rews = 0, 0, []
        ep_path_length, prev_pos, ep_start = 0.0, None, time.time()
        it = dict(successes=0, collisions=0, timeouts=0, ep_times=[], ep_count=0)
        for t in range(self.timesteps_per_batch):                                # :505
            batch_obs.append(obs)
            assert len(obs) == 16, f"Base state must be 16-d, got {len(obs)} at timestep {t_so_far + t}"
            action, log_prob = self.get_action(obs, t_so_far, one_round)         # :532
            curr_pos = np.array([self.env.position.x, self.env.position.y])      # :535-538
            if prev_pos is not None:
                ep_path_length += np.linalg.norm(curr_pos - prev_pos)
            prev_pos = curr_pos
            obs, rew, done, arrive = self.env.step(action, past_action)          # :541
            past_action = action
            episode_reward += rew
            ep_rews.append(rew)
            batch_acts.append(action)
            batch_log_probs.append(log_prob)
            one_round += 1
            timeout = one_round >= self.max_timesteps_per_episode                # :552
            if done or arrive or timeout:
                success = 1 if arrive else 0                                     # :558-560
                collision = 1 if (done and not arrive) else 0
                timeout_flag = 1 if (timeout and not done and not arrive) else 0
                ep_time = time.time() - ep_start
                self._log_episode_metrics(self.episode_count, t_so_far + np.sum(batch_lens) + one_round, success,
                                          collision, timeout_flag, one_round, episode_reward, ep_path_length, ep_time)
                self.episode_count += 1
                it["successes"] += success; it["collisions"] += collision; it["timeouts"] += timeout_flag
                it["ep_times"].append(ep_time); it["ep_count"] += 1
                batch_lens.append(one_round)
                batch_rews.append(ep_rews)
                ep_rews = []
                self.logger["Episode_Rewards"].append(episode_reward / one_round)
                episode_reward, one_round, ep_path_length, prev_pos = 0, 0, 0.0, None
                past_action = [0, 0]                                             # :591
                obs = self.env.reset()                                           # :593
                ep_start = time.time()
        batch_rews.append(ep_rews)                                               # :601 trailing partial episode
        b_obs = torch.from_numpy(np.array(batch_obs, dtype=np.float32))
        b_acts = torch.from_numpy(np.array(batch_acts, dtype=np.float32))
        b_logp = torch.from_numpy(np.array(batch_log_probs, dtype=np.float32))
        b_rtgs = self.compute_rtgs(batch_rews)                                   # :619
        self.logger["batch_rews"], self.logger["batch_lens"] = batch_rews, batch_lens
        d = self.device
        return b_obs.to(d), b_acts.to(d), b_logp.to(d), b_rtgs.to(d), batch_lens, it, None

    def _rollout_vec(self, t_so_far):
        """N robots x H steps on the device.  Episode protocol (reset, zero past action, per-agent
        one_round counter, outcome precedence) is the step kernel's auto-reset (ppo.py:549-593)."""
        env, H, N = self.env, self.horizon, self.env.num_envs
        L = _capi.lib()
        sp = _stream(self.device)
        if t_so_far > 50000 and self.cov_mat[0][0] >= 0.1:   # ppo.py:694-695, once per rollout (episode starts)
            self._decay_cov()
        env.clear_stats()                                    # in stream order: no read-back, no synchronisation
        if self.continue_episodes and getattr(self, "_rollouts_done", 0) > 0:
            self._b_obs[0].copy_(self._next_obs)                                 # episodes run on across rollouts
        else:
            env.reset(out=self._b_obs[0])                                        # ppo.py:486
        self._rollouts_done = getattr(self, "_rollouts_done", 0) + 1
        seed = 0 if self.seed is None else int(self.seed)
        # The step loop of ppo.py:505-549 is enqueued by the library (2 H launches, no Python in between) ONCE, into a
        # CUDA graph; every rollout then replays it.  What changes between rollouts (the exploration variance and the
        # Philox counter of the action noise) is read by the sampling epilogue from two device words.
        self._dyn_host[0] = int(np.float32(self.var).view(np.int32))
        self._dyn_host[1] = int(np.uint32(self._draw & 0xFFFFFFFF).view(np.int32))
        self._dyn.copy_(self._dyn_host, non_blocking=True)

        def enqueue(stream_ptr):
            _capi.check(L.navppo_rollout_ex(self._h, env._h, self.flat.data_ptr(), H, 1.0, seed, int(env.cfg.agent_id_offset), 0,
                                            self._b_obs.data_ptr(), self._next_obs.data_ptr(), self._b_act.data_ptr(),
                                            self._b_logp.data_ptr(), self._b_rew.data_ptr(), self._b_flags[0].data_ptr(),
                                            self._b_flags[1].data_ptr(), self._b_flags[2].data_ptr(), self._b_epret.data_ptr(),
                                            self._b_eppath.data_ptr(), self._b_eplen.data_ptr(), self._dyn.data_ptr(), stream_ptr))

        if not self.graph_rollout:
            enqueue(sp)
        else:
            if self._rollout_graph is None:
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side):
                        enqueue(side.cuda_stream)
                torch.cuda.current_stream(self.device).wait_stream(side)
                self._rollout_graph = g
            self._rollout_graph.replay()
        self._draw += H
        torch.amax(self._b_flags, dim=0, out=self._b_term)                       # done | arrive | timeout, ppo.py:553
        last_v = None
        if self.bootstrap_value:
            with torch.no_grad():
                last_v = self.critic(self._next_obs).reshape(-1).contiguous()
        _capi.check(L.navppo_rtg_scan(self._b_rew.data_ptr(), self._b_term.data_ptr(), None,
                                      None if last_v is None else last_v.data_ptr(), float(self.gamma), 1.0,
                                      self._b_rtg.data_ptr(), H, N, sp))
        st = env.stats()
        it = dict(successes=int(st.successes), collisions=int(st.collisions), timeouts=int(st.timeouts), ep_times=[],
                  ep_count=int(st.episodes), return_sum=float(st.return_sum), length_sum=float(st.length_sum),
                  path_sum=float(st.path_sum))
        # completed-episode lengths (ppo.py:583): the simulator writes an episode's length at its last step, so the
        # list is a device-side compaction of [H, N] at the episode-end flags; only the completed episodes' few
        # integers cross to the host (agent-major order, like the one-robot loop playing the agents in turn)
        term_t = self._b_term.t()
        ends = term_t.nonzero(as_tuple=False)              # [E, 2] = (agent, step), sorted by agent then time
        if ends.shape[0]:
            nn_d, tt_d = ends[:, 0], ends[:, 1]
            batch_lens = self._b_eplen[tt_d, nn_d].cpu().numpy().astype(np.int64)
        else:
            nn_d = tt_d = None
            batch_lens = np.zeros(0, np.int64)
        self.logger["batch_lens"] = batch_lens
        self.logger["batch_rews"] = []
        if self.log_episodes and batch_lens.size and self.rank == 0:
            self._log_vec_episodes(t_so_far, nn_d, tt_d, batch_lens)
        T = H * N
        return (self._b_obs.view(T, -1), self._b_act.view(T, 2), self._b_logp.view(T), self._b_rtg.view(T), batch_lens,
                it, None)

    # ------------------------------------------------------------------------------------ P5-P7
    def update(self, batch_obs, batch_acts, batch_log_probs, batch_rtgs, epochs=None):
        """The update of one PPO.learn iteration (ppo.py:275-397).  Returns a dict of the
        per-epoch metric arrays; nothing is read back from the device until all epochs are
        enqueued."""
        epochs = self.n_updates_per_iteration if epochs is None else epochs
        d = self.device
        obs = batch_obs.to(device=d, dtype=torch.float32).reshape(-1, layout.OBS_DIM).contiguous()
        act = batch_acts.to(device=d, dtype=torch.float32).reshape(-1, 2).contiguous()
        logp_old = batch_log_probs.to(device=d, dtype=torch.float32).reshape(-1).contiguous()
        rtg = batch_rtgs.to(device=d, dtype=torch.float32).reshape(-1).contiguous()
        T = obs.shape[0]
        self._handle_for(T)
        L, sp = _capi.lib(), _stream(d)
        adv = torch.empty(T, dtype=torch.float32, device=d)
        v = torch.empty(T, dtype=torch.float32, device=d)
        metrics = torch.zeros((max(epochs, 1), _capi.PPO_NUM_METRICS), dtype=torch.float64, device=d)
        before = self.flat.clone()
        if self.world == 1:
            _capi.check(L.navppo_update(self._h, self.flat.data_ptr(), self._exp_avg.data_ptr(), self._exp_avg_sq.data_ptr(),
                                        self._adam_step, obs.data_ptr(), act.data_ptr(), logp_old.data_ptr(),
                                        rtg.data_ptr(), T, self.var, epochs, adv.data_ptr(), v.data_ptr(),
                                        metrics.data_ptr(), sp))
        else:
            _capi.check(L.navppo_evaluate(self._h, self.flat.data_ptr(), obs.data_ptr(), act.data_ptr(), T, self.var,
                                          v.data_ptr(), adv.data_ptr(), sp))
            self._stats.zero_()
            _capi.check(L.navppo_adv_stats(rtg.data_ptr(), v.data_ptr(), T, self._stats.data_ptr(), sp))
            navdist.all_reduce_sum_(self._stats)                       # global mean / std (ppo.py:284)
            n_global = int(round(float(self._stats[2].item())))
            _capi.check(L.navppo_adv_normalize(rtg.data_ptr(), v.data_ptr(), T, self._stats.data_ptr(), adv.data_ptr(), sp))
            self._bind_peer(force=True)        # (handles are shared between trainers of one process)
            for e in range(epochs):
                row = metrics[e]
                if self._peer is not None:
                    # the rank's gradient goes straight into its peer-mapped buffer; the optimiser step adds the
                    # ranks' buffers over NVLink (one barrier + one kernel instead of ncclAllReduce + Adam)
                    b = self._peer["epoch"] & 1
                    _capi.check(L.navppo_grad(self._h, self.flat.data_ptr(), obs.data_ptr(), act.data_ptr(),
                                              logp_old.data_ptr(), adv.data_ptr(), rtg.data_ptr(), T, n_global, self.var,
                                              self._peer["grads"][b].data_ptr(), row.data_ptr(), sp))
                    self._peer["epoch"] += 1
                    _capi.check(L.navppo_adam_peer(self._h, self.flat.data_ptr(), self._exp_avg.data_ptr(),
                                                   self._exp_avg_sq.data_ptr(), self._adam_step + e + 1, b,
                                                   self._peer["epoch"], 1 if self._peer["nvls"] else 0, row.data_ptr(), sp))
                    continue
                _capi.check(L.navppo_grad(self._h, self.flat.data_ptr(), obs.data_ptr(), act.data_ptr(),
                                          logp_old.data_ptr(), adv.data_ptr(), rtg.data_ptr(), T, n_global, self.var,
                                          self._grad.data_ptr(), row.data_ptr(), sp))
                navdist.all_reduce_sum_(self._grad)                    # the one collective per epoch
                _capi.check(L.navppo_adam(self._h, self.flat.data_ptr(), self._grad.data_ptr(), self._exp_avg.data_ptr(),
                                          self._exp_avg_sq.data_ptr(), self._adam_step + e + 1, row.data_ptr(), sp))
            shares = metrics[:, :4].contiguous()           # each rank holds its share of the four batch means
            navdist.all_reduce_sum_(shares)
            metrics[:, :4] = shares
        self._adam_step += epochs
        self.V = v
        m = metrics.cpu().numpy()[:epochs]
        a_n, c_n = layout.ACTOR_PARAMS, _capi.PPO_CRITIC_OFFSET
        delta = self.flat - before
        return dict(actor_losses=m[:, _capi.M_ACTOR_LOSS], critic_losses=m[:, _capi.M_CRITIC_LOSS],
                    approx_kl=m[:, _capi.M_APPROX_KL], clip_frac=m[:, _capi.M_CLIP_FRAC],
                    actor_grad_norm=np.sqrt(m[:, _capi.M_ACTOR_GRAD_SQ]), critic_grad_norm=np.sqrt(m[:, _capi.M_CRITIC_GRAD_SQ]),
                    actor_param_delta=float(torch.linalg.vector_norm(delta[:a_n]).item()),
                    critic_param_delta=float(torch.linalg.vector_norm(delta[c_n:]).item()),
                    v_mean=float(v.mean().item()), adv=adv)

    # ------------------------------------------------------------------------------------
    def learn(self, total_timesteps, past_action=(0, 0)):
        """ppo.py:218-461."""
        if self.verbose and self.rank == 0:
            print(f"Learning... Running {self.max_timesteps_per_episode} timesteps per episode, "
                  f"{self.timesteps_per_batch} timesteps per batch for a total of {total_timesteps} timesteps", flush=True)
        t_so_far, i_so_far = 0, 0
        while t_so_far < total_timesteps:                                        # :245
            t0 = time.time()
            obs, acts, logp, rtgs, lens, iter_metrics, _ = self.rollout(past_action=list(past_action), t_so_far=t_so_far)
            torch.cuda.synchronize(self.device)
            t1 = time.time()
            if self.vectorized:
                # every simulated step counts: H x N per rank.  (The reference adds up completed-episode lengths, :258,
                # which equals the batch size on one robot because an episode is far shorter than a batch; with N
                # robots x H steps most episodes straddle rollouts and would never be counted.)
                steps = int(obs.shape[0]) * self.world
            else:
                steps = int(np.sum(lens))
            t_so_far += steps                                                    # :258
            i_so_far += 1
            self.logger.update(t_so_far=t_so_far, i_so_far=i_so_far, iter_metrics=iter_metrics)
            res = self.update(obs, acts, logp, rtgs)                             # :275-397
            torch.cuda.synchronize(self.device)
            t2 = time.time()
            with open(os.path.join(self.log_dir_path, "V_fun.txt"), "a+") as f:  # :278-282 (one line per iteration)
                f.write(f"{res['v_mean']}\n")
            self.logger["actor_losses"] = list(res["actor_losses"])
            self.logger["critic_losses"] = list(res["critic_losses"])
            for k in ("approx_kl", "clip_frac", "actor_grad_norm", "critic_grad_norm"):
                self.logger[k] = float(np.mean(res[k])) if len(res[k]) else 0.0
            # the reference logs the entropy of ONE T-dimensional Gaussian (ppo.py:330-331)
            T = int(obs.shape[0])
            self.logger["entropy"] = 0.5 * T * (1.0 + math.log(2.0 * math.pi)) + math.log(self.var)
            self.logger["actor_param_delta"] = res["actor_param_delta"]
            self.logger["critic_param_delta"] = res["critic_param_delta"]
            self.logger.update(rollout_time=t1 - t0, update_time=t2 - t1, iter_time=time.time() - t0,
                               actor_grad_steps=self.n_updates_per_iteration, critic_grad_steps=self.n_updates_per_iteration,
                               steps_per_sec=T * self.world / max(time.time() - t0, 1e-9))
            self._log_summary()
            if i_so_far % self.save_freq == 0 and self.rank == 0:                # :452-457
                self.save_checkpoint(i_so_far, t_so_far)
        return t_so_far

    def save_checkpoint(self, i_so_far, t_so_far):
        a = os.path.join(self.checkpoint_dir, f"actor_iter{i_so_far:04d}_step{t_so_far:08d}.pth")
        c = os.path.join(self.checkpoint_dir, f"critic_iter{i_so_far:04d}_step{t_so_far:08d}.pth")
        torch.save({k: v.detach().cpu().clone() for k, v in self.actor.state_dict().items()}, a)
        torch.save({k: v.detach().cpu().clone() for k, v in self.critic.state_dict().items()}, c)
        if self.verbose:
            print(f"[PPO] Saved checkpoint at iteration {i_so_far}, step {t_so_far}: {a}", flush=True)
        return a, c

    def _log_episode_metrics(self, episode_num, timestep, success, collision, timeout, length, ep_return, path_length,
                             ep_time):
        """ppo.py:739-750: same CSV columns."""
        with open(self.episode_csv_path, "a", newline="") as f:
            csv.writer(f).writerow([episode_num, timestep, success, collision, timeout, length, ep_return, path_length,
                                    ep_time])

    def _log_vec_episodes(self, t_so_far, agents, steps, lengths):
        """One csv row per episode completed in this rollout, the reference's columns (ppo.py:739-746).
        Episodes are listed agent by agent; `timestep` counts completed steps the way the one-robot
        loop would if it played the agents one after another; `time` is simulated time (0.2 s per
        step, gazebo.xacro:107) because N robots share the wall clock.  `agents` / `steps` are device
        index tensors: only the logged episodes' values are gathered and copied."""
        n = len(lengths)
        if self.max_logged_episodes is not None and n > self.max_logged_episodes:
            n = int(self.max_logged_episodes)
        a_, s_ = agents[:n], steps[:n]
        arrive = self._b_flags[1][s_, a_].cpu().numpy().astype(bool)
        done = self._b_flags[0][s_, a_].cpu().numpy().astype(bool)
        ret = self._b_epret[s_, a_].cpu().numpy()
        path = self._b_eppath[s_, a_].cpu().numpy()
        ts = t_so_far + np.cumsum(lengths)
        with open(self.episode_csv_path, "a", newline="") as f:
            w = csv.writer(f)
            for e in range(n):
                w.writerow([self.episode_count + e, int(ts[e]), int(arrive[e]), int(done[e] and not arrive[e]),
                            int(not done[e] and not arrive[e]), int(lengths[e]), float(ret[e]), float(path[e]),
                            0.2 * float(lengths[e])])
        self.episode_count += len(lengths)

    def _log_summary(self):
        """ppo.py:813-946 (condensed): stdout block + the same TensorBoard tags."""
        lg = self.logger
        im = lg.get("iter_metrics", {})
        n_ep = max(1, im.get("ep_count", 0))
        if self.vectorized:
            avg_len = im.get("length_sum", 0.0) / n_ep
            avg_ret = im.get("return_sum", 0.0) / n_ep
        else:
            avg_len = float(np.mean(lg["batch_lens"])) if len(lg["batch_lens"]) else 0.0
            avg_ret = float(np.mean([np.sum(e) for e in lg["batch_rews"] if len(e)])) if lg["batch_rews"] else 0.0
        a_loss = float(np.mean(lg["actor_losses"])) if len(lg["actor_losses"]) else 0.0
        c_loss = float(np.mean(lg["critic_losses"])) if len(lg["critic_losses"]) else 0.0
        scal = {"train/avg_episode_length": avg_len, "train/avg_episode_return": avg_ret, "loss/actor": a_loss,
                "loss/critic": c_loss, "ppo/approx_kl": lg["approx_kl"], "ppo/clip_frac": lg["clip_frac"],
                "ppo/entropy": lg["entropy"], "grad/actor_norm": lg["actor_grad_norm"],
                "grad/critic_norm": lg["critic_grad_norm"], "perf/steps_per_sec": lg["steps_per_sec"],
                "metrics/success_rate": im.get("successes", 0) / n_ep, "metrics/collision_rate": im.get("collisions", 0) / n_ep,
                "metrics/timeout_rate": im.get("timeouts", 0) / n_ep}
        lg["summary"] = scal
        if self.writer is not None and self.rank == 0:
            for k, v in scal.items():
                self.writer.add_scalar(k, v, lg["t_so_far"])
            self.writer.flush()
        if self.verbose and self.rank == 0:
            print(f"-------------------- Iteration #{lg['i_so_far']} --------------------\n"
                  f"Average Episodic Length: {avg_len:.2f}\nAverage Episodic Return: {avg_ret:.2f}\n"
                  f"Average Actor Loss: {a_loss:.5f}\nAverage Critic Loss: {c_loss:.5f}\n"
                  f"Timesteps So Far: {lg['t_so_far']}\nRollout {lg['rollout_time']:.3f}s  Update {lg['update_time']:.3f}s  "
                  f"Steps/sec {lg['steps_per_sec']:.1f}\n"
                  f"Success {scal['metrics/success_rate']:.3f}  Collision {scal['metrics/collision_rate']:.3f}  "
                  f"Timeout {scal['metrics/timeout_rate']:.3f}\n"
                  f"------------------------------------------------------", flush=True)
        lg["batch_lens"], lg["batch_rews"], lg["actor_losses"], lg["critic_losses"] = [], [], [], []
