#!/usr/bin/env python
"""bench.py — headline benchmark: env-steps/s of the stage_1 hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" of this benchmark is one pass of the hot path over one batch: H consecutive
Env.step calls for every one of the 8192 agents of a rank (BASELINE.json configs[1];
agents shard over ranks with no data-path collective -> weak scaling), issued as ONE launch
of the step kernel (navsim_rollout_scripted) that writes every step's observation, reward
and flags into the [H, N, .] rollout buffers.  Rank 0 prints ONE JSON line.  See DESIGN.md
"Measurement" for every figure's definition.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# Algorithmic HBM bytes per env-step of navsim_step_kernel (DESIGN.md "Data layout"):
#   read : pose 24 + goal 16 + past_dist 8 + prev_action 8 + ep stats 12 + steps 4 + draws 4 + action 8 = 84
#   write: pose 24 + past_dist 8 + prev_action 8 + ep stats 12 + steps 4 + obs 64 + rew 4 + flags 3     = 127
BYTES_PER_ENV_STEP = 84 + 127
BYTES_PER_ENV_STEP_SCRIPTED = BYTES_PER_ENV_STEP - 8  # actions drawn in-kernel, not read
# SURVEY.md 8(d), the contract figure: fp32 SoA 44 B read + 106 B written = 150 B (174 B with the pose kept in
# fp64, which is what this simulator stores); the 203 B above adds what the kernel also moves per step (episode
# return / path length / last move, draw counter, timeout flag, goal as fp64).  All three are reported.
BYTES_SURVEY = 150
BYTES_SURVEY_FP64_POSE = 174
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed `ncu --set full`
# capture of the same kernel and shape (profiles/README.md names the file); keyed by (N, H)
NCU_TRAFFIC_BYTES_PER_LAUNCH = {(8192, 128): 741632 + 18013952}   # profiles/r02z_step_8192_fused_ncu.csv


# FP32 work of one env-step on the house map by beam count: (fadd + fmul + 2 ffma) thread-level instruction counts of
# one fused launch from `ncu --metrics smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on.sum`
# (profiles/r02_c5_house_flops.txt), divided by the env-steps of that launch
NCU_FP32_FLOP_PER_ENV_STEP = {("house", 10): 4396.0, ("house", 12): 9012.0, ("house", 18): 9120.0, ("house", 24): 9210.0,
                              ("house", 36): 9434.0}
# FP32 ALU peak the c5 figures are quoted against: 148 SMs x 128 lanes x 2 flop x 1.965 GHz (nominal)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML in a thread DURING the timed region."""

    def __init__(self, index: int, period_s: float = 0.002):
        self.index, self.period, self.rows, self._stop, self._thr, self.err = index, period_s, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; map to the NVML index when it is set
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = f"nvml unavailable: {e}"
            return
        self._thr = threading.Thread(target=self._pump, daemon=True)
        self._thr.start()

    def _pump(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((sm, rs))
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                return
            time.sleep(self.period)

    def mark(self):
        return len(self.rows)

    def stop(self, lo: int = 0, hi: int | None = None):
        self._stop = True
        if self._thr is not None:
            self._thr.join(timeout=1)
        if self.err and not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err]}
        rows = self.rows[lo:hi] or self.rows
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80}
        reasons = sorted(k for k, b in bits.items() if any(r & b for _, r in rows))
        return {"sm_mhz": float(np.median([c for c, _ in rows])), "sm_max_mhz": self.sm_max, "samples": len(rows),
                "reasons": reasons}


def workload_string(n_per_gpu: int, horizon: int) -> str:
    """config.workload, identical in both arms."""
    return (f"stage_1 map, {n_per_gpu} agents/GPU, 10-beam LiDAR, 16-D obs / 2-D action, {horizon} env steps per bench "
            f"step, scripted Philox actions, auto-reset episodes (cap 500)")


class CpuPort:
    """The C restatement of the reference path (oracle/navsim_oracle.c) on host threads: test
    infrastructure, used here only as the reported CPU baseline / reference arm.  Nothing of the product
    package is imported on this path (the stage_1 geometry and the reference constants are restated in
    oracle/)."""

    def __init__(self, n_agents: int, threads: int):
        from oracle import binding
        self.n, self.threads = n_agents, threads
        cfg = binding.default_cfg(n_agents)
        cfg.seed = 0
        self.sim = binding.OracleSim(cfg, binding.stage_1_segments(), nthreads=threads)
        self.sim.reset()
        self.step0 = 0

    def run(self, nsteps: int) -> float:
        """nsteps consecutive Env.step calls for every agent (scripted actions, in C); seconds taken."""
        t0 = time.perf_counter()
        self.sim.run_scripted(nsteps, action_seed=0, step0=self.step0)
        self.step0 += nsteps
        return time.perf_counter() - t0


def cpu_port_throughput(n_agents: int, seconds: float, threads: int, chunk: int = 32):
    port = CpuPort(n_agents, threads)
    port.run(4)
    steps, dt = 0, 0.0
    while dt < seconds:
        dt += port.run(chunk)
        steps += chunk
    return n_agents * steps / dt, steps, dt


def tensor_roofline(useful_tflops, precision):
    """The update against the measured bf16 tensor peak (sustained: the kernel runs for ~175 ms): `useful` = the
    fp32-equivalent network flops per second the update delivers, `executed` = the bf16 MMA flops behind them (three
    passes per product in bf16x3: lo*hi + hi*lo + hi*hi), without the padding of the skinny products."""
    if precision == "fp32":
        return None
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, src = 1353.6, "fallback (this pool's sustained cuBLAS bf16 figure)"
    if os.path.exists(p):
        d = json.load(open(p))
        peak, src = float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", peak))), "measured (MEASURED_PEAKS.json, sustained)"
    executed = useful_tflops * (3.0 if precision == "bf16x3" else 1.0)
    return {"bound": "tensor", "useful": useful_tflops, "executed": executed, "peak": peak, "unit": "TFLOP/s",
            "frac_useful": useful_tflops / peak, "frac_executed": executed / peak, "peak_source": src,
            "note": "K = 16 / 32 and N = 16 .. 128 on every product: the MMAs are operand-fetch bound and the epilogues "
                    "(bias, LeakyReLU, bf16 hi | lo split) are CUDA-core work of the same order (DESIGN.md 6)"}


def training_bench(args, torch, dist, dev, rank, world, barrier):
    """env-steps/s of (ii) the rollout and (iii) the whole PPO iteration at N agents x H steps per GPU."""
    import tempfile

    from navbot_ppo_b200 import _capi
    from navbot_ppo_b200.env import VecEnv
    from navbot_ppo_b200.nets import NetActor, NetCritic
    from navbot_ppo_b200.ppo import PPO
    N, H = args.agents, args.horizon
    prec = {"fp32": _capi.PREC_FP32, "bf16x3": _capi.PREC_BF16X3, "bf16": _capi.PREC_BF16}[args.precision_one]
    env = VecEnv(N, map="stage_1", device=dev.index, seed=0, max_episode_steps=500, agent_id_offset=rank * N)
    with tempfile.TemporaryDirectory() as tmp:
        agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=N * H, max_timesteps_per_episode=500,
                    n_updates_per_iteration=args.epochs, gamma=0.99, lr=3e-4, clip=0.2, seed=0, output_dir=tmp,
                    method_name=f"bench{rank}", verbose=False, precision=prec, log_episodes=False)
        l0 = env.launch_count
        batch = agent.rollout([0, 0], 0)            # warm-up iteration
        agent.update(*batch[:4])
        launches_per_iter = env.launch_count - l0
        barrier()
        ro_ms, up_ms = 0.0, 0.0
        res = None
        for _ in range(args.train_iters):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            batch = agent.rollout([0, 0], 0)
            e[1].record()
            res = agent.update(*batch[:4])
            e[2].record()
            torch.cuda.synchronize()
            ro_ms += e[0].elapsed_time(e[1]); up_ms += e[1].elapsed_time(e[2])
        barrier()
        t = torch.tensor([ro_ms, up_ms, ro_ms + up_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ro_ms, up_ms, tot_ms = (float(x) / args.train_iters for x in t)
    steps = world * N * H
    flops = 3.0 * (98432 + 98368) * N * H * args.epochs       # fwd + bwd of both networks, SURVEY 8(d)
    return {"rollout_env_steps_per_s": steps / (ro_ms * 1e-3), "train_env_steps_per_s": steps / (tot_ms * 1e-3),
            "rollout_ms": ro_ms, "update_ms": up_ms, "iteration_ms": tot_ms, "epochs": args.epochs, "horizon": H,
            "samples_per_gpu": N * H, "precision": args.precision_one, "episode_csv": False,
            "update_tflops_per_gpu": flops / (up_ms * 1e-3) / 1e12,
            "update_tensor_roofline": tensor_roofline(flops / (up_ms * 1e-3) / 1e12, args.precision_one),
            "final_actor_loss": float(res["actor_losses"][-1]), "final_critic_loss": float(res["critic_losses"][-1]),
            "sim_launches_per_iteration": int(launches_per_iter)}


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the same path with every host core, on the SAME workload as
    the GPU arm (one bench step = H consecutive Env.step calls for all agents of the job).  (The reference
    itself is Python over ROS/Gazebo and cannot travel to this box; the oracle port is its restatement,
    pinned to it by tests/golden.)"""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_total = args.agents * max(1, args.gpus)
    port = CpuPort(n_total, threads)
    H = args.horizon
    for _ in range(max(args.warmup, 1)):
        port.run(H)
    secs = 0.0
    for _ in range(args.steps):
        secs += port.run(H)
    value = n_total * H * args.steps / secs
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args.agents, H), "agents_per_gpu": args.agents, "horizon": H,
                   "agents_total": n_total},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} bench steps of {H} Env.step calls over {n_total} agents "
                                   f"({n_total * H * args.steps} env-steps, {secs:.2f} s), C port "
                                   f"(oracle/navsim_oracle.c), {threads} pooled pthreads, no Python between steps"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=8192, help="agents per GPU (BASELINE configs[1])")
    ap.add_argument("--horizon", type=int, default=128, help="env steps per agent per bench step")
    ap.add_argument("--ref-seconds", type=float, default=2.0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the PPO rollout / training-iteration figures")
    ap.add_argument("--train-iters", type=int, default=2)
    ap.add_argument("--epochs", type=int, default=50, help="PPO epochs per iteration (main.py:471)")
    ap.add_argument("--precision", default="fp32,bf16x3",
                    help="comma list of GEMM arithmetic modes for the training figures: fp32 (CUDA cores, the reference's "
                         "arithmetic), bf16x3 (tcgen05, split-bf16 operands, ~16 mantissa bits), bf16 (tcgen05)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from navbot_ppo_b200.env import VecEnv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N, H, K, W = args.agents, args.horizon, args.steps, max(args.warmup, 3)
    peak_gbs, peak_src = load_peaks()
    env = VecEnv(N, map="stage_1", device=local, seed=0, max_episode_steps=500, agent_id_offset=rank * N)
    env.reset()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    bufs = env.rollout_scripted(H, action_seed=0)     # allocates the [H, N, .] rollout buffers

    def one_step():
        env.rollout_scripted(H, action_seed=0, out=bufs)

    for _ in range(W):
        one_step()
    barrier()
    launches0 = env.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    mark0 = sampler.mark()
    for k in range(K):
        flush.fill_(k & 0xFF)          # evict L2 between timed steps (untimed)
        ev[k][0].record()
        one_step()
        ev[k][1].record()
    barrier()
    clocks = sampler.stop(mark0, sampler.mark()) if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = env.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    env_steps_total = world * N * H * K
    value = env_steps_total / (ms_total * 1e-3)

    # ---- the same H steps as H single-step launches (what a per-step caller pays) ----------
    a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.step_scripted(1, 0)
    barrier()
    a_ev.record()
    for _ in range(H):
        env.step_scripted(1, 0)
    b_ev.record()
    torch.cuda.synchronize()
    single_launch_us = a_ev.elapsed_time(b_ev) * 1e3 / H
    # ---- and with the actions in a device tensor, one Env.step launch per step: what a policy-driven caller
    # ---- (PPO.rollout) pays for the simulator
    act_dev = torch.rand((N, 2), device=dev)
    env.step(act_dev)
    barrier()
    a_ev.record()
    for _ in range(H):
        env.step(act_dev)
    b_ev.record()
    torch.cuda.synchronize()
    policy_step_us = a_ev.elapsed_time(b_ev) * 1e3 / H
    t = torch.tensor([policy_step_us], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value_policy_step = world * N / (float(t.item()) * 1e-6)

    # ---- end to end through the host-buffer entry point (navsim_step_host) ----------------
    He = H
    hb = env.alloc_host_buffers()                       # page-locked numpy arrays owned by the caller
    hb["act"][:] = np.random.RandomState(rank).uniform(0, 1, size=(N, 2)).astype(np.float32)
    for _ in range(3):
        env.step_host(hb["act"], out=hb)
    barrier()
    e2e_steps = He * max(1, min(K, 8))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        env.step_host(hb["act"], out=hb)                # blocks until obs/reward/flags are in host memory
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * N * e2e_steps / float(t.item())
    # the asynchronous form: two buffer sets, step t + 1 is enqueued before step t's results are awaited (every step
    # still takes its actions from host memory and delivers its results to host memory)
    depth = 4                                            # NAVSIM_ASYNC_DEPTH
    sets = [hb] + [env.alloc_host_buffers() for _ in range(depth - 1)]
    for sb in sets[1:]:
        sb["act"][:] = hb["act"]

    def pipelined(nsteps):
        for i in range(nsteps):
            sb = sets[i % depth]
            # one library call per env step: enqueue step i, return once step i - 3's obs / reward / flags are in host
            # memory (in the buffer set the next call overwrites)
            env.step_host_pipelined(sb["act"], sb)
        env.wait(0)

    pipelined(8)
    barrier()
    t0 = time.perf_counter()
    pipelined(e2e_steps)
    e2e_async_dt = time.perf_counter() - t0
    t = torch.tensor([e2e_async_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_async_value = world * N * e2e_steps / float(t.item())
    # the same call with ordinary (pageable) numpy arrays, as a drop-in caller would pass them
    acts_pageable = np.array(hb["act"])
    for _ in range(3):
        env.step_host(acts_pageable)
    t0 = time.perf_counter()
    for _ in range(He):
        env.step_host(acts_pageable)
    e2e_pageable = N * He / (time.perf_counter() - t0)

    # ---- PPO on top of the simulator: rollout (policy forward + sampling + env step) and the
    # ---- full training iteration (rollout + reward-to-go + 50-epoch update), SURVEY.md 8(d)
    training = None
    if not args.no_train:
        training = {}
        for prec_name in args.precision.split(","):
            args.precision_one = prec_name.strip()
            training[args.precision_one] = training_bench(args, torch, dist, dev, rank, world, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the step kernel: live CUDA-event duration per launch -----------------
    per_launch_s = (ms_total * 1e-3) / K
    fused_bytes = (64 + 4 + 3) + (BYTES_PER_ENV_STEP_SCRIPTED - 71) / H   # outputs every step, state once per launch
    def by_def(nbytes):
        a_ = nbytes * N * H / per_launch_s / 1e9
        return {"bytes_per_env_step": nbytes, "achieved": a_, "frac": a_ / peak_gbs}
    achieved = BYTES_SURVEY * N * H / per_launch_s / 1e9        # the contract figure: SURVEY.md 8(d), 150 B / env-step
    roofline = {"kernel": "navsim_step_kernel", "bound": "hbm", "achieved": achieved, "peak": peak_gbs,
                "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH.get((N, H)),
                "definition": "SURVEY.md 8(d): 150 B per env-step (44 B read + 106 B written, fp32 SoA)",
                "by_definition": {"survey_150B": by_def(BYTES_SURVEY), "survey_fp64_pose_174B": by_def(BYTES_SURVEY_FP64_POSE),
                                  "kernel_state_203B": by_def(BYTES_PER_ENV_STEP_SCRIPTED)},
                "peak_source": peak_src, "bytes_per_env_step": BYTES_PER_ENV_STEP_SCRIPTED,
                "env_steps_per_launch": N * H, "us_per_launch": per_launch_s * 1e6, "us_per_env_step_batch": per_launch_s * 1e6 / H,
                "lanes_per_agent": env.lanes_per_agent,
                "fused_launch_bytes_per_env_step": fused_bytes,
                "traffic_note": "DRAM bytes of one launch under ncu (cold caches): below even the 72 B/step of outputs because "
                                "most of the 75 MB of rollout rows is still in the 126 MB L2 when the kernel ends",
                "note": "one launch = H steps with the agent state in registers: per env-step it really moves "
                        f"{fused_bytes:.1f} B (outputs + state/H) instead of the per-step-launch figure {BYTES_PER_ENV_STEP_SCRIPTED} B "
                        "that `achieved` is defined on; N=8192 agents = 55 per SM, so the launch is bound by one "
                        "agent-step's dependent-instruction latency, not by HBM; see roofline_sweep for the large-batch figure",
                "single_step_launch_us": single_launch_us,
                "single_step_launch_env_steps_per_s": N / (single_launch_us * 1e-6)}
    sweep = []
    if not args.no_sweep and world == 1:
        for n_big, h_big in ((65536, 64), (1 << 20, 16), (1 << 22, 8)):
            e2 = VecEnv(n_big, map="stage_1", device=local, seed=0)
            e2.reset()
            ob = e2.rollout_scripted(h_big, 0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            a.record()
            for _ in range(reps):
                e2.rollout_scripted(h_big, 0, out=ob)
            b.record()
            torch.cuda.synchronize()
            s = a.elapsed_time(b) * 1e-3 / (reps * h_big)
            gbs = BYTES_PER_ENV_STEP_SCRIPTED * n_big / s / 1e9
            e2.step_scripted(1, 0)
            a.record()
            for _ in range(10):
                e2.step_scripted(1, 0)
            b.record()
            torch.cuda.synchronize()
            s1 = a.elapsed_time(b) * 1e-3 / 10
            sweep.append({"agents": n_big, "steps_per_launch": h_big, "lanes_per_agent": e2.lanes_per_agent,
                          "us_per_env_step_batch": s * 1e6, "env_steps_per_s": n_big / s,
                          "achieved": gbs, "frac": gbs / peak_gbs,
                          "single_step_launch": {"us_per_launch": s1 * 1e6, "achieved": BYTES_PER_ENV_STEP_SCRIPTED * n_big / s1 / 1e9,
                                                 "frac": BYTES_PER_ENV_STEP_SCRIPTED * n_big / s1 / 1e9 / peak_gbs}})
            e2.close()
            del e2, ob

    # ---- the other BASELINE.json configurations (parity-tested in tests/): configs[2] per GPU, and configs[4] /
    # ---- SURVEY.md 8(d) c5 as written: house map, 4096 agents per GPU (32,768 over 8), start / goal from the
    # ---- GoalSpawnSampler tables, beam sweep B in {10, 12, 18, 24, 36}; run on every rank, max over ranks
    other = []
    c5 = []
    if not args.no_sweep:
        sweep_cfgs = [("configs[2] stage_2 16384 agents/GPU", "stage_2", 16384, 10, False)] if world == 1 else []
        sweep_cfgs += [(f"configs[4] house 4096 agents/GPU {b} beams, table start/goal", "house", 4096, b, True)
                       for b in (10, 12, 18, 24, 36)]
        for label, mp, n_c, beams, tables in sweep_cfgs:
            e3 = VecEnv(n_c, map=mp, device=local, seed=0, num_beams=beams, agent_id_offset=rank * n_c,
                        use_external_sampler=("small_house" if tables else False))
            e3.reset()
            ob = e3.rollout_scripted(H, 0)
            for _ in range(3):                      # let the robots spread out
                e3.rollout_scripted(H, 0, out=ob)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                e3.rollout_scripted(H, 0, out=ob)
            b.record()
            torch.cuda.synchronize()
            t3 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            s3 = float(t3.item()) * 1e-3 / (3 * H)
            nseg = int(e3.segments.shape[0])
            row = {"config": label, "agents_per_gpu": n_c, "agents_total": n_c * world, "beams": beams, "walls": nseg,
                   "lanes_per_agent": e3.lanes_per_agent, "us_per_env_step_batch": s3 * 1e6,
                   "env_steps_per_s": world * n_c / s3, "nominal_ray_wall_tests_per_s": world * n_c * beams * nseg / s3}
            if tables:
                gbs = BYTES_SURVEY * n_c / s3 / 1e9                       # per GPU
                row["hbm"] = {"bytes_per_env_step": BYTES_SURVEY, "achieved_GBps_per_gpu": gbs, "frac": gbs / peak_gbs}
                fl = NCU_FP32_FLOP_PER_ENV_STEP.get(("house", beams))
                row["fp32"] = None if fl is None else {
                    "flop_per_env_step": fl, "achieved_TFLOPs_per_gpu": fl * n_c / s3 / 1e12,
                    "peak_TFLOPs": FP32_PEAK_TFLOPS, "frac": fl * n_c / s3 / 1e12 / FP32_PEAK_TFLOPS,
                    "source": "thread-level FADD + FMUL + 2 FFMA counts of one launch under ncu (profiles/README.md), "
                              "divided by its env-steps"}
                c5.append(row)
            else:
                other.append(row)
            e3.close()
            del e3, ob

    # CPU baseline: rank 0 at N = 1 only (the contract); multi-GPU lines carry null
    threads = os.cpu_count() or 1
    cpu_baseline = None
    if world == 1:
        cpu_v, cpu_steps, cpu_dt = cpu_port_throughput(N, args.cpu_seconds, threads)
        cpu1_v, _, _ = cpu_port_throughput(N, min(3.0, args.cpu_seconds), 1)
        cpu_baseline = {"value": cpu_v, "unit": "env-steps/s", "cores": threads, "kind": "port",
                        "single_core_value": cpu1_v,
                        "sample": f"{cpu_dt:.1f} s ({cpu_steps} Env.step batches over {N} agents) of the C port "
                                  f"oracle/navsim_oracle.c on {threads} pthreads"}

    line = {
        "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(N, H), "agents_per_gpu": N, "horizon": H, "launches_per_bench_step": 1,
                   "l2": "flushed between timed steps (256 MiB write, untimed)",
                   "parallelism": f"agents sharded x{world}, no data-path collective"},
        "clocks": clocks,
        "value_policy_step": {"value": value_policy_step, "unit": "env-steps/s", "us_per_launch": policy_step_us,
                              "note": "actions from a device tensor, ONE navsim_step launch per env step (no fusion over "
                                      "steps): the simulator's share of a policy-driven rollout"},
        "e2e": {"value": e2e_async_value, "unit": "env-steps/s", "h2d_bytes_per_step": N * 8 * H,
                "d2h_bytes_per_step": N * (64 + 4 + 3) * H,
                "api": "navsim_step_host_pipelined (VecEnv.step_host_pipelined = navsim_step_host_async + navsim_wait of a steady "
                       "pipeline in one call), the host entry point for a caller that keeps several steps in flight: EVERY "
                       "env step takes its actions from page-locked host "
                       "memory and delivers obs / reward / flags to page-locked host memory (four buffer sets, 4 steps in "
                       "flight; the actions are staged by a host-to-device copy under the previous step's kernel, the step's "
                       "results go home in one copy-engine transfer under the next step's kernel); the timed loop waits "
                       "for every step's results",
                "blocking_value": e2e_value,
                "blocking_api": "navsim_step_host via VecEnv.step_host: the same buffers, one step at a time (launch, "
                                "zero-copy reads / writes over PCIe, stream synchronisation per step)",
                "pageable_buffers_value": e2e_pageable,
                "d2h_link_note": "tools/pcie_d2h.py on the same box class: 42.6 GB/s for one step's 581,632 B, i.e. at most "
                                 "6.0e8 env-steps/s for 8192 robots (tools/time_async.py has the host-side split)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_sweep": sweep,
        "other_configs": other,
        "c5_house_beam_sweep": c5,
        "training": training,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
