#!/usr/bin/env python
"""bench.py — headline benchmark: env-steps/s of the stage_1 hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" of this benchmark is one pass of the hot path over one batch: H consecutive
Env.step calls for every one of the 8192 agents of a rank (BASELINE.json configs[1];
agents shard over ranks with no data-path collective -> weak scaling).  Rank 0 prints ONE
JSON line.  See DESIGN.md "Measurement" for every figure's definition.
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


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_throughput(n_agents: int, seconds: float, threads: int):
    """The C oracle (oracle/navsim_oracle.c) stepping the same stage_1 workload on host cores."""
    from navbot_ppo_b200 import _capi, maps
    from oracle import binding
    cfg = _capi.default_cfg(n_agents)
    cfg.seed = 0
    sim = binding.OracleSim(cfg, maps.get_map("stage_1"), nthreads=threads)
    sim.reset()
    acts = [binding.scripted_actions(0, 0, t, n_agents) for t in range(16)]
    for t in range(4):
        sim.step(acts[t])
    t0 = time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < seconds:
        sim.step(acts[steps % 16])
        steps += 1
    dt = time.perf_counter() - t0
    return n_agents * steps / dt, steps, dt


def run_reference(args, rank, world):
    """--impl reference: the CPU port of the same path with every host core.  (The
    reference itself is Python over ROS/Gazebo and cannot travel to this box; the oracle
    port is its restatement, pinned to it by tests/golden.)"""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.agents
    per_step = []
    tot_steps = 0
    for i in range(args.warmup + args.steps):
        v, nsteps, dt = cpu_port_throughput(n, args.ref_seconds, threads)
        if i >= args.warmup:
            per_step.append((v, nsteps, dt))
            tot_steps += nsteps
    env_steps = sum(n * s for _, s, _ in per_step)
    secs = sum(d for _, _, d in per_step)
    value = env_steps / secs
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"stage_1 map, {n} agents, 10-beam LiDAR, 16-D obs / 2-D action, scripted actions",
                   "agents": n},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{args.ref_seconds:.0f} s of Env.step per bench step over {n} agents, "
                                   f"C port (oracle/navsim_oracle.c), {threads} pthreads"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=8192, help="agents per GPU (BASELINE configs[1])")
    ap.add_argument("--horizon", type=int, default=128, help="env steps per agent per bench step")
    ap.add_argument("--ref-seconds", type=float, default=2.0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-sweep", action="store_true")
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

    def one_step():
        env.step_scripted(H, action_seed=0)

    for _ in range(W):
        one_step()
    barrier()
    launches0 = env.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for k in range(K):
        flush.fill_(k & 0xFF)          # evict L2 between timed steps (untimed)
        ev[k][0].record()
        one_step()
        ev[k][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = env.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    env_steps_total = world * N * H * K
    value = env_steps_total / (ms_total * 1e-3)

    # ---- end to end through the host-buffer entry point (navsim_step_host) ----------------
    He = min(H, 32)
    acts = np.random.RandomState(rank).uniform(0, 1, size=(N, 2)).astype(np.float32)
    for _ in range(3):
        env.step_host(acts)
    barrier()
    t0 = time.perf_counter()
    for _ in range(He * max(1, K // 4)):
        env.step_host(acts)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * N * He * max(1, K // 4) / float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the step kernel: live CUDA-event duration per launch -----------------
    per_launch_s = (ms_total * 1e-3) / (K * H)
    achieved = BYTES_PER_ENV_STEP_SCRIPTED * N / per_launch_s / 1e9
    roofline = {"kernel": "navsim_step_kernel", "bound": "hbm", "achieved": achieved, "peak": peak_gbs,
                "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": None, "peak_source": peak_src,
                "bytes_per_env_step": BYTES_PER_ENV_STEP_SCRIPTED, "us_per_launch": per_launch_s * 1e6,
                "note": "at N=8192 one launch moves 1.7 MB: launch-latency bound; see roofline_sweep"}
    sweep = []
    if not args.no_sweep:
        for n_big in (65536, 1 << 20, 1 << 22):
            e2 = VecEnv(n_big, map="stage_1", device=local, seed=0)
            e2.reset()
            e2.step_scripted(5, 0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            a.record(); e2.step_scripted(reps, 0); b.record()
            torch.cuda.synchronize()
            s = a.elapsed_time(b) * 1e-3 / reps
            gbs = BYTES_PER_ENV_STEP_SCRIPTED * n_big / s / 1e9
            sweep.append({"agents": n_big, "us_per_launch": s * 1e6, "env_steps_per_s": n_big / s, "achieved": gbs,
                          "frac": gbs / peak_gbs})
            e2.close()
            del e2

    threads = os.cpu_count() or 1
    cpu_v, cpu_steps, cpu_dt = cpu_port_throughput(N, args.cpu_seconds, threads)
    cpu1_v, _, _ = cpu_port_throughput(N, min(3.0, args.cpu_seconds), 1)

    line = {
        "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"stage_1 map, {N} agents/GPU, 10-beam LiDAR, 16-D obs / 2-D action, "
                               f"{H} env steps per bench step, scripted Philox actions, auto-reset episodes (cap 500)",
                   "agents_per_gpu": N, "horizon": H, "l2": "flushed between timed steps (256 MiB write, untimed)",
                   "parallelism": f"agents sharded x{world}, no data-path collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": N * 8 * H,
                "d2h_bytes_per_step": N * (64 + 4 + 3) * H,
                "api": "navsim_step_host via VecEnv.step_host: host actions in, host obs/reward/flags out, every env step"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_sweep": sweep,
        "cpu_baseline": {"value": cpu_v, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "single_core_value": cpu1_v,
                         "sample": f"{cpu_dt:.1f} s ({cpu_steps} Env.step batches over {N} agents) of the C port "
                                   f"oracle/navsim_oracle.c on {threads} pthreads"},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
