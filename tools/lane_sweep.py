"""Time the step kernel for every lanes-per-agent split at a few batch sizes:
single-step launches and the fused H-step rollout launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from navbot_ppo_b200.env import VecEnv  # noqa: E402

H = 128
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [8192, 16384, 65536, 1 << 20]
map_name = sys.argv[2] if len(sys.argv) > 2 else "stage_1"
beams = int(sys.argv[3]) if len(sys.argv) > 3 else 10
for n in sizes:
    for lanes in (1, 2, 4, 8, 16, 32):
        if n * lanes > (1 << 22):
            continue
        env = VecEnv(n, map=map_name, device=0, seed=0, lanes_per_agent=lanes, num_beams=beams)
        env.reset()
        reps = max(1, min(8, (1 << 24) // (n * H)))
        out = env.rollout_scripted(H, 0)
        env.step_scripted(4, 0)
        torch.cuda.synchronize()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        for _ in range(reps * H):
            env.step_scripted(1, 0)
        b.record()
        for _ in range(reps):
            env.rollout_scripted(H, 0, out=out)
        c.record()
        torch.cuda.synchronize()
        t1 = a.elapsed_time(b) * 1e3 / (reps * H)
        t2 = b.elapsed_time(c) * 1e3 / (reps * H)
        print(f"{map_name} B={beams} N={n:8d} lanes={lanes:2d}  single-step launch {t1:8.2f} us/step ({n / t1 * 1e6:.3e} steps/s)   "
              f"fused {H}-step launch {t2:8.2f} us/step ({n / t2 * 1e6:.3e} steps/s)", flush=True)
        env.close()
        del env, out
