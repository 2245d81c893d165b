"""Rollout A/B: the one-launch fused rollout kernel, the CUDA graph of policy + step kernels with and without
programmatic dependent launches, and the fp32 policy kernel — identical buffers where the arithmetic is the same, and
the time per rollout.

    python tools/rollout_check.py [agents] [horizon]
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(N, H):
    import hashlib
    import tempfile

    import torch

    from navbot_ppo_b200 import _capi
    from navbot_ppo_b200.env import VecEnv
    from navbot_ppo_b200.nets import NetActor, NetCritic
    from navbot_ppo_b200.ppo import PPO
    env = VecEnv(N, map="stage_1", device=0, seed=0, max_episode_steps=500)
    with tempfile.TemporaryDirectory() as tmp:
        agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=N * H, max_timesteps_per_episode=500,
                    n_updates_per_iteration=1, seed=0, output_dir=tmp, method_name="ab", verbose=False,
                    precision=_capi.PREC_BF16X3, log_episodes=False)
        batch = agent.rollout([0, 0], 0)
        torch.cuda.synchronize()
        hs = hashlib.sha256()
        for t in batch[:4]:
            hs.update(t.detach().cpu().numpy().tobytes())
        ms = []
        for _ in range(5):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            e[0].record()
            agent.rollout([0, 0], 0)
            e[1].record()
            torch.cuda.synchronize()
            ms.append(e[0].elapsed_time(e[1]))
        print(f"RESULT {hs.hexdigest()[:16]} {min(ms):.3f} {sorted(ms)[2]:.3f}", flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(int(sys.argv[2]), int(sys.argv[3]))
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    out = {}
    for name, env in (("fused", {}), ("tc+chained", {"NAVPPO_ROLLOUT_FUSED": "0"}),
                      ("tc", {"NAVPPO_ROLLOUT_FUSED": "0", "NAVPPO_ROLLOUT_CHAIN": "0"}), ("fp32 policy kernel", {"NAVPPO_TC_INFER": "0"})):
        r = subprocess.run([sys.executable, __file__, "--child", str(N), str(H)], env={**os.environ, **env}, capture_output=True,
                           text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        if not line:
            print(name, "FAILED", r.stdout[-2000:], r.stderr[-2000:])
            return 1
        _, h, best, med = line[0].split()
        out[name] = h
        print(f"{name:20s} rollout {N} x {H}: best {best} ms, median {med} ms  ({N * H / (float(med) * 1e-3):.3e} env-steps/s)  digest {h}",
              flush=True)
    same = out["tc+chained"] == out["tc"] == out["fused"]
    print("fused == chained == plain launches (bit-identical batch):", same)
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
