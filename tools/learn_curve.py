"""Short end-to-end training run on one GPU: does the policy learn?  Prints one line per iteration.
usage: learn_curve.py [iterations] [agents] [horizon] [epochs] [fp32|bf16x3|bf16] [map]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from navbot_ppo_b200 import _capi  # noqa: E402
from navbot_ppo_b200.env import VecEnv  # noqa: E402
from navbot_ppo_b200.nets import NetActor, NetCritic  # noqa: E402
from navbot_ppo_b200.ppo import PPO  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
h = int(sys.argv[3]) if len(sys.argv) > 3 else 128
epochs = int(sys.argv[4]) if len(sys.argv) > 4 else 50
prec = {"fp32": _capi.PREC_FP32, "bf16x3": _capi.PREC_BF16X3, "bf16": _capi.PREC_BF16}[sys.argv[5] if len(sys.argv) > 5 else "bf16x3"]
map_name = sys.argv[6] if len(sys.argv) > 6 else "stage_1"
cont = len(sys.argv) > 7 and sys.argv[7] == "continue"   # keep episodes across rollouts + bootstrap the tail
env = VecEnv(n, map=map_name, device=0, seed=0, max_episode_steps=500)
rows = []
with tempfile.TemporaryDirectory() as tmp:
    agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=n * h, max_timesteps_per_episode=500,
                n_updates_per_iteration=epochs, gamma=0.99, lr=3e-4, clip=0.2, seed=0, output_dir=tmp, method_name="curve",
                verbose=False, precision=prec, log_episodes=False, save_freq=10 ** 9, continue_episodes=cont,
                bootstrap_value=cont)
    t_so_far = 0
    t0 = time.time()
    for it in range(iters):
        obs, acts, logp, rtgs, lens, m, _ = agent.rollout([0, 0], t_so_far)
        t_so_far += int(lens.sum())
        res = agent.update(obs, acts, logp, rtgs)
        ne = max(1, m["ep_count"])
        row = dict(iteration=it + 1, env_steps=(it + 1) * n * h, episodes=m["ep_count"], success=m["successes"] / ne,
                   collision=m["collisions"] / ne, timeout=m["timeouts"] / ne, avg_return=m["return_sum"] / ne,
                   avg_len=m["length_sum"] / ne, actor_loss=float(res["actor_losses"][-1]), critic_loss=float(res["critic_losses"][-1]),
                   approx_kl=float(res["approx_kl"][-1]), var=agent.var, wall_s=time.time() - t0)
        rows.append(row)
        print(json.dumps(row), flush=True)
torch.cuda.synchronize()
print(f"total {time.time() - t0:.1f} s for {iters * n * h} env steps -> {iters * n * h / (time.time() - t0):.3e} env-steps/s wall")
