"""One short vectorised PPO iteration (target for ncu captures of the trainer kernels).
usage: profile_ppo.py [agents] [horizon] [epochs] [fp32|bf16x3|bf16]"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from navbot_ppo_b200 import _capi  # noqa: E402
from navbot_ppo_b200.env import VecEnv  # noqa: E402
from navbot_ppo_b200.nets import NetActor, NetCritic  # noqa: E402
from navbot_ppo_b200.ppo import PPO  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
h = int(sys.argv[2]) if len(sys.argv) > 2 else 16
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
prec = {"fp32": _capi.PREC_FP32, "bf16x3": _capi.PREC_BF16X3, "bf16": _capi.PREC_BF16}[sys.argv[4] if len(sys.argv) > 4 else "bf16x3"]
env = VecEnv(n, map="stage_1", device=0, seed=0)
with tempfile.TemporaryDirectory() as tmp:
    agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=n * h, max_timesteps_per_episode=500,
                n_updates_per_iteration=epochs, seed=0, output_dir=tmp, method_name="prof", verbose=False, precision=prec)
    for _ in range(2):
        batch = agent.rollout([0, 0], 0)
        res = agent.update(*batch[:4])
    torch.cuda.synchronize()
    print("done", n, h, epochs, float(res["actor_losses"][-1]), float(res["critic_losses"][-1]))
