"""Run a few scripted simulator launches at a given batch size (target for ncu captures).
usage: profile_step.py [agents] [steps] [map] [lanes] [fused 0|1] [beams] [sampler world_type | -]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from navbot_ppo_b200 import maps  # noqa: E402
from navbot_ppo_b200.env import VecEnv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
map_name = sys.argv[3] if len(sys.argv) > 3 else "stage_1"
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
fused = int(sys.argv[5]) if len(sys.argv) > 5 else 0
beams = int(sys.argv[6]) if len(sys.argv) > 6 else 10
m = maps.synthetic_map(int(map_name[5:]), seed=5) if map_name.startswith("synth") else map_name
sampler = sys.argv[7] if len(sys.argv) > 7 and sys.argv[7] != "-" else False
env = VecEnv(n, map=m, device=0, seed=0, lanes_per_agent=lanes, num_beams=beams, use_external_sampler=sampler)
env.reset()
if fused:
    out = env.rollout_scripted(steps, 0)
    for _ in range(2):                 # let the robots spread out before the captured launch
        env.rollout_scripted(steps, 0, out=out)
    out = env.rollout_scripted(steps, 0, out=out)
else:
    for _ in range(steps):
        env.step_scripted(1, action_seed=0)
torch.cuda.synchronize()
print("done", n, steps, env.launch_count)
