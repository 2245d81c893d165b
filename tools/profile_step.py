"""Run a few scripted simulator steps at a given batch size (target for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from navbot_ppo_b200.env import VecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
map_name = sys.argv[3] if len(sys.argv) > 3 else "stage_1"
env = VecEnv(n, map=map_name, device=0, seed=0)
env.reset()
env.step_scripted(steps, action_seed=0)
torch.cuda.synchronize()
print("done", n, steps, env.launch_count)
