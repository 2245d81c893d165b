import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from navbot_ppo_b200.env import Env
env = Env(True, seed=0)
obs = env.reset()
past = np.zeros(2, np.float32)
rng = np.random.RandomState(0)
acts = np.stack([rng.uniform(0, 1, 2000), rng.uniform(-1, 1, 2000)], 1).astype(np.float32)
for t in range(200):
    obs, r, d, a = env.step(acts[t], past); past = acts[t]
    if d or a: env.reset(); past = np.zeros(2, np.float32)
t0 = time.perf_counter()
n = 0
for t in range(200, 2000):
    obs, r, d, a = env.step(acts[t], past); past = acts[t]; n += 1
    if d or a: env.reset(); past = np.zeros(2, np.float32)
dt = time.perf_counter() - t0
print(f"Env.step (one robot, drop-in class): {1e6 * dt / n:.1f} us/step = {n / dt:.0f} steps/s")
