"""Latency of the host-buffer step call (navsim_step_host) vs batch size, page-locked and pageable."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from navbot_ppo_b200.env import VecEnv  # noqa: E402

for n in (8, 1024, 8192, 65536):
    env = VecEnv(n, seed=0)
    env.reset()
    hb = env.alloc_host_buffers()
    hb["act"][:] = np.random.RandomState(0).uniform(0, 1, size=(n, 2)).astype(np.float32)
    pageable = np.array(hb["act"])
    for _ in range(20):
        env.step_host(hb["act"], out=hb)
    t0 = time.perf_counter()
    for _ in range(500):
        env.step_host(hb["act"], out=hb)
    t1 = time.perf_counter()
    for _ in range(20):
        env.step_host(pageable)
    t2 = time.perf_counter()
    for _ in range(200):
        env.step_host(pageable)
    t3 = time.perf_counter()
    act_d = torch.from_numpy(pageable).cuda()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    for _ in range(500):
        env.step(act_d)
        torch.cuda.synchronize()
    t5 = time.perf_counter()
    print(f"N={n:6d}  step_host pinned {1e6 * (t1 - t0) / 500:7.1f} us   pageable {1e6 * (t3 - t2) / 200:7.1f} us   "
          f"device step + sync {1e6 * (t5 - t4) / 500:7.1f} us", flush=True)
    env.close()
