"""Device/wall time of the PPO building blocks at rollout and update sizes."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from navbot_ppo_b200 import _capi  # noqa: E402
from navbot_ppo_b200.env import VecEnv  # noqa: E402
from navbot_ppo_b200.nets import NetActor, NetCritic, _stream  # noqa: E402
from navbot_ppo_b200.ppo import PPO  # noqa: E402

N, H = 8192, 128
prec = {"fp32": _capi.PREC_FP32, "bf16x3": _capi.PREC_BF16X3, "bf16": _capi.PREC_BF16}[sys.argv[1] if len(sys.argv) > 1 else "bf16x3"]
env = VecEnv(N, map="stage_1", device=0, seed=0)
tmp = tempfile.mkdtemp()
agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=N * H, max_timesteps_per_episode=500,
            n_updates_per_iteration=2, seed=0, output_dir=tmp, method_name="t", verbose=False, precision=prec)
L = _capi.lib()
dev = torch.device("cuda:0")
sp = _stream(dev)


def timeit(name, fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(reps):
        fn()
    b.record(); t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{name:44s} device {a.elapsed_time(b) * 1e3 / reps:10.1f} us   host-issue {t_issue * 1e6 / reps:8.1f} us", flush=True)


obs = env.reset()
act = torch.empty((N, 2), device=dev); logp = torch.empty(N, device=dev)
p_flat, p_obs, p_act, p_logp = agent.flat.data_ptr(), obs.data_ptr(), act.data_ptr(), logp.data_ptr()
timeit("navppo_act N=8192 (raw ctypes)", lambda: L.navppo_act(agent._h, p_flat, p_obs, N, 0.8, 0, 0, 1, None, p_act, p_logp, None, sp), 200)
timeit("env.step N=8192 (VecEnv.step)", lambda: env.step(act), 200)
batch = agent.rollout([0, 0], 0)
timeit("rollout H=128 (per env step)", lambda: agent.rollout([0, 0], 0), 3)
o, a_, lp, rtg = batch[:4]
T = o.shape[0]
v = torch.empty(T, device=dev); lg = torch.empty(T, device=dev)
timeit("navppo_evaluate T=1M", lambda: L.navppo_evaluate(agent._h, p_flat, o.data_ptr(), a_.data_ptr(), T, 0.8, v.data_ptr(), lg.data_ptr(), sp), 5)
adv = torch.randn(T, device=dev); grad = torch.zeros(_capi.PPO_FLAT, device=dev); met = torch.zeros(8, dtype=torch.float64, device=dev)
timeit("navppo_grad T=1M", lambda: L.navppo_grad(agent._h, p_flat, o.data_ptr(), a_.data_ptr(), lp.data_ptr(), adv.data_ptr(), rtg.data_ptr(), T, T, 0.8, grad.data_ptr(), met.data_ptr(), sp), 5)
timeit("update 2 epochs T=1M", lambda: agent.update(o, a_, lp, rtg), 2)
