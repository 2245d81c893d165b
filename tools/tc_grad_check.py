"""Compare navppo_grad in the tensor-core modes with the fp32 CUDA-core mode and the float64
oracle on a golden batch; print error levels and timings."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from navbot_ppo_b200 import _capi, layout  # noqa: E402
from navbot_ppo_b200.nets import _Handles  # noqa: E402

DEV = "cuda:0"


def flat_from(actor, critic):
    f = np.zeros(_capi.PPO_FLAT, np.float32)
    f[:layout.ACTOR_PARAMS] = actor
    f[_capi.PPO_CRITIC_OFFSET:_capi.PPO_CRITIC_OFFSET + layout.CRITIC_PARAMS] = critic
    return torch.from_numpy(f).to(DEV)


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ppo_learn_b.npz"))
    rng = np.random.RandomState(0)
    idx = rng.randint(0, len(g["obs"]), T) if T != len(g["obs"]) else np.arange(T)
    obs = g["obs"][idx] + (rng.normal(scale=0.01, size=(T, 16)).astype(np.float32) if T != len(g["obs"]) else 0)
    act, lp, rtg = g["acts"][idx], g["logp"][idx], g["rtgs"][idx]
    adv = rng.normal(size=T).astype(np.float32)
    flat = flat_from(g["actor_after"], g["critic_after"])
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    o, a_, l_, ad, rt = t(obs.astype(np.float32)), t(act), t(lp), t(adv), t(rtg)
    L = _capi.lib()
    res = {}
    for name, prec in (("fp32", _capi.PREC_FP32), ("bf16x3", _capi.PREC_BF16X3), ("bf16", _capi.PREC_BF16)):
        h = _Handles.get(torch.device(DEV), max(T, 1024), 0.2, 3e-4, prec)
        grad = torch.zeros(_capi.PPO_FLAT, device=DEV)
        met = torch.zeros(8, dtype=torch.float64, device=DEV)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for rep in range(reps + 1):
            if rep == 1:
                ev[0].record()
            rc = L.navppo_grad(h, flat.data_ptr(), o.data_ptr(), a_.data_ptr(), l_.data_ptr(), ad.data_ptr(), rt.data_ptr(), T, T,
                               float(g["var"]), grad.data_ptr(), met.data_ptr(), None)
            assert rc == 0, L.nav_last_error()
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps if reps else float("nan")
        res[name] = (grad.cpu().numpy().astype(np.float64), met.cpu().numpy(), ms)
        print(f"{name:7s} T={T} ms/grad={ms:.3f} metrics={met.cpu().numpy()[:4]}", flush=True)
    ref = res["fp32"][0]
    scale_a = np.abs(ref[:layout.ACTOR_PARAMS]).max()
    scale_c = np.abs(ref[_capi.PPO_CRITIC_OFFSET:]).max()
    for name in ("bf16x3", "bf16"):
        d = np.abs(res[name][0] - ref)
        print(f"{name:7s} vs fp32: actor max err {d[:layout.ACTOR_PARAMS].max() / scale_a:.3e} (rel to max {scale_a:.3e}), "
              f"critic {d[_capi.PPO_CRITIC_OFFSET:].max() / scale_c:.3e} (rel to max {scale_c:.3e}); "
              f"metric diffs {np.abs(res[name][1][:4] - res['fp32'][1][:4])}", flush=True)
        # where are the largest errors?
        off = layout.offsets("actor")
        for k, (o_, shp) in off.items():
            n = int(np.prod(shp))
            print(f"    actor {k:16s} err {d[o_:o_ + n].max() / scale_a:.3e}   |ref| max {np.abs(ref[o_:o_ + n]).max():.3e}")


if __name__ == "__main__":
    main()
