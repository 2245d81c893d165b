"""Check the tensor-core inference kernel (mlp_infer_ws_kernel, navppo_tcws.cu) through the C-ABI:
forward / act / evaluate of a NAVPPO_BF16X3 (and NAVPPO_BF16) handle against the fp32 CUDA-core kernel, ragged
sizes, and the time per call at the rollout's batch (8192 robots) and at the update's (1M samples).

    python tools/tc_infer_check.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from navbot_ppo_b200 import _capi, layout  # noqa: E402

DEV = "cuda:0"


def handle(prec, max_samples):
    cfg = _capi.default_ppo_cfg()
    cfg.device = 0
    cfg.max_samples = int(max_samples)
    cfg.precision = int(prec)
    h = ctypes.c_void_p()
    _capi.check(_capi.lib().navppo_create(ctypes.byref(h), ctypes.byref(cfg)))
    return h


def run(h, flat, obs, act_in, noise, var):
    L = _capi.lib()
    T = obs.shape[0]
    mu = torch.zeros(T, 2, device=DEV); v = torch.zeros(T, device=DEV)
    act = torch.zeros(T, 2, device=DEV); lp = torch.zeros(T, device=DEV); mu2 = torch.zeros(T, 2, device=DEV)
    v2 = torch.zeros(T, device=DEV); lp2 = torch.zeros(T, device=DEV)
    _capi.check(L.navppo_forward(h, flat.data_ptr(), obs.data_ptr(), T, mu.data_ptr(), v.data_ptr(), None))
    _capi.check(L.navppo_act(h, flat.data_ptr(), obs.data_ptr(), T, var, 7, 0, 3, noise.data_ptr(), act.data_ptr(), lp.data_ptr(),
                             mu2.data_ptr(), None))
    _capi.check(L.navppo_evaluate(h, flat.data_ptr(), obs.data_ptr(), act_in.data_ptr(), T, var, v2.data_ptr(), lp2.data_ptr(), None))
    torch.cuda.synchronize()
    return dict(mu=mu, v=v, act=act, lp=lp, mu2=mu2, v2=v2, lp2=lp2)


def timed(h, flat, obs, reps=20):
    L = _capi.lib()
    T = obs.shape[0]
    act = torch.zeros(T, 2, device=DEV); lp = torch.zeros(T, device=DEV)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for r in range(reps + 3):
        if r == 3:
            ev[0].record()
        _capi.check(L.navppo_act(h, flat.data_ptr(), obs.data_ptr(), T, 0.1, 7, 0, r, None, act.data_ptr(), lp.data_ptr(), None, None))
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps * 1e3


def main():
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ppo_learn_b.npz"))
    f = np.zeros(_capi.PPO_FLAT, np.float32)
    f[:layout.ACTOR_PARAMS] = g["actor_after"]
    f[_capi.PPO_CRITIC_OFFSET:_capi.PPO_CRITIC_OFFSET + layout.CRITIC_PARAMS] = g["critic_after"]
    flat = torch.from_numpy(f).to(DEV)
    var = float(g["var"])
    hs = {"fp32": handle(_capi.PREC_FP32, 1 << 20), "bf16x3": handle(_capi.PREC_BF16X3, 1 << 20), "bf16": handle(_capi.PREC_BF16, 1 << 20)}
    ok = True
    for T in (1, 127, 128, 129, 5000, 8192):
        rng = np.random.RandomState(T)
        idx = rng.randint(0, len(g["obs"]), T)
        obs = torch.from_numpy((g["obs"][idx] + rng.normal(scale=0.01, size=(T, 16))).astype(np.float32)).to(DEV)
        act_in = torch.from_numpy(np.ascontiguousarray(g["acts"][idx])).to(DEV)
        noise = torch.from_numpy(rng.normal(size=(T, 2)).astype(np.float32)).to(DEV)
        ref = run(hs["fp32"], flat, obs, act_in, noise, var)
        for name, tol in (("bf16x3", 2e-5), ("bf16", 2e-2)):
            out = run(hs[name], flat, obs, act_in, noise, var)
            worst = {k: float((out[k] - ref[k]).abs().max() / (1.0 + ref[k].abs().max())) for k in ref}
            good = all(w <= tol for w in worst.values()) and torch.equal(out["mu"], out["mu2"]) and torch.equal(out["v"], out["v2"])
            ok = ok and good
            print(f"T={T:5d} {name:7s} max rel err vs fp32: " + " ".join(f"{k}={w:.1e}" for k, w in worst.items()) + f"  {'ok' if good else 'BAD'}",
                  flush=True)
    for T in (8192, 1 << 20):
        obs = torch.randn(T, 16, device=DEV)
        for name in ("fp32", "bf16x3", "bf16"):
            print(f"navppo_act T={T}: {name:7s} {timed(hs[name], flat, obs, 20 if T < 100000 else 3):9.1f} us", flush=True)
    print("OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
