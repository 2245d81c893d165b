"""Cross-check the warp-specialised tcgen05 gradient kernel (navppo_tcws.cu) against the first,
single-role kernel (navppo_tc.cu, NAVPPO_TC_KERNEL=single): bit-identical gradients / metrics,
and the time per gradient pass of each.

    python tools/tc_ws_check.py [T] [reps]
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from navbot_ppo_b200 import _capi, layout  # noqa: E402

DEV = "cuda:0"


def handle(prec, max_samples, single):
    if single:
        os.environ["NAVPPO_TC_KERNEL"] = "single"
    else:
        os.environ.pop("NAVPPO_TC_KERNEL", None)
    cfg = _capi.default_ppo_cfg()
    cfg.device = 0
    cfg.max_samples = int(max_samples)
    cfg.precision = int(prec)
    h = ctypes.c_void_p()
    _capi.check(_capi.lib().navppo_create(ctypes.byref(h), ctypes.byref(cfg)))
    os.environ.pop("NAVPPO_TC_KERNEL", None)
    return h


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 33333
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ppo_learn_b.npz"))
    rng = np.random.RandomState(0)
    idx = rng.randint(0, len(g["obs"]), T)
    obs = (g["obs"][idx] + rng.normal(scale=0.01, size=(T, 16))).astype(np.float32)
    act, lp, rtg = g["acts"][idx], (g["logp"][idx] + rng.normal(scale=0.1, size=T)).astype(np.float32), g["rtgs"][idx]
    adv = rng.normal(size=T).astype(np.float32)
    f = np.zeros(_capi.PPO_FLAT, np.float32)
    f[:layout.ACTOR_PARAMS] = g["actor_after"]
    f[_capi.PPO_CRITIC_OFFSET:_capi.PPO_CRITIC_OFFSET + layout.CRITIC_PARAMS] = g["critic_after"]
    flat = torch.from_numpy(f).to(DEV)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)  # noqa: E731
    o, a_, l_, ad, rt = t(obs), t(act), t(lp), t(adv), t(rtg)
    L = _capi.lib()
    ok = True
    for pname, prec in (("bf16x3", _capi.PREC_BF16X3), ("bf16", _capi.PREC_BF16)):
        out = {}
        for kname, single in (("single-role", True), ("warp-specialised", False)):
            h = handle(prec, max(T, 1024), single)
            grad = torch.zeros(_capi.PPO_FLAT, device=DEV)
            met = torch.zeros(8, dtype=torch.float64, device=DEV)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for rep in range(reps + 1):
                if rep == 1:
                    ev[0].record()
                rc = L.navppo_grad(h, flat.data_ptr(), o.data_ptr(), a_.data_ptr(), l_.data_ptr(), ad.data_ptr(),
                                   rt.data_ptr(), T, T, float(g["var"]), grad.data_ptr(), met.data_ptr(), None)
                assert rc == 0, L.nav_last_error()
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / reps
            out[kname] = (grad.clone(), met.clone(), ms)
            print(f"{pname:7s} {kname:17s} T={T} ms/grad={ms:.3f} |grad|max={float(grad.abs().max()):.4e} "
                  f"metrics={met.cpu().numpy()[:4]}", flush=True)
            L.navppo_destroy(h)
        same_g = torch.equal(out["single-role"][0], out["warp-specialised"][0])
        same_m = torch.equal(out["single-role"][1], out["warp-specialised"][1])
        d = (out["single-role"][0] - out["warp-specialised"][0]).abs().max().item()
        print(f"{pname:7s} gradients bit-identical: {same_g} (max abs diff {d:.3e}); metrics identical: {same_m}; "
              f"speed-up {out['single-role'][2] / out['warp-specialised'][2]:.2f}x", flush=True)
        ok = ok and same_g and same_m
    print("OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
