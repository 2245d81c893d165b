"""Per-role cycle accounting of the warp-specialised tcgen05 gradient kernel (navppo_tc_profile):
where CTA (0, 0)'s epilogue warps, MMA-issuing thread and flush warp spend their cycles.

    python tools/tc_ws_profile.py [T]
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from navbot_ppo_b200 import _capi, layout  # noqa: E402
from tools.tc_ws_check import handle  # noqa: E402

DEV = "cuda:0"
EPI = ["wait Z/GH in TMEM", "wait tile columns free", "epilogue forward", "epilogue backward", "wait pass complete",
       "row-owner work", "-", "-"]
MMA = ["wait X/GU published", "wait epilogue of half-chunk", "wait weights", "wait dW flushed", "commits / other",
       "issue Z / GH", "issue U / GX", "issue dW"]
FL = ["wait dW complete", "flush", "-", "-", "-", "-", "-", "-"]


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1 << 20
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ppo_learn_b.npz"))
    rng = np.random.RandomState(0)
    idx = rng.randint(0, len(g["obs"]), T)
    obs = (g["obs"][idx] + rng.normal(scale=0.01, size=(T, 16))).astype(np.float32)
    act, lp, rtg = g["acts"][idx], g["logp"][idx], g["rtgs"][idx]
    adv = rng.normal(size=T).astype(np.float32)
    f = np.zeros(_capi.PPO_FLAT, np.float32)
    f[:layout.ACTOR_PARAMS] = g["actor_after"]
    f[_capi.PPO_CRITIC_OFFSET:_capi.PPO_CRITIC_OFFSET + layout.CRITIC_PARAMS] = g["critic_after"]
    flat = torch.from_numpy(f).to(DEV)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)  # noqa: E731
    o, a_, l_, ad, rt = t(obs), t(act), t(lp), t(adv), t(rtg)
    L = _capi.lib()
    h = handle(_capi.PREC_BF16X3, max(T, 1024), False)
    grad = torch.zeros(_capi.PPO_FLAT, device=DEV)
    met = torch.zeros(8, dtype=torch.float64, device=DEV)
    prof = torch.zeros(32 + 4 * 512, dtype=torch.int64, device=DEV)

    def run():
        rc = L.navppo_grad(h, flat.data_ptr(), o.data_ptr(), a_.data_ptr(), l_.data_ptr(), ad.data_ptr(), rt.data_ptr(), T, T,
                           float(g["var"]), grad.data_ptr(), met.data_ptr(), None)
        assert rc == 0, L.nav_last_error()

    run()
    torch.cuda.synchronize()
    L.navppo_tc_profile(prof.data_ptr())
    run()
    torch.cuda.synchronize()
    L.navppo_tc_profile(None)
    pc = prof.cpu().numpy()
    tiles = (T + 127) // 128
    per_cta = -(-tiles // 74)
    print(f"T={T}: {tiles} tiles, CTA (0,0) walks {per_cta}; cycles per tile by role and category")
    for name, base, labels in (("epilogue warp 0 (row owner)", 0, EPI), ("epilogue warp 4", 8, EPI), ("consuming MMA warp (U, dW, GX)", 16, MMA),
                               ("flush warp 8", 24, FL)):
        tot = pc[base:base + 8].sum()
        print(f"  {name}: total {tot / per_cta:9.0f}")
        for i, lab in enumerate(labels):
            if lab != "-":
                print(f"      {lab:30s} {pc[base + i] / per_cta:9.0f}  ({100.0 * pc[base + i] / max(tot, 1):5.1f} %)")


    # event trace of one tile: every mark is "the segment that just ended was of this category"
    if "--trace" in sys.argv:
        names = {0: ("E0", EPI), 1: ("E4", EPI), 2: ("MMA", MMA), 3: ("FL", FL)}
        ev = []
        t0 = min(int(pc[32 + r * 512]) for r in range(4) if pc[32 + r * 512 + 511] > 0)
        for r in range(4):
            base = 32 + r * 512
            n = int(pc[base + 511])
            prev = int(pc[base]) & ((1 << 56) - 1)
            for k in range(1, n):
                v = int(pc[base + k])
                cat, t = (v >> 56) & 0xFF, v & ((1 << 56) - 1)
                ev.append((prev - t0, t - t0, names[r][0], names[r][1][cat]))
                prev = t
        ev.sort()
        print("trace of one tile: start end dur role segment")
        for a_, b_, r, c in ev:
            if b_ - a_ >= int(os.environ.get("TRACE_MIN", "150")):
                print(f"  {a_:7d} {b_:7d} {b_ - a_:6d}  {r:4s} {c}")


if __name__ == "__main__":
    main()
