"""Timeline (ns, %globaltimer) of one CTA of the tensor-core inference kernel, from the hooks navppo_tc_profile
switches on: where a rollout step's policy call spends its time at 8192 robots (one tile per CTA).

    python tools/tc_infer_profile.py [T]
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from navbot_ppo_b200 import _capi  # noqa: E402

DEV = "cuda:0"
NAMES = ["kernel entry", "prologue done (barriers, TMEM, biases)", "dependency wait passed", "x0 published",
         "pass 0: first Z in TMEM", "pass 0: epilogues done", "pass 0: U complete", "y1 published",
         "pass 1: first Z in TMEM", "pass 1: epilogues done", "pass 1: U complete", "outputs written", "TMEM freed"]


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    cfg = _capi.default_ppo_cfg()
    cfg.device = 0
    cfg.max_samples = 1 << 16
    cfg.precision = _capi.PREC_BF16X3
    h = ctypes.c_void_p()
    L = _capi.lib()
    _capi.check(L.navppo_create(ctypes.byref(h), ctypes.byref(cfg)))
    flat = (torch.randn(_capi.PPO_FLAT, device=DEV) * 0.05).contiguous()
    obs = torch.randn(T, 16, device=DEV)
    act = torch.zeros(T, 2, device=DEV); lp = torch.zeros(T, device=DEV)
    prof = torch.zeros(4096 + 64, dtype=torch.int64, device=DEV)
    for r in range(5):
        if r == 4:
            L.navppo_tc_profile(prof.data_ptr())
        _capi.check(L.navppo_act(h, flat.data_ptr(), obs.data_ptr(), T, 0.1, 7, 0, r, None, act.data_ptr(), lp.data_ptr(), None, None))
    torch.cuda.synchronize()
    L.navppo_tc_profile(None)
    t = prof[4096:4096 + len(NAMES)].cpu().numpy()
    print(f"tensor-core inference kernel, T={T}, CTA (0, 0), first row owner: ns since kernel entry")
    for k, name in enumerate(NAMES):
        print(f"  {int(t[k] - t[0]):7d}  (+{int(t[k] - t[max(k - 1, 0)]):6d})  {name}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
