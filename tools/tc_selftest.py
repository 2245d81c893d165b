"""Probe the tcgen05 operand-layout conventions with the one-CTA GEMM self-test
(navppo_tc_selftest).  Each variant runs in its own process: a wrong descriptor may fault.

    python tools/tc_selftest.py            # all cases, theory + swapped-stride variants
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [  # (a_mode, b_mode, N, K)
    (0, 0, 128, 32), (0, 0, 128, 16), (0, 0, 16, 128), (0, 0, 32, 128),
    (0, 1, 128, 32), (0, 1, 128, 16), (0, 1, 32, 128),
    (1, 1, 48, 128), (1, 1, 32, 128), (1, 1, 16, 128),
]


def strides(a_mode, b_mode, N, K, swap_a=False, swap_b=False):
    if a_mode == 0:
        a = [128 * 16, 128, 2 * 128 * 16]
    else:
        a = [128, K * 16, 128]
    if b_mode == 0:
        b = [N * 16, 128, 2 * N * 16]
    else:
        b = [128, K * 16, 128]
    if swap_a:
        a[0], a[1] = a[1], a[0]
    if swap_b:
        b[0], b[1] = b[1], b[0]
    return a + b


def run_one(a_mode, b_mode, N, K, swap_a, swap_b):
    import ctypes
    import numpy as np
    import torch
    from navbot_ppo_b200 import _capi
    rng = np.random.RandomState(N * 1000 + K + a_mode * 7 + b_mode * 13)
    A = rng.normal(size=(128, K)).astype(np.float32)
    B = rng.normal(size=(N, K)).astype(np.float32)
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    d = torch.zeros((128, N), device="cuda")
    st = (ctypes.c_uint32 * 8)(*strides(a_mode, b_mode, N, K, swap_a, swap_b), a_mode, b_mode)
    rc = _capi.lib().navppo_tc_selftest(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, a_mode, b_mode, st, None)
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    err = np.abs(d.cpu().numpy() - ref).max()
    print(f"a_mode={a_mode} b_mode={b_mode} N={N:3d} K={K:3d} swap_a={int(swap_a)} swap_b={int(swap_b)} rc={rc} "
          f"max_err={err:.4e} ref_scale={np.abs(ref).max():.2f} {'OK' if err < 0.05 else 'WRONG'}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_one(*[int(x) for x in sys.argv[1:7]])
    else:
        theory_only = os.environ.get("TC_THEORY_ONLY")
        for case in CASES:
            for sa in (0, 1):
                for sb in (0, 1):
                    if theory_only and (sa or sb):
                        continue
                    try:
                        out = subprocess.run([sys.executable, __file__, *map(str, case), str(sa), str(sb)],
                                             capture_output=True, text=True, timeout=120)
                        print(out.stdout.strip() or f"case {case} sa={sa} sb={sb}: no output; stderr tail: {out.stderr[-300:]}",
                              flush=True)
                    except subprocess.TimeoutExpired:
                        print(f"case {case} sa={sa} sb={sb}: TIMEOUT", flush=True)
