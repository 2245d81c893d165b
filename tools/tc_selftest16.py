"""One-CTA BF16 / split-BF16 GEMM self-test on the tcgen05 tensor cores (navppo_tc_selftest_bf16)
over every operand role of the fused update kernel.  One process per case."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [  # (a_mode, b_mode, N, K)
    (0, 0, 128, 32), (0, 0, 128, 16), (0, 0, 16, 128), (0, 0, 32, 128),      # Z = X Wa^T, U = H Wb^T
    (0, 1, 128, 32), (0, 1, 128, 16), (0, 1, 32, 128),                        # GH = GU Wb, GX = GZ Wa
    (1, 1, 48, 128), (1, 1, 32, 128), (1, 1, 16, 128),                        # dW = GZ^T X, H^T GU
]


def run_one(a_mode, b_mode, N, K, passes):
    import numpy as np
    import torch
    from navbot_ppo_b200 import _capi
    rng = np.random.RandomState(N * 1000 + K + a_mode * 7 + b_mode * 13)
    A = rng.normal(size=(128, K)).astype(np.float32)
    B = rng.normal(size=(N, K)).astype(np.float32)
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    d = torch.full((128, N), -7.0, device="cuda")
    rc = _capi.lib().navppo_tc_selftest_bf16(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, a_mode, b_mode, passes, None)
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    err = np.abs(d.cpu().numpy() - ref).max()
    tol = 0.2 if passes == 1 else 2e-3
    print(f"a_mode={a_mode} b_mode={b_mode} N={N:3d} K={K:3d} passes={passes} rc={rc} max_err={err:.3e} "
          f"ref_scale={np.abs(ref).max():.2f} {'OK' if err < tol else 'WRONG'}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_one(*[int(x) for x in sys.argv[1:6]])
    else:
        for case in CASES:
            for passes in (1, 3):
                try:
                    out = subprocess.run([sys.executable, __file__, *map(str, case), str(passes)], capture_output=True,
                                         text=True, timeout=120)
                    print(out.stdout.strip() or f"case {case}: no output; stderr tail: {out.stderr[-300:]}", flush=True)
                except subprocess.TimeoutExpired:
                    print(f"case {case} passes={passes}: TIMEOUT", flush=True)
