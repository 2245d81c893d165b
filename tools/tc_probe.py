"""Which shared-memory word does the tensor core fetch for logical element (n, k) of an
MN-major B operand?  A = K-major one-hot rows (known-good path), B = raw image whose word w
holds a code; D[m][n] then reads back the code of B_hw(n, k = m % 8)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from navbot_ppo_b200 import _capi  # noqa: E402


def probe(N, lbo, sbo, b_major, label, operand="B"):
    K = 8
    img_words = N * K
    outs = []
    for code in ("granule", "lane"):
        img = np.arange(img_words)
        img = (img // 4 if code == "granule" else img % 4).astype(np.float32)
        onehot = np.zeros((128, K), np.float32)
        onehot[np.arange(128), np.arange(128) % 8] = 1.0
        if operand == "B":
            A, B = onehot, img
            a_mode, b_mode = 0, 2
            st = [2048, 128, 4096, lbo, sbo, 128, 0, b_major]
        a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        d = torch.full((128, N), -1.0, device="cuda")
        arr = (ctypes.c_uint32 * 8)(*st)
        rc = _capi.lib().navppo_tc_selftest(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, a_mode, b_mode, arr, None)
        torch.cuda.synchronize()
        outs.append(d.cpu().numpy()[:8].astype(int))     # rows m = 0..7 -> k = m
    g, l = outs
    word = g * 4 + l
    print(f"--- {label}: N={N} lbo={lbo} sbo={sbo} b_major={b_major}: word index fetched for (k rows, n cols 0..11)")
    for k in range(8):
        print(f"k={k}: " + " ".join(f"{w:4d}" for w in word[k][:12]) + " ... " + " ".join(f"{w:4d}" for w in word[k][-4:]))


if __name__ == "__main__":
    which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    cfgs = [(32, 128, 8 * 16, 1, "theory MN-major (sbo = K*16, lbo = 128)"),
            (32, 8 * 16, 128, 1, "swapped"),
            (32, 128, 128, 1, "both 128"),
            (32, 512, 128, 1, "lbo 512 sbo 128"),
            (32, 32 * 16, 128, 0, "K-major reference (lbo = N*16, sbo = 128)")]
    N, lbo, sbo, bm, label = cfgs[which]
    probe(N, lbo, sbo, bm, label)
