"""Device -> pinned-host copy rate at the size one host-buffer env step moves (8192 robots: 581,632 B), the ceiling of
bench.py's e2e.async_pipelined_value:  python tools/pcie_d2h.py"""
import torch

for nbytes in (581632, 8 << 20, 256 << 20):
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    reps = max(5, min(2000, (2 << 30) // nbytes))
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    s = e0.elapsed_time(e1) * 1e-3 / reps
    print(f"D2H {nbytes:>10d} B: {s * 1e6:8.1f} us per copy, {nbytes / s / 1e9:6.1f} GB/s"
          + (f"  -> at most {8192 / s:.3e} env-steps/s for 8192 robots" if nbytes == 581632 else ""))
