"""In-tree build of the CUDA library (sm_100a only) with plain nvcc.

    python -m navbot_ppo_b200.build [--force]

Produces navbot_ppo_b200/libnavbot_b200.so.  The simulator translation unit is compiled
with -fmad=false (bit-exact physics, see csrc/navsim_math.h); the PPO/MLP unit keeps FMA
contraction.  nvcc cross-compiles without a GPU, so this also runs on the CPU-only box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libnavbot_b200.so")
OBJ = os.path.join(PKG, "csrc", "_obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math",
          "-Xcompiler", "-ffp-contract=off", "-I", os.path.join(ROOT, "include")]

# (source, extra flags)
UNITS = [
    ("navsim_kernels.cu", ["-fmad=false"]),
    ("navppo_kernels.cu", []),
    # -fmad=false: navppo_tcws.cu steps the simulator inside its fused rollout kernel (navsim_device.cuh), whose fp64
    # chains must round exactly like navsim_kernels.cu's; navppo_tc.cu is its bit-exact cross-check and follows suit
    # (every fused multiply-add of the network arithmetic is an explicit fmaf)
    ("navppo_tc.cu", ["-fmad=false"]),
    ("navppo_tcws.cu", ["-fmad=false"]),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    objs, jobs = [], []
    for src, extra in UNITS:
        spath = os.path.join(CSRC, src)
        if not os.path.exists(spath):
            continue
        opath = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", spath, "-o", opath]
        stamp = opath + ".cmd"                       # a changed command line (flags) rebuilds too
        same_cmd = os.path.exists(stamp) and open(stamp).read() == " ".join(cmd)
        if force or not same_cmd or _stale(opath, [spath] + headers):
            jobs.append((cmd, stamp))
        objs.append(opath)

    def compile_unit(job):
        cmd, stamp = job
        run = [cmd[0], "-Xptxas=-v", *cmd[1:]] if verbose else cmd
        if verbose:
            print(" ".join(run), flush=True)
        subprocess.run(run, check=True)
        with open(stamp, "w") as f:
            f.write(" ".join(cmd))

    if jobs:                                         # the translation units compile side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
            list(pool.map(compile_unit, jobs))
    if force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "--cudart", "static", "-Xlinker", "--no-undefined",
               "-lpthread", "-ldl", "-lrt"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
