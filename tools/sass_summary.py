"""Per-kernel SASS mnemonic counts of the in-tree library (cuobjdump -sass): the evidence that the tensor-core
kernels are tcgen05 / TMEM / TMA code and that the multicast Adam variant reads through multimem (LDGMC).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "navbot_ppo_b200", "libnavbot_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "ELECT", "LDGMC", "FADD2", "FMUL2", "FFMA2",
         "HMMA", "FFMA", "RED", "ATOM", "LDG", "STG", "LDS", "STS", "SHFL", "ACQBULK", "USETMAXREG", "STL", "LDL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB)}: arch {', '.join(arch)}; instruction counts per kernel (static)")
    print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA),")
    print("# SYNCS = mbarrier ops, LDGMC = multimem.ld_reduce (NVLS), FADD2 / FMUL2 = packed fp32x2")
    for name, c in zip(demangle, kernels.values()):
        short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
        short = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", short)
        cols = " ".join(f"{k}={c[k]}" for k in WATCH if c[k])
        print(f"{short[:70]:70s} total={c['_total']:6d}  {cols}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
