"""Aggregate an ncu source page (--page source --csv --print-source cuda,sass) by CUDA-C line:
instructions executed and stall samples per source line.  usage: ncu_source_lines.py rep [top]"""
import csv
import io
import subprocess
import sys
import collections

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# locate header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ci = {n: i for i, n in enumerate(hdr)}
# first "Source" col = CUDA-C text, second = SASS
src_cols = [i for i, n in enumerate(hdr) if n == "Source"]
agg = collections.OrderedDict()
files = {}
cur_file = ""
tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        if r and r[0] == "File Path":
            cur_file = r[1]
        continue
    try:
        ln = r[ci["Line No"]]
        inst = int(r[ci["Instructions Executed"]] or 0)
        samp = int(r[ci["# Samples"]] or 0)
    except ValueError:
        continue
    key = (ln, r[src_cols[0]].strip()[:100])
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += inst; a[1] += samp; a[2] += 1
    tot_i += inst; tot_s += samp
print(f"total instructions executed {tot_i}, samples {tot_s}")
for (ln, text), (inst, samp, nsass) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * inst / max(tot_i, 1):5.1f}% inst {100 * samp / max(tot_s, 1):5.1f}% stall  sass={nsass:4d}  L{ln:>5s}  {text}")
