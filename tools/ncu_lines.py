#!/usr/bin/env python
"""Per-source-line shares of an ncu capture:  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > lines.csv
   python tools/ncu_lines.py lines.csv [top]   -> warp instructions, stall samples and shared-memory wavefronts per file and per line."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
per_line = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < 10 or r[0] in ("", "Function Name"):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def num(k):
        try:
            return int(r[hdr[k]])
        except (ValueError, KeyError, IndexError):
            return 0
    per_line.append((cur, line, r[1].strip()[:110], num("Instructions Executed"), num("# Samples"), num("L1 Wavefronts Shared"), num("Thread Instructions Executed")))
tot_i = sum(x[3] for x in per_line) or 1
tot_s = sum(x[4] for x in per_line) or 1
tot_w = sum(x[5] for x in per_line) or 1
by_file = collections.defaultdict(lambda: [0, 0, 0])
for f, l, s, i, sm, w, t in per_line:
    by_file[f][0] += i; by_file[f][1] += sm; by_file[f][2] += w
print(f"total warp instructions {tot_i:.4g}, samples {tot_s}, shared wavefronts {tot_w:.4g}")
for f, (i, sm, w) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:18s} inst {i / tot_i:6.3f}  samples {sm / tot_s:6.3f}  smem wavefronts {w / tot_w:6.3f}")
print("top lines by stall samples:")
for f, l, s, i, sm, w, t in sorted(per_line, key=lambda x: -x[4])[:top]:
    print(f"  {f}:{l:<4d} inst {i / tot_i:6.3f} samples {sm / tot_s:6.3f} lanes {t / max(i, 1):5.1f} | {s}")
