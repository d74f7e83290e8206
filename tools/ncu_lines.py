#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample totals of one kernel from an .ncu-rep
(ncu --set full --import-source on).  Usage: ncu_lines.py report.ncu-rep kernel [launch-skip] [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# regex: so that template instances (k_cns_dp<8, 128>) and mangled forms match by their base name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-skip", skip,
                      "--launch-count", "1", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next((i for i, r in enumerate(rows) if r and r[0] == "Line No"), None)
if hi is None:
    sys.exit(f"no source page for kernel '{kern}' (launch-skip {skip}) in {rep}: " + out[:300].replace("\n", " "))
h = rows[hi]
ci, cs = h.index("Instructions Executed"), h.index("# Samples")
agg = {}
fname = ""
for r in rows[:hi]:
    if r and r[0] == "File Path":
        fname = r[1]
for r in rows[hi + 1:]:
    if len(r) < len(h):
        if r and r[0] == "File Path":
            fname = r[1]
        continue
    if r[2] != "-":          # SASS row
        continue
    try:
        key = (fname.split("/")[-1], int(r[0]), r[1].strip()[:90])
        agg[key] = (agg.get(key, (0, 0))[0] + int(r[ci]), agg.get(key, (0, 0))[1] + int(r[cs]))
    except ValueError:
        pass
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {ti}, samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]}:{k[1]:4d} inst {100 * v[0] / ti:5.1f}%  samples {100 * v[1] / ts:5.1f}%  {k[2]}")
