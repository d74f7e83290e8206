#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total ms, share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
tot, n = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    k = r[ki].split("(")[0].split("::")[-1][-32:]
    tot[k] = tot.get(k, 0) + float(r[vi].replace(",", ""))
    n[k] += 1
s = sum(tot.values())
print(f"{'kernel':34s} {'n':>4s} {'total ms':>10s} {'share':>7s}")
for k, v in tot.items():
    print(f"{k:34s} {n[k]:4d} {v / 1e6:10.2f} {v / s * 100:6.1f}%")
print(f"{'all':34s} {sum(n.values()):4d} {s / 1e6:10.2f}")
