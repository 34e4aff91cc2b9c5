#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the ordered list."""
import collections
import csv
import sys

path = sys.argv[1]
verbose = len(sys.argv) > 2
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
seq = []
for row in rows:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("irb::", "").replace("(anonymous namespace)::", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    grid = row.get("Grid Size", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    seq.append((name, grid, v))
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.3f} ms over {len(seq)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={v[0]:4d} ms={v[1]:9.3f} share={v[1] / tot:.3f}")
if verbose:
    for i, (n, g, v) in enumerate(seq):
        print(i, n[:60], g, f"{v:.4f}")
