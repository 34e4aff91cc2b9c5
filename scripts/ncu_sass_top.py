#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export (gzipped): per kernel, the SASS lines with the
most warp-stall samples and the dominant stall reason.  Usage: ncu_sass_top.py file.csv.gz [kernel_index] [n]"""
import csv, gzip, io, sys
csv.field_size_limit(10**9)
txt = gzip.open(sys.argv[1], "rt").read()
blocks = txt.split('"Kernel Name",')[1:]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
for k, b in enumerate(blocks):
    rd = list(csv.reader(io.StringIO(b)))
    name, hdr = rd[0][0], rd[1]
    rows = [r for r in rd[2:] if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    s = ix["Warp Stall Sampling (All Samples)"]
    tot = sum(int(r[s]) for r in rows)
    print(k, name[:100], "lines", len(rows), "samples", tot)
    if which is None or which != k:
        continue
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {c: sum(int(r[ix[c]]) for r in rows) for c in stall}
    print("  totals:", sorted(agg.items(), key=lambda x: -x[1])[:8])
    top = sorted(range(len(rows)), key=lambda i: -int(rows[i][s]))[:n]
    for i in sorted(top):
        r = rows[i]
        st = sorted(((int(r[ix[c]]), c[6:]) for c in stall), reverse=True)[:2]
        print(f"  {i:5d} {r[ix['Source']][:72]:72s} {r[s]:>6s} x{r[ix['Instructions Executed']]:>9s} {st}")
