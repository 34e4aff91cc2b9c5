#!/usr/bin/env python
"""Per-kernel stall breakdown from an `ncu --page source --csv --print-source sass` export (gzip ok).
Sums the per-instruction warp-stall samples by stall reason, overall and for the hottest instructions,
optionally restricted to the warps of one role by an instruction-address window.
Usage: ncu_sass_stalls.py sass.csv[.gz] [kernel-substring] [top-n]"""
import csv, gzip, sys, collections
fn = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
op = gzip.open if fn.endswith(".gz") else open
rows = list(csv.reader(op(fn, "rt")))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        i = j
        if want not in name:
            continue
        ix = {h: k for k, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = collections.Counter()
        samples = 0
        for r in body:
            for c in stall_cols:
                v = int(r[ix[c]] or 0)
                tot[c] += v
            samples += int(r[ix["# Samples"]] or 0)
        print("==", name[:100], "instructions", len(body), "samples", samples)
        for c, v in tot.most_common(12):
            print(f"   {c:28s} {v:8d} {100.0 * v / max(1, samples):5.1f} %")
        hot = sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:topn]
        for r in hot:
            reasons = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
            print(f"   {int(r[ix['# Samples']]):6d}  {body.index(r):5d} {r[ix['Source']].strip()[:70]:70s} " +
                  " ".join(f"{c}={v}" for v, c in reasons if v))
    else:
        i += 1
