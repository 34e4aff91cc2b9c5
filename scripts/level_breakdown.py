#!/usr/bin/env python
"""Per-(kernel family, level) times from IRB_PROFILE_DUMP launch CSVs; two files -> side by side.
Usage: level_breakdown.py a.csv [b.csv]"""
import collections, csv, sys
def load(f):
    agg = collections.OrderedDict()
    for r in csv.reader(open(f)):
        k = (r[1], float(r[3]))
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[2])
    return agg
A = load(sys.argv[1]); B = load(sys.argv[2]) if len(sys.argv) > 2 else None
names = collections.defaultdict(float)
for (n, b), (c, ms) in A.items(): names[n] += ms
print("total A %.2f ms" % sum(v[1] for v in A.values()), ("total B %.2f ms" % sum(v[1] for v in B.values())) if B else "")
for (n, b), (c, ms) in sorted(A.items(), key=lambda x: (-names[x[0][0]], -x[0][1])):
    line = f"{n:26s} {b/1e6:9.1f} MB x{c:3d} {ms:7.3f} ms {b/1e9/(ms/c/1e3):7.0f} GB/s"
    if B:
        m = [(k, v) for k, v in B.items() if k[0] == n and abs(k[1] - b) < 1]
        if m: line += f"   | B {m[0][1][1]:7.3f} ms"
    print(line)
if B:
    for k, v in B.items():
        if not any(k2[0] == k[0] and abs(k2[1] - k[1]) < 1 for k2 in A): print("only in B:", k, v)
