#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library: counts of the Blackwell mnemonics (UTCHMMA = tcgen05.mma, UTMALDG /
UTMASTG / UTMAREDG = bulk-tensor load / store / reduce, LDTM = tcgen05.ld, UBLKCP = bulk copy, FFMA2 / HFMA2 packed
math, HMMA = legacy mma.sync) in every kernel of libirb200.so.  Runs on the build host (cuobjdump only, no GPU).

    python scripts/sass_summary.py > profiles/r02_sass_summary.json
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "image_restoration_models_b200", "libirb200.so")
MNEMONICS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "HFMA2", "HMMA",
             "FFMA", "MUFU", "LDG", "STG", "LDS", "STS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    kernels, cur, arch = collections.OrderedDict(), None, set()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_instructions"] += 1
            for mn in MNEMONICS:
                if op == mn or (mn in ("FFMA", "LDG", "STG", "LDS", "STS", "MUFU", "SYNCS") and op.startswith(mn) and
                                not (mn == "FFMA" and op.startswith("FFMA2"))):
                    kernels[cur][mn] += 1
    rows = []
    for name, c in kernels.items():
        full = demangle(name).replace("(anonymous namespace)::", "").replace("irb::", "").replace("void ", "", 1)
        short = full.split("(", 1)[0]
        rows.append({"kernel": short[:110], "instructions": c["_instructions"],
                     **{mn: c[mn] for mn in MNEMONICS if c[mn]}})
    total = collections.Counter()
    for r in rows:
        for k, v in r.items():
            if k not in ("kernel",):
                total[k] += v
    out = {"library": os.path.relpath(LIB, ROOT), "arch": sorted(arch), "n_kernels": len(rows), "totals": dict(total),
           "tcgen05_kernels": [r["kernel"] for r in rows if r.get("UTCHMMA")],
           "tma_kernels": [r["kernel"] for r in rows if r.get("UTMALDG") or r.get("UTMASTG") or r.get("UTMAREDG")],
           "kernels": rows}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
