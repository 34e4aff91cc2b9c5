#!/bin/bash
# Timing-only experiments on ffn_fused.cu (IRB_FUSED_DBG bits: 1 no W_in reloads, 2 no xn reloads, 4 no taps, 8 no W_out reloads).
# The debug instantiations are compiled only with -DIRB_FUSED_EXPERIMENTS (add it to NVCC_FLAGS in build.py for the session).
TAG=${1:-dbg}
OUT=gpurun_out
mkdir -p $OUT
for D in ${DBG_LIST:-0 1 3 11 4 5 15}; do
  IRB_FUSED_DBG=$D IRB_PROFILE_DUMP=$OUT/launch_half_${TAG}_dbg$D.csv timeout 300 python bench.py --steps 2 --warmup 3 --mode half --no-cpu-baseline > $OUT/bench_half_${TAG}_dbg$D.json 2> $OUT/bench_half_${TAG}_dbg$D.err
  echo "dbg=$D exit $?"
  python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/launch_half_${TAG}_dbg$D.csv")))
agg = collections.defaultdict(list)
for r in rows:
    if r[1] in ("gdfn_fused",): agg[(r[1], r[3])].append(float(r[2]))
for k, v in sorted(agg.items()): print("  ", k, len(v), "avg ms %.4f" % (sum(v) / len(v)))
PY
done
