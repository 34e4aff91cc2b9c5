#!/bin/bash
# Timing experiments for ffn_fused.cu (half mode): debug switches + one ncu full capture of a full-resolution launch.
TAG=${1:-fx}
OUT=gpurun_out
mkdir -p $OUT
for D in ${DBG_LIST:-0 1 3 7 15}; do
  IRB_FUSED_DBG=$D IRB_PROFILE_DUMP=$OUT/launch_half_${TAG}_dbg$D.csv timeout 300 python bench.py --steps 2 --warmup 3 --mode half --no-cpu-baseline > $OUT/bench_half_${TAG}_dbg$D.json 2> $OUT/bench_half_${TAG}_dbg$D.err
  echo "dbg=$D exit $?"
  python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/launch_half_${TAG}_dbg$D.csv")))
agg = collections.defaultdict(list)
for r in rows:
    if r[1] in ("gdfn_fused", "layernorm"): agg[(r[1], r[3])].append(float(r[2]))
for k, v in sorted(agg.items()): print("  ", k, len(v), "avg ms %.4f" % (sum(v) / len(v)))
PY
done
if [ "${SKIP_NCU:-0}" = "0" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ffn_fused" -s ${NCU_SKIP:-18} -c 1 \
      -f -o $OUT/prof_fused_$TAG python bench.py --steps 1 --warmup 0 --mode half --no-cpu-baseline > $OUT/ncu_fused_$TAG.log 2>&1
  echo "ncu exit $?"; tail -3 $OUT/ncu_fused_$TAG.log
fi
