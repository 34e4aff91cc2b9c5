#!/bin/bash
# A/B timing of ffn_fused.cu variants (half-mode bench, per-launch CUDA-event times) after a block-parity check of each.
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
for V in "" "IRB_FUSED_FHFMA=1"; do
  N=${V:-default}
  env $V timeout 300 python -m pytest tests/test_gpu_parity.py -k "block or gray_64 or motion" -x -q --timeout 200 > $OUT/pytest_${TAG}_$N.log 2>&1
  echo "[$N] parity exit $?"; tail -3 $OUT/pytest_${TAG}_$N.log
  env $V IRB_PROFILE_DUMP=$OUT/launch_half_${TAG}_$N.csv timeout 300 python bench.py --steps 3 --warmup 3 --mode half --no-cpu-baseline > $OUT/bench_half_${TAG}_$N.json 2> $OUT/bench_half_${TAG}_$N.err
  echo "[$N] bench exit $?"
  python - <<PY
import csv, collections, json
rows = list(csv.reader(open("$OUT/launch_half_${TAG}_$N.csv")))
agg = collections.defaultdict(list)
for r in rows:
    if r[1] in ("gdfn_fused",): agg[(r[1], r[3])].append(float(r[2]))
for k, v in sorted(agg.items()): print("  ", k, len(v), "avg ms %.4f" % (sum(v) / len(v)))
d = json.load(open("$OUT/bench_half_${TAG}_$N.json")); print("   half Mpix/s", d["value"], "ms", d["ms_per_step"], "| fp32", d["other_mode"]["value"])
PY
done
