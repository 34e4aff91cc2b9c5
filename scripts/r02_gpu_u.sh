#!/bin/bash
# Round-2 GPU pass U: full GPU test suite + block timings (+ optional short bench)
TAG=${1:-r02u}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/status_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 ${PYTEST_K:+-k "$PYTEST_K"} > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt; tail -4 $OUT/pytest_$TAG.log
timeout 300 python scripts/bench_kernels.py --blocks --levels > $OUT/blocks_$TAG.log 2>&1
echo "blocks exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_$TAG.log | cut -c1-900
if [ "${BENCH:-0}" = "1" ]; then
  IRB_PROFILE_DUMP=$OUT/launch_fp32_$TAG.csv timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
  echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; head -c 400 $OUT/bench_$TAG.json; echo
fi
