#!/bin/bash
# A/B of one environment switch on the same box: block parity, (optionally) the full GPU suite, then block kernels and the
# bench with the switch off / on, twice.   Usage: bash scripts/r02_gpu_r.sh <tag> <ENV_SWITCH> [full]
TAG=${1:-r02r}; SW=${2:-IRB_OLD_RING_SPLIT}; FULL=${3:-}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/status_$TAG.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -k "block" -x -q --timeout 300 > $OUT/pytest_block_$TAG.log 2>&1
RC=$?; echo "block tests exit $RC" | tee -a $OUT/status_$TAG.txt; tail -5 $OUT/pytest_block_$TAG.log
if [ "$RC" != "0" ]; then exit 1; fi
if [ -n "$FULL" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > $OUT/pytest_$TAG.log 2>&1
  echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt; tail -6 $OUT/pytest_$TAG.log
  cp $OUT/parity.json $OUT/parity_$TAG.json 2>/dev/null
fi
for V in new old new2 old2; do
  if [[ $V == old* ]]; then export $SW=1; else unset $SW; fi
  timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_${TAG}_$V.log 2>&1
  echo "blocks $V exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_${TAG}_$V.log | cut -c1-420
  IRB_PROFILE_DUMP=$OUT/launch_${TAG}_$V.csv timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-tiled > $OUT/bench_${TAG}_$V.json 2> $OUT/bench_${TAG}_$V.err
  echo "bench $V exit $?" | tee -a $OUT/status_$TAG.txt; head -c 260 $OUT/bench_${TAG}_$V.json; echo; tail -2 $OUT/bench_${TAG}_$V.err
done
unset $SW
cat $OUT/status_$TAG.txt
