#!/bin/bash
# Quick GPU-box pass after a kernel change: block parity, full parity suite, benches in both modes with per-launch
# CUDA-event dumps (+ optional A/B against the multi-kernel GDFN).  Usage (under gpurun): bash scripts/check_ffn_fused.sh <tag>
TAG=${1:-ff}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -k "block" -x -q --timeout 300 > $OUT/pytest_block_$TAG.log 2>&1
RC=$?; echo "block tests exit $RC"; tail -25 $OUT/pytest_block_$TAG.log
if [ "$RC" != "0" ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; tail -8 $OUT/pytest_$TAG.log
cp $OUT/parity.json $OUT/parity_$TAG.json 2>/dev/null
IRB_PROFILE_DUMP=$OUT/launch_fp32_$TAG.csv timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench fp32 exit $?"; tail -c 2800 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
IRB_PROFILE_DUMP=$OUT/launch_half_$TAG.csv timeout 600 python bench.py --steps 5 --warmup 3 --mode half --no-cpu-baseline > $OUT/bench_half_$TAG.json 2> $OUT/bench_half_$TAG.err
echo "bench half exit $?"; tail -c 2800 $OUT/bench_half_$TAG.json; tail -3 $OUT/bench_half_$TAG.err
if [ "${AB:-0}" = "1" ]; then
  IRB_NO_FFN_FUSED=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_nofuse.json 2> $OUT/bench_${TAG}_nofuse.err
  echo "bench fp32 (unfused) exit $?"; head -c 300 $OUT/bench_${TAG}_nofuse.json
fi
