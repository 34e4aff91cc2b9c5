#!/bin/bash
# Round-2 GPU pass I: both TMEM-direct fused kernels with producer warps: parity, block timings, timing experiments, bench.
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity.json $OUT/status_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "${PYTEST_K:-block or fresh or guard or golden}" > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt; tail -5 $OUT/pytest_$TAG.log
timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_$TAG.log 2>&1
echo "blocks exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_$TAG.log | cut -c1-700
for D in ${DBG_LIST:-2 4 16 22 246}; do
  IRB200_LIB=$PWD/build_ab/libirb200_dbg.so IRB_FUSED_DBG=$D timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_${TAG}_dbg$D.log 2>&1
  echo "gdfn dbg $D: $(grep C96 $OUT/blocks_${TAG}_dbg$D.log | grep fp32 | python -c 'import sys,json; [print(json.loads(l)["kernels"].get("gdfn_fused")) for l in sys.stdin]')" | tee -a $OUT/status_$TAG.txt
done
for D in ${AF_DBG_LIST:-2 48}; do
  IRB200_LIB=$PWD/build_ab/libirb200_dbg.so IRB_AF_DBG=$D timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_${TAG}_afdbg$D.log 2>&1
  echo "front dbg $D: $(grep C96 $OUT/blocks_${TAG}_afdbg$D.log | grep fp32 | python -c 'import sys,json; [print(json.loads(l)["kernels"].get("mdta_fused_front")) for l in sys.stdin]')" | tee -a $OUT/status_$TAG.txt
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-tiled > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; head -c 300 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
