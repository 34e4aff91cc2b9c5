#!/bin/bash
# Round-2 GPU pass G: the TMEM-direct fused MDTA front behind a LayerNorm pass: parity, block timings, timing experiments
# (which role paces it).
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity.json $OUT/status_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "${PYTEST_K:-block or fresh or guard or golden}" > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt; tail -5 $OUT/pytest_$TAG.log
timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_$TAG.log 2>&1
echo "blocks exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_$TAG.log | cut -c1-700
for D in ${DBG_LIST:-2 8 16 32 48 56}; do
  IRB200_LIB=$PWD/build_ab/libirb200_dbg.so IRB_AF_DBG=$D timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_${TAG}_dbg$D.log 2>&1
  echo "dbg $D: $(grep C96 $OUT/blocks_${TAG}_dbg$D.log | grep fp32 | python -c 'import sys,json; [print(json.loads(l)["kernels"].get("mdta_fused_front")) for l in sys.stdin]')" | tee -a $OUT/status_$TAG.txt
done
