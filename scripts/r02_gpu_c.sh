#!/bin/bash
# Round-2 GPU pass C: block-level kernel timings of the fused MDTA front (+ timing experiments) and one ncu --set full
# capture of it with per-instruction stall sampling.
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block or fresh or guard" > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee $OUT/status_$TAG.txt; tail -3 $OUT/pytest_$TAG.log
timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_$TAG.log 2>&1
echo "blocks exit $?" | tee -a $OUT/status_$TAG.txt; cat $OUT/blocks_$TAG.log | cut -c1-900
for D in ${DBG_LIST:-1 2 8 11}; do
  IRB200_LIB=$PWD/build_ab/libirb200_dbg.so IRB_AF_DBG=$D timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_${TAG}_dbg$D.log 2>&1
  echo "dbg $D: $(grep C96 $OUT/blocks_${TAG}_dbg$D.log | grep fp32 | python -c 'import sys,json; [print(json.loads(l)["kernels"].get("mdta_fused_front")) for l in sys.stdin]')" | tee -a $OUT/status_$TAG.txt
done
NCU_K=attn_fused NCU_CS=96 NCU_MODES=0 bash scripts/ncu_blocks.sh $TAG > $OUT/ncu_blocks_$TAG.log 2>&1
tail -5 $OUT/ncu_blocks_$TAG.log
cat $OUT/status_$TAG.txt
