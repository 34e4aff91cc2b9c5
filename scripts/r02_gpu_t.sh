#!/bin/bash
# Round-2 GPU pass T: experiments build only (IRB_AF_DBG bits: 64 clock64 instrumentation of the fused front, 128 no proxy fence)
TAG=${1:-r02t}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/status_$TAG.txt
for D in ${DBG_LIST:-0 64 128 192}; do
IRB200_LIB=$PWD/build_ab/libirb200_dbg.so IRB_AF_DBG=$D timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_${TAG}_dbg$D.log 2>&1
echo "dbg$D exit $?" | tee -a $OUT/status_$TAG.txt
grep "af-dbg" $OUT/blocks_${TAG}_dbg$D.log | tail -16 | sort | tee -a $OUT/status_$TAG.txt
grep fp32 $OUT/blocks_${TAG}_dbg$D.log | grep C96 | cut -c1-400 | tee -a $OUT/status_$TAG.txt
done
