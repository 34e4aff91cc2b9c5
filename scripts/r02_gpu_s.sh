#!/bin/bash
# Round-2 GPU pass S: where a depthwise warp of the fused MDTA front spends its time (experiments build, IRB_AF_DBG=64:
# clock64 around the accumulator waits / the units / the X-tile wait), the re-parallelised fold kernel, block parity.
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/status_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "${PYTEST_K:-block or fresh or guard or golden}" > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt; tail -5 $OUT/pytest_$TAG.log
timeout 300 python scripts/bench_kernels.py --blocks --levels > $OUT/blocks_$TAG.log 2>&1
echo "blocks exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_$TAG.log | cut -c1-700
IRB_FOLD_PG1=1 timeout 300 python scripts/bench_kernels.py --blocks --levels > $OUT/blocks_${TAG}_pg1.log 2>&1
echo "blocks pg1 exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_${TAG}_pg1.log | cut -c1-700
IRB200_LIB=$PWD/build_ab/libirb200_dbg.so IRB_AF_DBG=64 timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_${TAG}_dbg64.log 2>&1
echo "dbg64 exit $?" | tee -a $OUT/status_$TAG.txt
grep "af-dbg" $OUT/blocks_${TAG}_dbg64.log | tail -8 | tee -a $OUT/status_$TAG.txt
