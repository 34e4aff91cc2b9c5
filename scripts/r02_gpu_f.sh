#!/bin/bash
# Round-2 GPU pass F: the TMEM-direct fused MDTA front (attn_fused.cu v2): parity tests, block-level timings against v1.
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity.json
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_metrics.py -m gpu -q -x --timeout 600 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee $OUT/status_$TAG.txt; tail -15 $OUT/pytest_$TAG.log
cp $OUT/parity.json $OUT/parity_$TAG.json 2>/dev/null
timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_$TAG.log 2>&1
echo "blocks exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_$TAG.log | cut -c1-600
IRB_ATTN_FUSED_V1=1 timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_${TAG}_v1.log 2>&1
echo "blocks v1 exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_${TAG}_v1.log | cut -c1-600
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-tiled > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; head -c 400 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
cat $OUT/status_$TAG.txt
