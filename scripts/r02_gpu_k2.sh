#!/bin/bash
# Short evidence refresh (bench in both modes + block timings) after a small change; the ncu records stay those of the last full pass
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/status_$TAG.txt
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
IRB_PROFILE_DUMP=$OUT/launch_fp32_$TAG.csv timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; head -c 300 $OUT/bench_$TAG.json; echo
timeout 300 python bench.py --steps 10 --warmup 3 --mode half --no-cpu-baseline --no-eager > $OUT/bench_half_$TAG.json 2> $OUT/bench_half_$TAG.err
echo "bench half exit $?" | tee -a $OUT/status_$TAG.txt
timeout 200 python scripts/bench_kernels.py --blocks --levels > $OUT/blocks_$TAG.log 2>&1
echo "blocks exit $?" | tee -a $OUT/status_$TAG.txt
