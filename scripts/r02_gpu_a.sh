#!/bin/bash
# Round-2 first GPU pass: full GPU test suite (incl. 512x512 oracle parity + fp16 range tests), smoke, bench (with the
# tiled config-4 object and the eager-CUDA baseline), dense tensor peaks, compute-sanitizer.
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
rm -f $OUT/parity.json
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee $OUT/status_$TAG.txt; tail -6 $OUT/pytest_$TAG.log
cp $OUT/parity.json $OUT/parity_$TAG.json 2>/dev/null
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" | tee -a $OUT/status_$TAG.txt; tail -2 $OUT/smoke_$TAG.log
IRB_PROFILE_DUMP=$OUT/launch_fp32_$TAG.csv timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; head -c 600 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
timeout 300 python scripts/measure_tensor_peaks.py > $OUT/tensor_peaks_$TAG.log 2>&1
echo "tensor peaks exit $?" | tee -a $OUT/status_$TAG.txt; tail -1 $OUT/tensor_peaks_$TAG.log | cut -c1-600
SAN_TIMEOUT=360 bash scripts/sanitize.sh $TAG memcheck racecheck synccheck
cat $OUT/status_$TAG.txt
