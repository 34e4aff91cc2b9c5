#!/bin/bash
# Round-2 evidence pass E (after the container re-creation): dense tensor peaks first (bench.py reads them), the full GPU
# test suite, smoke, the bench with the per-launch dump, half-mode bench, ncu launch list + DRAM traffic of one forward,
# block-level kernel timings and the batch-1 latencies.
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 300 python scripts/measure_tensor_peaks.py > $OUT/tensor_peaks_$TAG.log 2>&1
echo "tensor peaks exit $?" | tee $OUT/status_$TAG.txt; tail -1 $OUT/tensor_peaks_$TAG.log | cut -c1-600
[ -s $OUT/tensor_peaks.json ] && cp $OUT/tensor_peaks.json profiles/tensor_peaks.json
rm -f $OUT/parity.json
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt; tail -8 $OUT/pytest_$TAG.log
cp $OUT/parity.json $OUT/parity_$TAG.json 2>/dev/null
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" | tee -a $OUT/status_$TAG.txt; tail -2 $OUT/smoke_$TAG.log
IRB_PROFILE_DUMP=$OUT/launch_fp32_$TAG.csv timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; head -c 600 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
timeout 900 python bench.py --steps 10 --warmup 3 --mode half --no-cpu-baseline --no-eager > $OUT/bench_half_$TAG.json 2> $OUT/bench_half_$TAG.err
echo "bench half exit $?" | tee -a $OUT/status_$TAG.txt; head -c 300 $OUT/bench_half_$TAG.json; echo
bash scripts/ncu_traffic.sh $TAG fp32 | tail -2
timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_$TAG.log 2>&1
echo "blocks exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_$TAG.log | cut -c1-600
timeout 600 python scripts/bench_latency.py > $OUT/latency_$TAG.json 2> $OUT/latency_$TAG.err
echo "latency exit $?" | tee -a $OUT/status_$TAG.txt
cat $OUT/status_$TAG.txt
