#!/bin/bash
# Round-2 evidence pass (final code): tensor peaks, full GPU suite, smoke, benches (fp32 with every baseline leg, half, the
# reference arm), ncu launch list + DRAM traffic, ncu --set full of the two fused kernels, batch-1 latencies, SASS summary.
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity.json $OUT/status_$TAG.txt
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt; tail -4 $OUT/pytest_$TAG.log
cp $OUT/parity.json $OUT/parity_$TAG.json 2>/dev/null
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" | tee -a $OUT/status_$TAG.txt; tail -2 $OUT/smoke_$TAG.log
IRB_PROFILE_DUMP=$OUT/launch_fp32_$TAG.csv timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; head -c 300 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
timeout 900 python bench.py --steps 10 --warmup 3 --mode half --no-cpu-baseline --no-eager > $OUT/bench_half_$TAG.json 2> $OUT/bench_half_$TAG.err
echo "bench half exit $?" | tee -a $OUT/status_$TAG.txt; head -c 200 $OUT/bench_half_$TAG.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
echo "bench reference exit $?" | tee -a $OUT/status_$TAG.txt; head -c 300 $OUT/bench_ref_$TAG.json; echo
bash scripts/ncu_traffic.sh $TAG fp32 | tail -2
timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_$TAG.log 2>&1
echo "blocks exit $?" | tee -a $OUT/status_$TAG.txt
timeout 600 python scripts/bench_latency.py > $OUT/latency_$TAG.json 2> $OUT/latency_$TAG.err
echo "latency exit $?" | tee -a $OUT/status_$TAG.txt
NCU_K="ffn_fused|attn_fused" NCU_CS=96 NCU_MODES=0 bash scripts/ncu_blocks.sh $TAG > $OUT/ncu_blocks_$TAG.log 2>&1
python scripts/summarize_ncu_full.py $OUT/ncu_raw_$TAG.csv $OUT/ncu_summary_$TAG.csv | tee -a $OUT/status_$TAG.txt
cp $OUT/ncu_raw_$TAG.csv $OUT/ncu_raw_keep_$TAG.csv 2>/dev/null
timeout 900 python scripts/bench_configs.py --steps 5 > $OUT/configs_$TAG.log 2>&1
echo "configs exit $?" | tee -a $OUT/status_$TAG.txt; cp $OUT/configs.json $OUT/configs_$TAG.json 2>/dev/null
cat $OUT/status_$TAG.txt
