#!/bin/bash
# One-GPU evidence pass: parity tests, smoke, both benches (+ the reference arm), ncu launch lists with DRAM traffic,
# one ncu --set full capture of the fused GDFN kernel and of the TMA contractions, per-config throughput.
# Usage (under gpurun): bash scripts/final_1gpu.sh <tag>
TAG=${1:-fin}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee $OUT/status_$TAG.txt; tail -4 $OUT/pytest_$TAG.log
cp $OUT/parity.json $OUT/parity_$TAG.json 2>/dev/null
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" | tee -a $OUT/status_$TAG.txt; tail -2 $OUT/smoke_$TAG.log
IRB_PROFILE_DUMP=$OUT/launch_fp32_$TAG.csv timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; head -c 700 $OUT/bench_$TAG.json; echo
IRB_PROFILE_DUMP=$OUT/launch_half_$TAG.csv timeout 900 python bench.py --steps 10 --warmup 3 --mode half --no-cpu-baseline > $OUT/bench_half_$TAG.json 2> $OUT/bench_half_$TAG.err
echo "bench half exit $?" | tee -a $OUT/status_$TAG.txt; head -c 300 $OUT/bench_half_$TAG.json; echo
IRB_NO_FFN_FUSED=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_nofuse_$TAG.json 2> $OUT/bench_nofuse_$TAG.err
echo "bench (two-kernel GDFN, A/B) exit $?" | tee -a $OUT/status_$TAG.txt; head -c 200 $OUT/bench_nofuse_$TAG.json; echo
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
echo "bench reference exit $?" | tee -a $OUT/status_$TAG.txt; head -c 400 $OUT/bench_ref_$TAG.json; echo
bash scripts/ncu_traffic.sh $TAG fp32 | tail -2
bash scripts/ncu_traffic.sh $TAG half | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ffn_fused" -s 18 -c 1 \
    -f -o $OUT/prof_fused_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_fused_$TAG.log 2>&1
echo "ncu fused exit $?" | tee -a $OUT/status_$TAG.txt
timeout 900 ncu --set full --clock-control none -k "regex:tma_gemm|attn_front|layernorm_rows" -s 60 -c 6 \
    -f -o $OUT/prof_blocks_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_blocks_$TAG.log 2>&1
echo "ncu blocks exit $?" | tee -a $OUT/status_$TAG.txt
timeout 900 python scripts/bench_configs.py --steps 5 > $OUT/configs_$TAG.log 2>&1
echo "configs exit $?" | tee -a $OUT/status_$TAG.txt; cp $OUT/configs.json $OUT/configs_$TAG.json 2>/dev/null; tail -12 $OUT/configs_$TAG.log | cut -c1-260
cat $OUT/status_$TAG.txt
