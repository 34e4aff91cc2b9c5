#!/bin/bash
# Round-2 GPU pass B: GPU tests with the fused MDTA front, benches with and without it (A/B), ncu launch list + traffic.
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity.json
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee $OUT/status_$TAG.txt; tail -8 $OUT/pytest_$TAG.log
cp $OUT/parity.json $OUT/parity_$TAG.json 2>/dev/null
IRB_PROFILE_DUMP=$OUT/launch_fp32_$TAG.csv timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; head -c 500 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
IRB_NO_ATTN_FUSED=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-tiled > $OUT/bench_noattnfused_$TAG.json 2> $OUT/bench_noattnfused_$TAG.err
echo "bench (two-kernel front, A/B) exit $?" | tee -a $OUT/status_$TAG.txt; head -c 300 $OUT/bench_noattnfused_$TAG.json; echo
timeout 900 python bench.py --steps 10 --warmup 3 --mode half --no-cpu-baseline --no-eager > $OUT/bench_half_$TAG.json 2> $OUT/bench_half_$TAG.err
echo "bench half exit $?" | tee -a $OUT/status_$TAG.txt; head -c 300 $OUT/bench_half_$TAG.json; echo
cat $OUT/status_$TAG.txt
