#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== tc_unit" | tee $OUT/status_$TAG.txt
timeout 1500 python scripts/tc_unit.py > $OUT/tc_unit_$TAG.log 2>&1
echo "tc_unit exit $?" | tee -a $OUT/status_$TAG.txt
tail -30 $OUT/tc_unit_$TAG.log
echo "== pytest -m gpu" | tee -a $OUT/status_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt
tail -15 $OUT/pytest_$TAG.log
echo "== smoke" | tee -a $OUT/status_$TAG.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" | tee -a $OUT/status_$TAG.txt
tail -5 $OUT/smoke_$TAG.log
echo "== bench" | tee -a $OUT/status_$TAG.txt
rm -f $OUT/launch_fp32_$TAG.csv $OUT/launch_half_$TAG.csv
IRB_PROFILE_DUMP=$OUT/launch_fp32_$TAG.csv timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
BENCH_RC=$?
echo "bench exit $BENCH_RC" | tee -a $OUT/status_$TAG.txt
tail -c 1500 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
IRB_PROFILE_DUMP=$OUT/launch_half_$TAG.csv timeout 900 python bench.py --steps 5 --warmup 3 --mode half --no-cpu-baseline > $OUT/bench_half_$TAG.json 2> $OUT/bench_half_$TAG.err
echo "bench half exit $?" | tee -a $OUT/status_$TAG.txt
tail -c 1500 $OUT/bench_half_$TAG.json; tail -5 $OUT/bench_half_$TAG.err
if [ "$BENCH_RC" = "0" ] && [ "${SKIP_NCU:-0}" = "0" ]; then
  echo "== ncu launch list" | tee -a $OUT/status_$TAG.txt
  timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/plain_$TAG.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:${NCU_LIST_REGEX:-gemm|dwconv|gram|fold|copy_channels|k1_|k4_|k6_|tc_}" -c ${NCU_LIST_COUNT:-380} --csv \
      --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_list_$TAG.log 2>&1
  echo "ncu list exit $?" | tee -a $OUT/status_$TAG.txt
  echo "== ncu full capture" | tee -a $OUT/status_$TAG.txt
  timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:${NCU_FULL_REGEX:-gemm_simt}" -s ${NCU_FULL_SKIP:-330} -c ${NCU_FULL_COUNT:-3} \
      -f -o $OUT/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
  echo "ncu full exit $?" | tee -a $OUT/status_$TAG.txt
  if [ -n "$NCU_FULL_REGEX2" ]; then
    timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$NCU_FULL_REGEX2" -s ${NCU_FULL_SKIP2:-0} -c ${NCU_FULL_COUNT2:-2} \
        -f -o $OUT/prof2_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full2_$TAG.log 2>&1
    echo "ncu full2 exit $?" | tee -a $OUT/status_$TAG.txt
  fi
fi
cat $OUT/status_$TAG.txt
