#!/bin/bash
# Programmatic-dependent-launch A/B: block parity, full GPU suite, then bench / block kernels / batch-1 latency with
# the launch attribute on (IRB_PDL=1) and off (the default) on the same box.
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/status_$TAG.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -k "block" -x -q --timeout 300 > $OUT/pytest_block_$TAG.log 2>&1
RC=$?; echo "block tests exit $RC" | tee -a $OUT/status_$TAG.txt; tail -5 $OUT/pytest_block_$TAG.log
if [ "$RC" != "0" ]; then exit 1; fi
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt; tail -6 $OUT/pytest_$TAG.log
cp $OUT/parity.json $OUT/parity_$TAG.json 2>/dev/null
for V in pdl nopdl pdl2 nopdl2; do
  if [[ $V == nopdl* ]]; then unset IRB_PDL; else export IRB_PDL=1; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-tiled > $OUT/bench_${TAG}_$V.json 2> $OUT/bench_${TAG}_$V.err
  echo "bench $V exit $?" | tee -a $OUT/status_$TAG.txt; head -c 260 $OUT/bench_${TAG}_$V.json; echo; tail -2 $OUT/bench_${TAG}_$V.err
done
for V in pdl nopdl; do
  if [[ $V == nopdl* ]]; then unset IRB_PDL; else export IRB_PDL=1; fi
  timeout 600 python scripts/bench_latency.py > $OUT/latency_${TAG}_$V.json 2> $OUT/latency_${TAG}_$V.err
  echo "latency $V exit $?" | tee -a $OUT/status_$TAG.txt; head -c 1500 $OUT/latency_${TAG}_$V.json; echo
done
unset IRB_PDL
cat $OUT/status_$TAG.txt
