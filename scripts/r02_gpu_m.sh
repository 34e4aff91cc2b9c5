#!/bin/bash
# A/B of the barrier-poll sleep of the one-thread roles (block-level kernel timings)
OUT=gpurun_out; mkdir -p $OUT
for V in default mma0 mma100; do
  if [ $V = default ]; then LIB=""; else LIB=$PWD/build_ab/libirb200_$V.so; fi
  IRB200_LIB=$LIB timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_poll_$V.log 2>&1
  echo "$V: $(grep 'C96\|C48' $OUT/blocks_poll_$V.log | grep fp32 | python -c 'import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d["block"], d["kernels"]["gdfn_fused"]["ms"], d["kernels"]["mdta_fused_front"]["ms"], end="  ")')"
done
