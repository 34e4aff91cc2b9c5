#!/bin/bash
# A/B of library variants on the block shapes of every level: bash scripts/r02_gpu_ab.sh <tag> <variant> [<variant> ...]
# (variant = suffix of build_ab/libirb200_<variant>.so; "default" = the shipped library)
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/status_$TAG.txt
for V in "$@"; do
  if [ $V = default ]; then LIB=""; else LIB=$PWD/build_ab/libirb200_$V.so; fi
  IRB200_LIB=$LIB timeout 300 python scripts/bench_kernels.py --blocks --levels > $OUT/blocks_${TAG}_$V.log 2>&1
  echo "== $V exit $?" | tee -a $OUT/status_$TAG.txt
  grep fp32 $OUT/blocks_${TAG}_$V.log | python -c '
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d["block"], " ".join("%s=%.4f" % (k, v["ms"]) for k, v in d["kernels"].items()))' | tee -a $OUT/status_$TAG.txt
done
