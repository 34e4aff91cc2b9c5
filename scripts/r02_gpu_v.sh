#!/bin/bash
# Round-2 GPU pass V: fold kernel grid.z / part-group sweep on the block shapes of every level
TAG=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/status_$TAG.txt
for V in "0 0" "1 0" "2 0" "4 0" "0 2" "1 2"; do
  set -- $V
  IRB_FOLD_ZB=$1 IRB_FOLD_PG=$2 timeout 300 python scripts/bench_kernels.py --blocks --levels > $OUT/blocks_${TAG}_zb$1_pg$2.log 2>&1
  echo "zb $1 pg $2: $(grep fp32 $OUT/blocks_${TAG}_zb$1_pg$2.log | python -c 'import sys,json; print(" ".join("%s=%.4f" % (json.loads(l)["block"], json.loads(l)["kernels"]["softmax_fold"]["ms"]) for l in sys.stdin))')" | tee -a $OUT/status_$TAG.txt
done
