#!/bin/bash
# ncu --set full over one TransformerBlock per (mode, C); exports CSV pages on the box (the .ncu-rep embeds the
# whole cubin per kernel and exceeds what gpurun copies back).  Usage: bash scripts/ncu_blocks.sh <tag>
TAG=${1:-blk}
OUT=gpurun_out
mkdir -p $OUT
timeout 800 ncu --set full --clock-control none --profile-from-start off ${NCU_K:+-k regex:$NCU_K} -f -o /tmp/prof_$TAG \
    python scripts/bench_kernels.py --ncu > $OUT/ncu_$TAG.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $OUT/ncu_raw_$TAG.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page details --csv > $OUT/ncu_details_$TAG.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/ncu_sass_$TAG.csv 2>/dev/null
gzip -f $OUT/ncu_sass_$TAG.csv
ls -la $OUT /tmp/prof_$TAG.ncu-rep
