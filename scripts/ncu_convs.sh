#!/bin/bash
# ncu --set full over the 3x3 convolution kernels of one bench step (CSV pages exported on the box)
TAG=${1:-conv}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none -k "regex:conv3" -c 8 -f -o /tmp/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-tiled > $OUT/ncu_$TAG.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $OUT/ncu_raw_$TAG.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/ncu_sass_$TAG.csv 2>/dev/null
gzip -f $OUT/ncu_sass_$TAG.csv
ls -la $OUT | tail -5
