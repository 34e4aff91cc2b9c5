#!/bin/bash
# ncu launch list of ONE forward of the bench workload: duration and DRAM bytes of every launch, joined with the
# library's own per-launch family tags.  Usage (under gpurun): bash scripts/ncu_traffic.sh <tag> [fp32|half]
TAG=${1:-t}
MODE=${2:-fp32}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/ncu_tags_${TAG}_$MODE.csv
IRB_NCU_RANGE=1 IRB_PROFILE_DUMP=$OUT/ncu_tags_${TAG}_$MODE.csv timeout 1200 ncu \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file $OUT/ncu_launches_${TAG}_$MODE.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --mode $MODE > $OUT/ncu_traffic_${TAG}_$MODE.log 2>&1
echo "ncu traffic ($MODE) exit $?"
python scripts/summarize_traffic.py $OUT/ncu_launches_${TAG}_$MODE.csv $OUT/ncu_tags_${TAG}_$MODE.csv \
    $OUT/launches_${TAG}_$MODE.csv $OUT/traffic_${TAG}_$MODE.json
