#!/bin/bash
# compute-sanitizer pass over the hot path (SURVEY.md section 5 "Race detection"): memcheck, racecheck, synccheck and
# initcheck on scripts/sanitize_target.py (every block width, a whole small Restormer, DnCNN; results checked against
# the oracle).  Usage (under gpurun): bash scripts/sanitize.sh <tag> [tools...]
TAG=${1:-san}; shift
TOOLS=${@:-memcheck racecheck synccheck}
OUT=gpurun_out
mkdir -p $OUT
SAN=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
for tool in $TOOLS; do
  for what in blocks model; do
    log=$OUT/sanitizer_${TAG}_${tool}_${what}.log
    timeout ${SAN_TIMEOUT:-420} $SAN --tool $tool --print-limit 20 --error-exitcode 99 python scripts/sanitize_target.py $what > $log 2>&1
    rc=$?
    echo "$tool $what exit $rc : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)" | tee -a $OUT/sanitizer_${TAG}_summary.txt
  done
done
