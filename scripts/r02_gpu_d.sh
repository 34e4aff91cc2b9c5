#!/bin/bash
# Round-2 GPU pass D: graph-cache test + batch-1 latencies, ffn_fused issue-loop A/B at block level.
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graph or block or fresh or guard or golden" > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee $OUT/status_$TAG.txt; tail -4 $OUT/pytest_$TAG.log
timeout 600 python scripts/bench_latency.py > $OUT/latency_$TAG.json 2> $OUT/latency_$TAG.err
echo "latency exit $?" | tee -a $OUT/status_$TAG.txt; python -c "
import json; d=json.load(open('$OUT/latency_$TAG.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {a:(round(b['host_ms_call_plus_sync'],3), round(b['device_ms'],3)) for a,b in v.items() if isinstance(b,dict)})
"
timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_$TAG.log 2>&1
echo "blocks exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_$TAG.log | cut -c1-400
IRB_FFN_GENERIC_ISSUE=1 timeout 300 python scripts/bench_kernels.py --blocks > $OUT/blocks_${TAG}_generic.log 2>&1
echo "blocks (generic ffn issue loop) exit $?" | tee -a $OUT/status_$TAG.txt; grep fp32 $OUT/blocks_${TAG}_generic.log | cut -c1-400
cat $OUT/status_$TAG.txt
