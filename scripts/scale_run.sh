#!/bin/bash
# Scaling evidence on N GPUs of one box: bench.py (config 2, weak scaling) and the tiled harness (configs 4 and 5).
# Usage (under gpurun --gpus N): bash scripts/scale_run.sh N
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29601 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/scale_bench_n$N.json 2> $OUT/scale_bench_n$N.err
echo "bench n=$N exit $?"
$TR --master-port 29602 scripts/bench_tiled.py --frames 4 > $OUT/scale_tiled_c4_n$N.json 2> $OUT/scale_tiled_c4_n$N.err
echo "tiled c4 n=$N exit $?"
$TR --master-port 29603 scripts/bench_tiled.py --frames 4 --mode half > $OUT/scale_tiled_c4_half_n$N.json 2>> $OUT/scale_tiled_c4_n$N.err
$TR --master-port 29604 scripts/bench_tiled.py --frames 2 --config 5 > $OUT/scale_tiled_c5_n$N.json 2> $OUT/scale_tiled_c5_n$N.err
echo "tiled c5 n=$N exit $?"
grep -h '"metric"' $OUT/scale_tiled_c4_n$N.json $OUT/scale_tiled_c4_half_n$N.json $OUT/scale_tiled_c5_n$N.json | cut -c1-330
python - <<PY
import json
for l in open("$OUT/scale_bench_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l); print("bench", d["n_gpus"], round(d["value"], 2), "Mpix/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 2), d["clocks"])
PY
