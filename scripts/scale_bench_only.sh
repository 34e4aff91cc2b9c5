#!/bin/bash
# bench.py (config 2, weak scaling) on N GPUs of one box.  Usage (under gpurun --gpus N): bash scripts/scale_bench_only.sh N
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $OUT/scale_bench_n${N}_R.json 2> $OUT/scale_bench_n${N}_R.err
echo "bench n=$N exit $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29602 scripts/bench_tiled.py --frames 4 > $OUT/scale_tiled_c4_n${N}_R.json 2> $OUT/scale_tiled_c4_n${N}_R.err
echo "tiled c4 n=$N exit $?"
grep -h '"metric"' $OUT/scale_tiled_c4_n${N}_R.json | cut -c1-330
python - <<PY
import json
for l in open("$OUT/scale_bench_n${N}_R.json"):
    if l.startswith("{"):
        d = json.loads(l); print("bench", d["n_gpus"], round(d["value"], 2), "Mpix/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 2), d["clocks"], "other", round(d["other_mode"]["value"], 2))
PY
