#!/bin/bash
# Round-2 GPU pass N: norm1 of the next block from the GDFN epilogue: parity, bench A/B against the LayerNorm pass
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity.json $OUT/status_$TAG.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt; tail -5 $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-tiled > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; tail -2 $OUT/bench_$TAG.err
IRB_NO_NORM1_CHAIN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-tiled > $OUT/bench_${TAG}_nochain.json 2> $OUT/bench_${TAG}_nochain.err
python - <<PY
import json
for f in ("$OUT/bench_$TAG.json", "$OUT/bench_${TAG}_nochain.json"):
    d = json.load(open(f))
    ks = {k["name"]: (k["launches"], k["ms"]) for k in d["kernels"]}
    print(f, round(d["value"], 2), "Mpix/s", round(d["ms_per_step"], 2), "ms", "half", round(d["other_mode"]["value"], 2), {n: ks[n] for n in ("gdfn_fused", "layernorm", "mdta_fused_front", "attn_out_1x1")})
PY
