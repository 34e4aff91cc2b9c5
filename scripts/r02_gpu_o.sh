#!/bin/bash
# Round-2 GPU pass O: the 16-bit plan at the wide levels of the fp32 mode: full parity suite, bench A/B (IRB_NO_WIDE16)
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity.json $OUT/status_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt; tail -12 $OUT/pytest_$TAG.log
cp $OUT/parity.json $OUT/parity_$TAG.json 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-tiled > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?" | tee -a $OUT/status_$TAG.txt; tail -2 $OUT/bench_$TAG.err
IRB_NO_WIDE16=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-tiled > $OUT/bench_${TAG}_nowide16.json 2> $OUT/bench_${TAG}_nowide16.err
python - <<PY
import json
for f in ("$OUT/bench_$TAG.json", "$OUT/bench_${TAG}_nowide16.json"):
    d = json.load(open(f))
    print(f, round(d["value"], 2), "Mpix/s", round(d["ms_per_step"], 2), "ms", "half", round(d["other_mode"]["value"], 2), "GB", round(d["step_algorithmic_GB"],1))
p = json.load(open("$OUT/parity_$TAG.json"))
worst = sorted(((v.get("max_abs", 0), k) for k, v in p.items() if "bf16" not in k and "metrics" not in k and "max_abs" in v), reverse=True)[:6]
print(worst)
PY
