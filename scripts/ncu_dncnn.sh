#!/bin/bash
# ncu --set full over a few DnCNN body-layer launches (batch 32 of 256x256); CSV pages exported on the box.
TAG=${1:-dn}
MODE=${2:-fp32}
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/dn.py <<PY
import sys, torch
sys.path.insert(0, "$PWD")
import image_restoration_models_b200 as M, oracle
torch.set_grad_enabled(False)
dsd = oracle.synth_state_dict(oracle.dncnn_schema(1, 1, 64, 17, "BR"), 8)
d = M.DnCNN(1, 1, 64, 17, "BR").eval(); d.load_state_dict(dsd, strict=True); d = d.cuda().set_mode("$MODE")
x = oracle.synth_image((32, 1, 256, 256), 9, 25.0).cuda()
d(x); torch.cuda.synchronize()
torch.cuda.profiler.start(); d(x); torch.cuda.synchronize(); torch.cuda.profiler.stop()
PY
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:conv3 -c 2 -f -o /tmp/prof_$TAG python /tmp/dn.py > $OUT/ncu_$TAG.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $OUT/ncu_raw_$TAG.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/ncu_sass_$TAG.csv 2>/dev/null
gzip -f $OUT/ncu_sass_$TAG.csv
