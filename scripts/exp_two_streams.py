#!/usr/bin/env python
"""Experiment: the bench batch (8 x 512 x 512, config 2) as ONE forward against two half-batches on two streams (the
small-footprint HBM-bound kernels of one stream can share an SM with the FP32-pipe-bound fused kernels of the other).
    python scripts/exp_two_streams.py > gpurun_out/two_streams.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_restoration_models_b200 as M  # noqa: E402
import oracle  # noqa: E402

torch.set_grad_enabled(False)
dev = "cuda:0"
kw = oracle.RESTORMER_TASKS["gray_denoise"]
sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 5)
m = M.Restormer(**kw, bias=False).eval()
m.load_state_dict(sd, strict=True)
m = m.to(dev)
x = torch.rand(8, 1, 512, 512, device=dev)
parts = int(os.environ.get("PARTS", "2"))
xs = list(x.chunk(parts))
streams = [torch.cuda.Stream() for _ in range(parts)]


def one():
    return m(x)


def split():
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(cur)
    outs = []
    for s, xi in zip(streams, xs):
        s.wait_event(ev)
        with torch.cuda.stream(s):
            outs.append(m(xi))
    for s in streams:
        cur.wait_stream(s)
    return outs


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


y1 = one()
y2 = torch.cat(split())
res = {"max_abs_diff": float((y1 - y2).abs().max()), "parts": parts}
for rep in range(2):
    res[f"one_ms_{rep}"] = timeit(one)
    res[f"split_ms_{rep}"] = timeit(split)
print(json.dumps(res, indent=1))
