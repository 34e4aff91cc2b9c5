#!/usr/bin/env python
"""Batch-1 latency of the reference harness's real call pattern (one tile per forward, src/utils.py:403-419): Restormer on a
256x256 and a 512x512 tile, DnCNN-S on 256x256; plain launches vs the CUDA-graph cache vs eager PyTorch (cuDNN / cuBLAS, the
reference's own GPU path, TF32 on).  Host wall time per call with a synchronise (what the harness sees) and device time.

    python scripts/bench_latency.py > gpurun_out/latency.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_restoration_models_b200 as M  # noqa: E402
import oracle  # noqa: E402

torch.set_grad_enabled(False)


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    # per-call latency as the harness sees it: call + synchronise
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        torch.cuda.synchronize()
    sync = (time.perf_counter() - t0) / n * 1e3
    return {"device_ms": e0.elapsed_time(e1) / n, "host_ms_pipelined": wall, "host_ms_call_plus_sync": sync}


def main():
    out = {"what": "batch-1 latency, fp32 mode; eager = the oracle's ATen op sequence on CUDA tensors with TF32 on"}
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    for task, hw in (("color_denoise", 256), ("color_denoise", 512), ("motion_deblur", 512)):
        kw = oracle.RESTORMER_TASKS[task]
        sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 7)
        m = M.Restormer(**kw, bias=False).eval()
        m.load_state_dict(sd, strict=True)
        m = m.cuda()
        x = oracle.synth_image((1, 3, hw, hw), 11, 25.0).cuda()
        sdc = {k: v.cuda() for k, v in sd.items()}
        m.set_cuda_graphs(False)
        plain = timeit(lambda: m(x))
        m.set_cuda_graphs(True)
        graph = timeit(lambda: m(x))
        row = {"plain": plain, "graph": graph,
               "eager_tf32": timeit(lambda: oracle.restormer_forward(sdc, x), n=10, warm=3),
               "launches": m.launches_per_forward()}
        out[f"restormer_{task}_1x3x{hw}x{hw}"] = row
    dsd = oracle.synth_state_dict(oracle.dncnn_schema(1, 1, 64, 17, "R"), 8)
    d = M.DnCNN(1, 1, 64, 17, "R").eval()
    d.load_state_dict(dsd, strict=True)
    d = d.cuda()
    dx = oracle.synth_image((1, 1, 256, 256), 9, 25.0).cuda()
    dsdc = {k: v.cuda() for k, v in dsd.items()}
    d.set_cuda_graphs(False)
    dplain = timeit(lambda: d(dx))
    d.set_cuda_graphs(True)
    dgraph = timeit(lambda: d(dx))
    out["dncnn_s_1x1x256x256"] = {"plain": dplain, "graph": dgraph,
                                  "eager_tf32": timeit(lambda: oracle.dncnn_forward(dsdc, dx), n=10, warm=3), "launches": 17}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
