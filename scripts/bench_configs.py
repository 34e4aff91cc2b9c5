#!/usr/bin/env python
"""Single-GPU throughput of every BASELINE.json configuration's shape (synthetic data, random-init weights), both
arithmetic modes, device-resident inputs, CUDA-event timed.  Config 2 is bench.py's headline; this script is the table
for the others.  Prints one JSON line per (config, mode); writes gpurun_out/configs.json.

    python scripts/bench_configs.py [--steps 5]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import image_restoration_models_b200 as M
    from image_restoration_models_b200 import tiling
    import oracle
    from oracle.make_golden_tiling import make_image
    torch.set_grad_enabled(False)
    dev = torch.device("cuda", 0)
    out = []

    def emit(row):
        print(json.dumps(row), flush=True)
        out.append(row)

    # config 1: DnCNN-S gray sigma=25 (17 x conv3x3 + BN + ReLU), one 256x256 image; plus a batch for throughput
    dsd = oracle.synth_state_dict(oracle.dncnn_schema(1, 1, 64, 17, "BR"), 8)
    for mode in ("fp32", "half"):
        d = M.DnCNN(1, 1, 64, 17, "BR").eval()
        d.load_state_dict(dsd, strict=True)
        d = d.to(dev)
        if hasattr(d, "set_mode"):
            d.set_mode(mode)
        for B in (1, 32):
            x = oracle.synth_image((B, 1, 256, 256), 9, 25.0).to(dev)
            ms = timed(lambda: d(x), args.steps * 4)
            emit({"config": 1, "name": f"DnCNN-S gray 256x256 batch {B}", "mode": mode, "ms": ms,
                  "mpix_per_s": B * 256 * 256 / 1e6 / (ms / 1e3)})

    def restormer(task, mode):
        kw = oracle.RESTORMER_TASKS[task]
        m = M.Restormer(**kw, bias=False).eval()
        m.load_state_dict(oracle.synth_state_dict(oracle.restormer_schema(**kw), 7), strict=True)
        return m.to(dev).set_mode(mode), kw

    for mode in ("fp32", "half"):
        # config 2: gray denoise, batch 8 of 512x512
        m, kw = restormer("gray_denoise", mode)
        x = oracle.synth_image((8, 1, 512, 512), 100, 25.0).to(dev)
        ms = timed(lambda: m(x), args.steps)
        emit({"config": 2, "name": "Restormer gray denoise 8 x 512x512", "mode": mode, "ms": ms,
              "mpix_per_s": 8 * 512 * 512 / 1e6 / (ms / 1e3)})
        del m
        # config 3: real denoise (SIDD shape), batch 32 of 256x256
        m, kw = restormer("real_denoise", mode)
        x = oracle.synth_image((32, 3, 256, 256), 101, 25.0).to(dev)
        ms = timed(lambda: m(x), args.steps)
        emit({"config": 3, "name": "Restormer real denoise 32 x 256x256", "mode": mode, "ms": ms,
              "mpix_per_s": 32 * 256 * 256 / 1e6 / (ms / 1e3)})
        del m
        # config 4: motion deblur, 1280x720 uint8 frames through the device-side tiled harness (6 tiles of 512x512)
        m, kw = restormer("motion_deblur", mode)
        frames = [make_image("uint8", 720, 1280, 3, 100 + i) for i in range(4)]
        ms = timed(lambda: [tiling._run_local(m, f, dev, 512, 96, True, 6) for f in frames], max(1, args.steps // 2), warmup=1)
        emit({"config": 4, "name": "Restormer motion deblur 4 x 1280x720 (tiled 512/96, uint8 in/out, host frames)",
              "mode": mode, "ms": ms, "mpix_per_s": 4 * 720 * 1280 / 1e6 / (ms / 1e3),
              "computed_tile_mpix_per_s": 4 * 6 * 512 * 512 / 1e6 / (ms / 1e3)})
        del m
        # config 5: dual-pixel defocus, 1680x1120 uint16 frames with 6 channels (12 tiles of 512x512)
        m, kw = restormer("defocus_dual", mode)
        frames = [make_image("uint16", 1120, 1680, 6, 200 + i) for i in range(2)]
        ms = timed(lambda: [tiling._run_local(m, f, dev, 512, 96, True, 12) for f in frames], max(1, args.steps // 2), warmup=1)
        emit({"config": 5, "name": "Restormer dual-pixel defocus 2 x 1680x1120 (tiled 512/96, uint16 6-ch in, 3-ch out)",
              "mode": mode, "ms": ms, "mpix_per_s": 2 * 1120 * 1680 / 1e6 / (ms / 1e3),
              "computed_tile_mpix_per_s": 2 * 12 * 512 * 512 / 1e6 / (ms / 1e3)})
        del m
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
