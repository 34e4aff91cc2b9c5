#!/usr/bin/env python
"""BASELINE config 4: Restormer motion deblur on GoPro-shape 1280x720 uint8 frames through the device-side tiled
harness (512x512 tiles, overlap 96 -> 6 tiles per frame, src/configs.py:29-33), frames partitioned over the GPUs
of one box (no data-path collective).  Launch under torchrun for N > 1.  Prints one JSON line on rank 0.

    python scripts/bench_tiled.py --frames 8 [--mode half]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_tiled.py --frames 8
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8, help="frames PER GPU (weak scaling)")
    ap.add_argument("--mode", default="fp32", choices=["fp32", "half"])
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--config", type=int, default=4, choices=[4, 5],
                    help="4: motion deblur 1280x720 uint8 (6 tiles); 5: dual-pixel defocus 1680x1120 uint16 6-ch (12 tiles)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import image_restoration_models_b200 as M
    from image_restoration_models_b200 import tiling
    import oracle
    from oracle.make_golden_tiling import make_image
    torch.set_grad_enabled(False)
    task, dt, fh, fw, fc, ntile = (("motion_deblur", "uint8", 720, 1280, 3, 6) if args.config == 4 else
                                   ("defocus_dual", "uint16", 1120, 1680, 6, 12))
    kw = oracle.RESTORMER_TASKS[task]
    model = M.Restormer(**kw, bias=False).eval()
    model.load_state_dict(oracle.synth_state_dict(oracle.restormer_schema(**kw), 7), strict=True)
    model = model.to(dev).set_mode(args.mode)
    frames = [make_image(dt, fh, fw, fc, 100 + rank * 1000 + i) for i in range(args.frames)]
    # frames are independent: every rank restores its own frames (group=None inside would split tiles instead)
    def run(fs):
        return [tiling._run_local(model, f, dev, 512, 96, True, ntile) for f in fs]
    for _ in range(args.warmup):
        run(frames[:1])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = run(frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    if rank == 0:
        mpix = world * args.frames * fh * fw / 1e6
        print(json.dumps({"metric": "restormer_tiled_deblur_output_mpix_per_s", "config": args.config, "task": task,
                          "frame": [fh, fw, fc], "value": mpix / (ms / 1e3), "unit": "Mpix/s",
                          "n_gpus": world, "frames_per_gpu": args.frames, "tiles_per_frame": ntile, "mode": args.mode,
                          "ms_total": ms, "ms_per_frame": ms / args.frames,
                          "computed_tile_mpix_per_s": world * args.frames * ntile * 512 * 512 / 1e6 / (ms / 1e3),
                          "checksum": int(np.sum(outs[0].astype(np.int64)))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
