#!/usr/bin/env python
"""Benchmark of the hot path: Restormer forward, Mpix/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A "step" is one forward over one batch of synthetic images.  Workload at every N: BASELINE config 2,
"Restormer gray Gaussian denoise (1->1 ch, dim=48, blocks [4,6,6,8]), batch 8 of synthetic 512x512",
one such batch PER GPU (weak scaling; images are independent, there is no collective on the data path).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TASK = "gray_denoise"
BATCH, HEIGHT, WIDTH = 8, 512, 512
SIGMA = 25.0
METRIC = "restormer_fwd_mpix_per_s"
DTYPES = {
    # arithmetic the path computes in; both modes accumulate in fp32 and keep the residual stream, LayerNorm
    # statistics, softmax and GELU in fp32, and both meet the north-star parity bar (max-abs <= 1e-3, dPSNR <= 0.01 dB)
    "fp32": "fp32 residual stream / accumulators / statistics / softmax / GELU, tf32 operands in the 3x3 convolutions; fp16 (same "
            "10-bit mantissa) where a tensor is only ever a tensor-core operand (norm1 / norm2 output, v, folded attention "
            "matrix, the fused kernels' on-chip operands) and for the intermediates of the two low-resolution levels (C > 128)",
    "half": "fp16 intermediates + fp16 tensor-core operands, fp32 accumulate and fp32 residual stream",
    "bf16": "bf16 intermediates + bf16 tensor-core operands, fp32 accumulate and fp32 residual stream (outside the 1e-3 "
            "parity bar: reported separately)",
}
UNIT = "Mpix/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_model(device):
    import torch
    import image_restoration_models_b200 as M
    import oracle
    kw = oracle.RESTORMER_TASKS[TASK]
    model = M.Restormer(**kw, bias=False).eval()
    model.load_state_dict(oracle.synth_state_dict(oracle.restormer_schema(**kw), 7), strict=True)
    return model.to(device), kw


def cpu_reference_run(steps, warmup, sample_hw=None, budget_s=200.0):
    """The reference algorithm (oracle port == the same ATen ops the reference modules call) on all host cores."""
    import torch
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_grad_enabled(False)
    kw = oracle.RESTORMER_TASKS[TASK]
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 7)
    if sample_hw is None:
        # calibrate on a 64x64 crop, then take the largest square crop of one 512x512 image that keeps the run bounded
        x = oracle.synth_image((1, kw["inp_channels"], 64, 64), 17, SIGMA)
        oracle.restormer_forward(sd, x)
        t0 = time.perf_counter(); oracle.restormer_forward(sd, x); t64 = time.perf_counter() - t0
        sample_hw = 64
        for hw in (512, 256, 128):
            if t64 * (hw / 64.0) ** 2 * 1.3 * (steps + warmup) <= budget_s:
                sample_hw = hw
                break
    x = oracle.synth_image((1, kw["inp_channels"], sample_hw, sample_hw), 17, SIGMA)
    for _ in range(warmup):
        oracle.restormer_forward(sd, x)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        oracle.restormer_forward(sd, x)
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    mpix = sample_hw * sample_hw / 1e6 / mean
    return mpix, mean * 1e3, cores, f"1 image of {sample_hw}x{sample_hw} (crop of one 512x512 batch element), fp32, {steps} timed runs"


def load_tensor_peaks():
    """profiles/tensor_peaks.json: dense tf32 / fp16 / bf16 matmul peaks measured on this pool's B200 with the
    MEASURED_PEAKS recipe (scripts/measure_tensor_peaks.py: torch.matmul 8192^3, best of 10 and 4 s back to back)."""
    path = os.path.join(ROOT, "profiles", "tensor_peaks.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            pass
    return None


def gpu_eager_baseline(dev, steps):
    """What `model.cuda()` of the reference runs (scripts/tests.py:428, src/utils.py:412-417): the same ATen op sequence
    (oracle port) on CUDA tensors through cuDNN / cuBLAS, config 2, CUDA-event timed, TF32 off and on.  Runs after, and
    outside, every timed region of the product."""
    import torch
    import oracle
    kw = oracle.RESTORMER_TASKS[TASK]
    sd = {k: v.to(dev) for k, v in oracle.synth_state_dict(oracle.restormer_schema(**kw), 7).items()}
    x = oracle.synth_image((BATCH, kw["inp_channels"], HEIGHT, WIDTH), 100, SIGMA).to(dev)
    out = {}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True
            for _ in range(2):
                oracle.restormer_forward(sd, x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                oracle.restormer_forward(sd, x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": BATCH * HEIGHT * WIDTH / 1e6 / (ms / 1e3), "unit": UNIT, "ms_per_step": ms}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = saved
    out["what"] = ("eager PyTorch (cuDNN / cuBLAS ATen ops, the reference's own GPU path) on the same B200, same batch, "
                   f"{steps} timed forwards after 2 warm-up, cudnn.benchmark on")
    del sd, x
    torch.cuda.empty_cache()
    return out


TILED_FRAME = (720, 1280, 3)          # BASELINE config 4: GoPro-shape uint8 frames
TILED_PATCH, TILED_OVERLAP = 512, 96  # src/configs.py:29-33 -> 6 tiles per frame


def tiled_config4(dev, rank, world, mode, barrier, reduce_max, distinct=8, passes=6):
    """BASELINE config 4 under the same clock as the headline: Restormer motion deblur (WithBias) on 1280x720 uint8 frames
    through the device-side tiled harness (512 / 96 -> 6 tiles per frame), `distinct * passes` frames PER GPU, frames
    partitioned over the ranks (no collective).  uint8 host frames in, uint8 host frames out (pinned staging, copies
    inside the timed region): this number is end to end by construction."""
    import torch
    import image_restoration_models_b200 as M
    from image_restoration_models_b200 import tiling
    import oracle
    from oracle.make_golden_tiling import make_image
    kw = oracle.RESTORMER_TASKS["motion_deblur"]
    model = M.Restormer(**kw, bias=False).eval()
    model.load_state_dict(oracle.synth_state_dict(oracle.restormer_schema(**kw), 7), strict=True)
    model = model.to(dev).set_mode(mode)
    fh, fw, fc = TILED_FRAME
    frames = [make_image("uint8", fh, fw, fc, 100 + rank * 1000 + i) for i in range(distinct)]
    pipe = tiling.FramePipeline(model, dev, frames[0].shape, frames[0].dtype, patch_size=TILED_PATCH,
                                patch_overlap=TILED_OVERLAP, pad=tiling.pad, tile_batch=6)
    ntile = pipe.geo.T
    pipe.run(frames[:3], copy_out=False)                 # warm-up: packs the weights, allocates the workspace
    work = frames * passes
    barrier()
    t0 = time.perf_counter()
    pipe.run(work, copy_out=False)                       # ends with a device synchronise: every result is in host memory
    wall_ms = (time.perf_counter() - t0) * 1e3
    # device time: event on the copy stream before the first H2D -> event on the output stream after the last D2H
    ms = reduce_max(pipe.last_device_ms)
    wall_ms = reduce_max(wall_ms)
    nfr = len(work)
    out_mpix = world * nfr * fh * fw / 1e6
    tile_mpix = world * nfr * ntile * TILED_PATCH * TILED_PATCH / 1e6
    launches = nfr * (model.launches_per_forward() + 3)  # + tile gather, prediction copy, blend
    del pipe, model
    torch.cuda.empty_cache()
    return {"metric": "restormer_tiled_deblur_output_mpix_per_s", "value": out_mpix / (ms / 1e3), "unit": UNIT,
            "computed_tile_mpix_per_s": tile_mpix / (ms / 1e3), "n_gpus": world, "frames_per_gpu": nfr,
            "tiles_per_frame": ntile, "frame": [fh, fw, fc], "patch": TILED_PATCH, "overlap": TILED_OVERLAP,
            "ms_total": ms, "ms_per_frame": ms / nfr, "host_wall_ms": wall_ms, "mode": mode, "scaling": "weak",
            "e2e": {"value": out_mpix / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": fh * fw * fc,
                    "d2h_bytes_per_step": fh * fw * fc, "step": "one frame"},
            "gpu_launches": launches,
            "workload": "Restormer motion deblur (3->3 ch, WithBias), 1280x720 uint8 frames, 512x512 tiles with overlap 96, "
                        f"{nfr} frames per GPU ({distinct} distinct), frames partitioned over the GPUs (BASELINE config 4)",
            "pipeline": "pinned double-buffered staging; H2D, compute (gather -> forward -> blend) and D2H on three streams"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="fp32", choices=["fp32", "half"])
    ap.add_argument("--no-tiled", action="store_true", help="skip the tiled config-4 object")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-CUDA baseline")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
    steps = max(args.steps, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything a library prints there (e.g. the NCCL version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    config = {"workload": "Restormer gray Gaussian denoise (1->1 ch, dim=48, blocks [4,6,6,8]), batch 8 of synthetic "
                          "512x512 per GPU (BASELINE config 2, the config the metric is quoted on; north_star's colour 512x512 "
                          "denoise forward differs only in patch_embed / output: 3 instead of 1 image channels)",
              "task": TASK, "batch_per_gpu": BATCH, "height": HEIGHT, "width": WIDTH, "weights": "random-init (seeded)",
              "partition": "by image, one batch per GPU, no collective",
              "cache": "working set ~12 GB per step >> 126 MB L2, so every step streams from HBM"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        mpix, ms, cores, sample = cpu_reference_run(steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": mpix, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": mpix, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    import oracle
    from image_restoration_models_b200 import _native
    torch.set_grad_enabled(False)
    model, kw = build_model(dev)
    model.set_mode(args.mode)
    x_host = oracle.synth_image((BATCH, kw["inp_channels"], HEIGHT, WIDTH), 100 + rank, SIGMA).pin_memory()
    y_host = torch.empty((BATCH, kw["out_channels"], HEIGHT, WIDTH), dtype=torch.float32).pin_memory()
    x_dev = x_host.to(dev)
    pix_per_step = BATCH * HEIGHT * WIDTH

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ----------------------------------------------------------
    for _ in range(warmup):
        y = model(x_dev)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        y = model(x_dev)
    e1.record()
    barrier()
    ms_total = reduce_max(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / steps
    value = world * pix_per_step / 1e6 / (ms_step / 1e3)

    # ---- end to end through the public API: pinned host input -> forward -> host output ---------
    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        yd = model(xd)
        y_host.copy_(yd, non_blocking=True)
    e2e_step()
    barrier()
    e0.record()
    for _ in range(steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = reduce_max(e0.elapsed_time(e1)) / steps
    e2e_value = world * pix_per_step / 1e6 / (e2e_ms / 1e3)

    # ---- per-kernel breakdown (separate pass: two events per launch) ------------------------------
    # (IRB_NCU_RANGE=1: this one forward is the cudaProfiler range that scripts/ncu_traffic.sh captures)
    ncu_range = os.environ.get("IRB_NCU_RANGE") == "1"
    if ncu_range:
        torch.cuda.profiler.start()
    with _native.kernel_profile() as prof:
        model(x_dev)
        torch.cuda.synchronize()
    if ncu_range:
        torch.cuda.profiler.stop()
    rows = sorted(prof.rows, key=lambda r: -r["ms"])
    hbm_peak, bf16_peak, peak_kind = load_peaks()
    tpeaks = load_tensor_peaks()
    # tensor roofline denominators (dense, sustained): tf32 for the fp32-mode contractions that take tf32 operands, fp16
    # for those that take fp16 operands (fused GDFN, attention front / output in fp32 mode; everything in half mode)
    tf32_peak = tpeaks["tf32_tflops_sustained"] if tpeaks else bf16_peak / 2.0
    f16_peak = tpeaks["fp16_tflops_sustained"] if tpeaks else bf16_peak
    tpeak_src = ("profiles/tensor_peaks.json (torch.matmul 8192^3, sustained)" if tpeaks else
                 "MEASURED_PEAKS.json bf16 sustained (tf32 taken as half of it: no measured tf32 peak on file)")
    # (fp32 mode: everything but the 3x3 convolutions and reduce_chan takes fp16 operands -- the fused kernels, the attention
    # output, and the 16-bit plan of the two low-resolution levels)
    F16_FAMILIES = {"gdfn_fused", "dwconv_qkv_gram", "attn_out_1x1", "mdta_fused_front", "ln_qkv_1x1", "ln_project_in_1x1",
                    "ffn_project_out_1x1", "dwconv_gate_project_out", "dwconv3x3_gelu_gate"}
    step_ms_sum = sum(r["ms"] for r in rows)
    kernels = []
    for r in rows:
        gbs = r["bytes"] / 1e9 / (r["ms"] / 1e3) if r["ms"] > 0 else 0.0
        tfs = r["flops"] / 1e12 / (r["ms"] / 1e3) if r["ms"] > 0 else 0.0
        tp = f16_peak if (args.mode == "half" or r["name"] in F16_FAMILIES) else tf32_peak
        kernels.append({"name": r["name"], "launches": r["launches"], "ms": round(r["ms"], 3),
                        "GBps": round(gbs, 1), "hbm_frac": round(gbs / hbm_peak, 4),
                        "TFLOPs": round(tfs, 2), "tensor_frac": round(tfs / tp, 4),
                        "best_frac": round(max(gbs / hbm_peak, tfs / tp), 4)})
    # The two fused kernels are bound by the FP32 pipe (depthwise taps / gate on CUDA cores), not by HBM or the tensor pipe:
    # useful FP32-pipe operations per pixel = 27 hp (GDFN: 18 hp taps + 9 hp gate) / 29 C (front: 27 C taps + 2 C norms),
    # against 128 FMA/clk/SM x 148 SMs x the SM clock sampled during the timed region
    _d, _nb, _nr = kw.get("dim", 48), kw.get("num_blocks", [4, 6, 6, 8]), kw.get("num_refinement_blocks", 4)
    _pix = BATCH * HEIGHT * WIDTH
    _hp_of = lambda c: -(-int(c * kw.get("ffn_expansion_factor", 2.66)) // 16) * 16
    _cfg = [(_d, _pix, _nb[0]), (2 * _d, _pix // 4, 2 * _nb[1]), (2 * _d, _pix, _nb[0] + _nr)]   # (C, pixels, launches)
    _clk = ((clocks or {}).get("sm_mhz") if rank == 0 else None) or 1965.0
    _fma_peak = 128.0 * 148 * _clk * 1e6
    # Measured issue rates on this pool's B200 (scripts/probe_ffma2.cu, profiles/r02_probe_ffma2.jsonl): scalar FFMA 126
    # FMA/clk/SM, packed FFMA2 106.7 FMA/clk/SM (2.40 cycles per instruction per scheduler).  The fused kernels use FFMA2
    # (half the issue slots per FMA: their streams are 37 % non-FMA instructions), so 106.7 is the ceiling they run against.
    _ffma2_ceiling = 106.7 / 128.0
    _useful = {"gdfn_fused": sum(n * p * 27.0 * _hp_of(c) for c, p, n in _cfg),
               "mdta_fused_front": sum(n * p * 29.0 * c for c, p, n in _cfg)}
    for k in kernels:
        if k["name"] in _useful and k["ms"] > 0:
            k["fp32_pipe_frac"] = round(_useful[k["name"]] / (k["ms"] / 1e3) / _fma_peak, 4)
            k["ffma2_ceiling_frac"] = round(k["fp32_pipe_frac"] / _ffma2_ceiling, 4)
            k["best_frac"] = round(max(k["best_frac"], k["fp32_pipe_frac"]), 4)
    top = rows[0]
    top_ms_per_launch = top["ms"] / top["launches"]
    achieved = top["bytes"] / top["launches"] / 1e9 / (top_ms_per_launch / 1e3)
    # DRAM bytes per launch of that family from the committed ncu pass of this same command (scripts/ncu_traffic.sh)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", f"traffic_{args.mode}.json")
    if os.path.exists(tpath):
        try:
            fam = json.load(open(tpath))["families"].get(top["name"])
            if fam and fam["launches"] == top["launches"]:
                traffic = fam["dram_bytes_per_launch"]
                traffic_src = f"profiles/traffic_{args.mode}.json (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
        except Exception:
            pass
    roofline = {"kernel": top["name"], "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "avg_launch_ms": top_ms_per_launch, "share_of_step": top["ms"] / step_ms_sum,
                "algorithmic_bytes_per_launch": top["bytes"] / top["launches"]}
    top_tfs = top["flops"] / 1e12 / (top["ms"] / 1e3)
    top_tp = f16_peak if (args.mode == "half" or top["name"] in F16_FAMILIES) else tf32_peak
    roofline["tensor"] = {"achieved": top_tfs, "peak": top_tp, "unit": "TFLOP/s", "frac": top_tfs / top_tp,
                          "peak_source": tpeak_src}
    if top["name"] == "gdfn_fused":
        # The fused GDFN moves 10*C bytes per pixel where the kernels it replaces moved ~55*C: neither HBM nor the tensor
        # pipe bounds it.  Its limiter is the FP32 pipe: per pixel 9 FMA for each of the 2*hp depthwise outputs and ~12
        # FMA-equivalents per gated element for the erf gate.  Reported against 128 FMA/clk/SM x 148
        # SMs x the SM clock sampled in this run; `bound` says so instead of naming a roofline the kernel is not on.
        hp_total = 0.0      # sum over launches of pixels * hp  ==  flops term 36*hp*pix / 36
        d = kw.get("dim", 48)
        nb, nr = kw.get("num_blocks", [4, 6, 6, 8]), kw.get("num_refinement_blocks", 4)
        pix = BATCH * HEIGHT * WIDTH
        hp_of = lambda c: -(-int(c * kw.get("ffn_expansion_factor", 2.66)) // 16) * 16
        launches_cfg = [(d, pix, nb[0]), (2 * d, pix // 4, 2 * nb[1]), (2 * d, pix, nb[0] + nr)]   # (C, pixels, launches)
        fma = sum(n * p * (18.0 + 9.0) * hp_of(c) for c, p, n in launches_cfg)
        clk = (clocks or {}).get("sm_mhz") if rank == 0 else None
        clk = clk or 1965.0
        fma_peak = 128.0 * 148 * clk * 1e6
        fma_rate = fma / (top["ms"] / 1e3)
        roofline.update({"bound": "fp32_pipe", "hbm": {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                                                       "frac": achieved / hbm_peak},
                         "achieved": fma_rate / 1e12, "peak": fma_peak / 1e12, "unit": "TFMA/s",
                         "frac": fma_rate / fma_peak,
                         "peak_source": f"128 FMA/clk/SM x 148 SMs x {clk:.0f} MHz (SM clock sampled during the timed region)",
                         "useful_fma_per_step": fma,
                         "frac_of_measured_ffma2_rate": fma_rate / fma_peak / (106.7 / 128.0),
                         "measured_issue_rates": "scalar FFMA 126 FMA/clk/SM, packed FFMA2 106.7 FMA/clk/SM (scripts/probe_ffma2.cu)",
                         "limiter": "fp32 FMA pipe (depthwise taps + GELU gate on CUDA cores), not HBM or the tensor pipe"})
        ppath = os.path.join(ROOT, "profiles", "fused_gdfn_pipes.json")
        if os.path.exists(ppath):
            try:
                roofline["limiter_evidence"] = json.load(open(ppath))
            except Exception:
                pass
        # the largest family that IS bound by HBM, same definition of achieved / peak
        # (a family whose FP32-pipe or tensor fraction exceeds its HBM fraction is not HBM-bound: the fused front is skipped)
        rest = [r for r, k in zip(rows, kernels) if r["name"] != "gdfn_fused" and r["bytes"] > 0 and
                k["hbm_frac"] >= max(k.get("fp32_pipe_frac", 0.0), k["tensor_frac"])]
        if rest:
            h = rest[0]
            h_ms = h["ms"] / h["launches"]
            h_ach = h["bytes"] / h["launches"] / 1e9 / (h_ms / 1e3)
            roofline["top_hbm_bound_kernel"] = {"kernel": h["name"], "bound": "hbm", "achieved": h_ach, "peak": hbm_peak,
                                                "unit": "GB/s", "frac": h_ach / hbm_peak, "avg_launch_ms": h_ms,
                                                "share_of_step": h["ms"] / step_ms_sum}
    # share of the step spent in kernels at >= 0.6 of their best roofline (north_star: >= 0.6 per kernel)
    roofline["step_share_at_0p6"] = sum(r["ms"] for r, k in zip(rows, kernels) if k["best_frac"] >= 0.6) / step_ms_sum
    step_bytes = sum(r["bytes"] for r in rows)
    step_flops = sum(r["flops"] for r in rows)

    # ---- the other arithmetic mode, same workload, device-resident timing only ---------------------
    other = "half" if args.mode == "fp32" else "fp32"
    model.set_mode(other)
    for _ in range(warmup):
        y = model(x_dev)
    barrier()
    e0.record()
    for _ in range(steps):
        y = model(x_dev)
    e1.record()
    barrier()
    other_ms = reduce_max(e0.elapsed_time(e1)) / steps
    other_mode = {"mode": other, "value": world * pix_per_step / 1e6 / (other_ms / 1e3), "unit": UNIT,
                  "ms_per_step": other_ms,
                  "dtype": DTYPES[other], "parity": "same bar as the headline mode (tests/test_gpu_parity.py)"}
    # bf16 mode (BASELINE config 3 "fp32 and bf16"), reported separately: outside the parity bar by construction
    model.set_mode("bf16")
    for _ in range(warmup):
        y = model(x_dev)
    barrier()
    e0.record()
    for _ in range(steps):
        y = model(x_dev)
    e1.record()
    barrier()
    bf16_ms = reduce_max(e0.elapsed_time(e1)) / steps
    bf16_mode = {"mode": "bf16", "value": world * pix_per_step / 1e6 / (bf16_ms / 1e3), "unit": UNIT, "ms_per_step": bf16_ms,
                 "dtype": DTYPES["bf16"],
                 "parity": "NOT inside the 1e-3 bar (8 mantissa bits); measured max-abs in profiles/r02_parity.json, "
                           "tests/test_gpu_parity.py::test_restormer_bf16_mode_reported_separately"}
    model.set_mode(args.mode)

    # ---- BASELINE config 4 (tiled full-frame deblur) at the same N, frames partitioned over the ranks ----------
    tiled = None
    if not args.no_tiled:
        model._workspace = None            # release the headline workload's scratch first
        torch.cuda.empty_cache()
        tiled = tiled_config4(dev, rank, world, args.mode, barrier, reduce_max)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        mpix, ms, cores, sample = cpu_reference_run(3, 1, sample_hw=256)
        cpu_baseline = {"value": mpix, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                        "ms_per_sample": ms}

    eager = None
    if world == 1 and not args.no_eager:
        try:
            model._workspace = None
            torch.cuda.empty_cache()
            eager = gpu_eager_baseline(dev, 3)
        except Exception as exc:            # a baseline must never take the bench line down
            eager = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPES[args.mode], "mode": args.mode, "other_mode": other_mode, "bf16_mode": bf16_mode,
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": y_host.numel() * 4, "ms_per_step": e2e_ms},
            "gpu_launches": steps * model.launches_per_forward(),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "gpu_eager_baseline": eager, "tiled_config4": tiled,
            "step_algorithmic_GB": step_bytes / 1e9, "step_TFLOP": step_flops / 1e12,
            "step_hbm_frac": step_bytes / 1e9 / (ms_step / 1e3) / hbm_peak,
            "kernels": kernels}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
