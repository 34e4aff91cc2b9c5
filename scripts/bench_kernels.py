#!/usr/bin/env python
"""Kernel-level timing of the 1x1 contraction on Restormer block shapes (CUDA events via the library's profiler).
Usage: IRB200_LIB=/path/to/libirb200.so python scripts/bench_kernels.py [--half]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image_restoration_models_b200 import _native  # noqa: E402

# name, k1, k2, N, ln_mode, resid
SHAPES = [
    ("K1_c48", 48, 0, 144, 1, False), ("K4_c48", 48, 0, 48, 0, True), ("K5_c48", 48, 0, 256, 1, False),
    ("K6_c48", 128, 0, 48, 0, True),
    ("K1_c96", 96, 0, 288, 1, False), ("K4_c96", 96, 0, 96, 0, True), ("K5_c96", 96, 0, 512, 1, False),
    ("K6_c96", 256, 0, 96, 0, True),
]


def run(name, k1, k2, N, ln, resid, rows, half):
    lib = _native.lib()
    dev = "cuda"
    K = k1 + k2
    a_half = int(half and ln == 0)
    y_half = int(half and not resid)
    a1 = torch.randn(rows, k1, device=dev)
    if a_half:
        a1 = a1.half()
    w = torch.randn(N, K, device=dev) / K ** 0.5
    lw = torch.ones(k1, device=dev)
    y = torch.zeros(rows, N, device=dev, dtype=torch.float16 if y_half else torch.float32)
    scratch = torch.empty((N * (K + 64) + rows * K) * 4 + 1024, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def call():
        _native.check(lib.ir_test_conv1x1(0, a1.data_ptr(), k1, k1, 0, 0, 0, w.data_ptr(), 0, ln, lw.data_ptr(), lw.data_ptr(),
                                          y.data_ptr() if resid else 0, N, y.data_ptr(), N, 8, rows // 8, N, 1,
                                          a_half, int(half), y_half, scratch.data_ptr(), scratch.numel(), st))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    with _native.kernel_profile() as prof:
        for _ in range(5):
            call()
        torch.cuda.synchronize()
    r = [x for x in prof.rows if x["name"] == "other"][0]
    ms = r["ms"] / r["launches"]
    return {"name": name, "ms": round(ms, 4), "GBps": round(r["bytes"] / r["launches"] / 1e9 / (ms / 1e3), 1)}


def main():
    half = "--half" in sys.argv
    rows = 8 * 512 * 512
    only = [a.split("=")[1] for a in sys.argv if a.startswith("--only=")]
    out = [run(*s, rows, half) for s in SHAPES if not only or s[0] in only]
    print(json.dumps({"lib": os.environ.get("IRB200_LIB", "default"), "half": half, "results": out}))




def run_block(Cc, heads, B, H, W, mode, ncu=False):
    """All eight kernels of one TransformerBlock at full resolution, per-family times from the profiler."""
    import numpy as np
    lib = _native.lib()
    dev = "cuda"
    hidden = int(Cc * 2.66)
    g = torch.Generator().manual_seed(0)
    shapes = [(Cc,), (heads, 1, 1), (3 * Cc, Cc, 1, 1), (3 * Cc, 1, 3, 3), (Cc, Cc, 1, 1), (Cc,), (2 * hidden, Cc, 1, 1),
              (2 * hidden, 1, 3, 3), (Cc, hidden, 1, 1)]
    params = [(torch.rand(s, generator=g) - 0.5).to(dev) * 0.2 + (1.0 if len(s) == 1 else 0.0) for s in shapes]
    nbytes = lib.ir_block_packed_bytes(Cc, heads, 2.66, 0, 0, mode)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _native.check(lib.ir_block_pack_weights(Cc, heads, 2.66, 0, 0, _native.ptr_array(params), len(params), packed.data_ptr(),
                                            nbytes, mode, st))
    ws = torch.empty(lib.ir_block_workspace_bytes(Cc, heads, 2.66, B, H, W, mode), dtype=torch.uint8, device=dev)
    x = torch.randn(B * H * W, Cc, device=dev)
    def call():
        _native.check(lib.ir_block_forward(Cc, heads, 2.66, 0, 0, packed.data_ptr(), x.data_ptr(), B, H, W, ws.data_ptr(),
                                           ws.numel(), mode, st))
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    if ncu:
        torch.cuda.profiler.start()
        call()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return {}
    with _native.kernel_profile() as prof:
        for _ in range(3):
            call()
        torch.cuda.synchronize()
    return {r["name"]: {"ms": round(r["ms"] / r["launches"], 4),
                        "GBps": round(r["bytes"] / r["launches"] / 1e9 / (r["ms"] / r["launches"] / 1e3), 0)} for r in prof.rows}


if "--ncu" in sys.argv:
    # one profiled forward of one block per (mode, C): run under `ncu --profile-from-start off`
    Bn = 2
    modes = [int(m) for m in os.environ.get("NCU_MODES", "0,1").split(",")]
    cs = [int(c) for c in os.environ.get("NCU_CS", "48,96,192").split(",")]
    # pixel counts of the bench workload's levels (8 x 512 x 512 input): C = 192 -> 131 k pixels, C = 384 -> 32 k pixels
    side = {48: 512, 96: 512, 192: 256, 384: 128}
    for mode in modes:
        for Cc, heads in ((48, 1), (96, 1), (192, 4), (384, 8)):
            if Cc not in cs:
                continue
            lib = _native.lib()
            run_block(Cc, heads, Bn, side[Cc], side[Cc], mode, ncu=True)
    sys.exit(0)

if "--blocks" in sys.argv:
    for mode, nm in ((0, "fp32"), (1, "half")):
        for Cc, heads in ((48, 1), (96, 1)):
            print(json.dumps({"block": f"C{Cc}h{heads}", "mode": nm, "kernels": run_block(Cc, heads, 8, 512, 512, mode)}))
        if "--levels" in sys.argv:
            # the lower-resolution levels of the bench workload (8 x 512 x 512 input)
            for Cc, heads, hw in ((96, 2, 256), (192, 4, 128), (384, 8, 64)):
                print(json.dumps({"block": f"C{Cc}h{heads}@{hw}", "mode": nm, "kernels": run_block(Cc, heads, 8, hw, hw, mode)}))

if __name__ == "__main__" and "--blocks" not in sys.argv:
    main()
