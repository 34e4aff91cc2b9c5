#!/usr/bin/env python
"""Dense tensor-core peaks of this pool's B200 with the MEASURED_PEAKS.json recipe: torch.matmul 8192^3 (2*N^3 flops),
best of 10 (burst) and back to back for 4 s (sustained), for tf32 (fp32 inputs, allow_tf32), fp16 and bf16.
cuBLAS here is the yardstick, not the product.  Writes gpurun_out/tensor_peaks.json (copied to profiles/)."""
import json
import os
import subprocess
import time

import torch

N = 8192


def measure(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(N, N, device="cuda", dtype=dtype)
    b = torch.randn(N, N, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.time()
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            a @ b
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    flops = 2.0 * N ** 3
    return flops / (best / 1e3) / 1e12, flops * n / (e0.elapsed_time(e1) / 1e3) / 1e12


def main():
    out = {"how": "torch.matmul 8192^3 (2*N^3), best of 10 = burst, back to back for 4 s = sustained; CUDA events",
           "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    for name, dtype, tf32 in (("tf32", torch.float32, True), ("fp16", torch.float16, False),
                              ("bf16", torch.bfloat16, False), ("fp32_simt", torch.float32, False)):
        burst, sus = measure(dtype, tf32)
        out[f"{name}_tflops"] = round(burst, 1)
        out[f"{name}_tflops_sustained"] = round(sus, 1)
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"],
                           capture_output=True, text=True).stdout.strip()
        out["nvidia_smi_after"] = q
    except Exception:
        pass
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/tensor_peaks.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
