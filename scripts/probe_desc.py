#!/usr/bin/env python
"""Hardware probe: which base-offset value makes a row-shifted SWIZZLE_128B operand descriptor read the rows TMA wrote?
Small-integer inputs make the tf32 products exact.  Prints, per shift, the base-offset values that give the right result."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image_restoration_models_b200 import _native

lib = _native.lib()
g = torch.Generator().manual_seed(0)
a = torch.randint(-8, 9, (160, 32), generator=g).float().cuda()
w = torch.randint(-4, 5, (32, 32), generator=g).float().cuda()
st = torch.cuda.current_stream().cuda_stream
for shift in range(0, 18):
    ref = a[shift:shift + 128] @ w.t()
    good = []
    for bo in range(8):
        d = torch.full((128, 32), float("nan"), device="cuda")
        _native.check(lib.ir_probe_shifted_descriptor(a.data_ptr(), w.data_ptr(), d.data_ptr(), shift, bo, st))
        torch.cuda.synchronize()
        if torch.equal(d, ref):
            good.append(bo)
    print(f"shift {shift:2d}: base_offset values that match = {good}", flush=True)
