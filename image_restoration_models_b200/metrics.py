"""PSNR / SSIM of a restored image on the GPU: mirror of ``calculate_metrics`` (/root/reference/src/utils.py:134-156),
which scripts/tests.py calls once per restored image (:61, :120, :179, :238, :289, :339, :392) on the uint8 / uint16 HWC
array ``run_model_inference`` returns.  Same arguments and return value (two Python floats); the images may also be CUDA
tensors (e.g. the device-resident output of ``tiling.FramePipeline``), in which case nothing but the two results crosses
PCIe.  No CPU fallback: the arithmetic is ``ir_image_metrics`` of libirb200.so.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native

_CODES = {torch.uint8: 0, torch.uint16: 1, torch.float32: 2}


def _to_device(img, device):
    if isinstance(img, np.ndarray):
        if img.dtype == np.float64:
            img = img.astype(np.float32)
        img = torch.from_numpy(np.ascontiguousarray(img))
    if not isinstance(img, torch.Tensor):
        raise TypeError("calculate_metrics: expected a numpy array or a torch tensor")
    return img.to(device, non_blocking=True).contiguous()


def calculate_metrics(pred, target, data_range: int | float | None = None, device=None):
    """(psnr, ssim) between prediction and target: HWC (3 or 1 channels) or HW images of equal shape and dtype."""
    if device is None:
        device = pred.device if isinstance(pred, torch.Tensor) and pred.is_cuda else torch.device("cuda")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("calculate_metrics runs on CUDA only (no CPU fallback)")
    if tuple(pred.shape) != tuple(target.shape):
        raise ValueError("Input images must have the same dimensions.")          # skimage's check_shape_equality
    p, t = _to_device(pred, device), _to_device(target, device)
    if p.dtype != t.dtype:
        raise ValueError(f"calculate_metrics: pred is {p.dtype} but target is {t.dtype}")
    if p.dtype not in _CODES:
        raise ValueError(f"calculate_metrics: unsupported dtype {p.dtype} (uint8, uint16 or float32)")
    if data_range is None:                       # utils.py:137-143
        data_range = 255 if p.dtype == torch.uint8 else 65535 if p.dtype == torch.uint16 else 1.0
    if p.dim() == 2:
        H, W, C = p.shape[0], p.shape[1], 1
    elif p.dim() == 3 and p.shape[2] in (1, 3):
        H, W, C = p.shape
    else:
        raise ValueError("calculate_metrics: expected an HW, HWx1 or HWx3 image")
    lib = _native.lib()
    ws = torch.empty(max(lib.ir_image_metrics_workspace_bytes(H, W, C), 256), dtype=torch.uint8, device=device)
    out = torch.empty(3, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        _native.check(lib.ir_image_metrics(p.data_ptr(), t.data_ptr(), _CODES[p.dtype], H, W, C, float(data_range),
                                           out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    psnr, ssim, _ = out.cpu().tolist()
    return psnr, ssim
