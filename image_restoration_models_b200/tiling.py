"""Tiled inference with the tile cut, normalisation, reflect pad, Gaussian-window blend and output conversion on
the GPU, and all tiles of an image batched through the model.

Mirror of the reference harness ``run_model_inference`` (/root/reference/src/utils.py:353-454) and its helpers
``get_gaussian_weights`` (:314-350), ``pad`` (:174-181), ``normalize`` (:159-171) for the Restormer / DnCNN path:
same arguments, same tile grid, same result (bit-exact given the same tile predictions; the reference runs one
batch-1 forward per tile with a D2H copy and a numpy accumulate in between).

Multi-GPU: tiles are independent, so ranks take contiguous slices of the tile list; predictions are all-gathered
(the only exchange step, 3 MB per 512x512 tile) and every rank blends in the reference's tile order, which keeps the
result bit-identical to the single-GPU one.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from . import _native

_DTYPES = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}


def get_gaussian_weights(height: int, width: int, n_channels=3, sigma_scale=0.125):
    """Same window as the reference (centre = size/2.0, sigma = size*sigma_scale; float64 math, float32 result)."""
    y_grid, x_grid = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    g = np.exp(-((y_grid - height / 2.0) ** 2 / (2 * (height * sigma_scale) ** 2)
                 + (x_grid - width / 2.0) ** 2 / (2 * (width * sigma_scale) ** 2)))
    return np.repeat(g[:, :, np.newaxis], n_channels, axis=2).astype(np.float32)


def tile_grid(h: int, w: int, patch_size, patch_overlap: int):
    """Tile origins in the reference's loop order (src/utils.py:383-392)."""
    if patch_size:
        patch_size = min(patch_size, max(h, w))
        stride = max(patch_size - patch_overlap, 1)
        h_idx = list(range(0, h - patch_size, stride)) + [max(h - patch_size, 0)]
        w_idx = list(range(0, w - patch_size, stride)) + [max(w - patch_size, 0)]
    else:
        patch_size = max(h, w)
        h_idx, w_idx = [0], [0]
    return h_idx, w_idx, patch_size


def padded_extent(n: int, factor: int = 8) -> int:
    """Extent after the reference's pad(): next multiple of `factor` only when not already one (src/utils.py:174-181)."""
    return n if n % factor == 0 else ((n + factor) // factor) * factor


def partition(n_items: int, world: int):
    """Contiguous slices of the tile list, one per rank (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((start, start + n))
        start += n
    return out


class CudaBackend:
    """Tile gather / blend kernels of libirb200.so (ir_tile_gather, ir_tile_blend)."""

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tiled inference runs on CUDA only (no CPU fallback)")

    def upload(self, img: np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(img)).to(self.device, non_blocking=True)

    def gather(self, img_dev, dtype_code, divisor, H, W, Cc, xy_dev, T, th, tw, TH, TW):
        out = torch.empty((T, Cc, TH, TW), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(_native.lib().ir_tile_gather(img_dev.data_ptr(), dtype_code, float(divisor), H, W, Cc,
                                                   xy_dev.data_ptr(), T, th, tw, TH, TW, out.data_ptr(), stream))
        return out

    def blend(self, pred, xy_dev, T, th, tw, TH, TW, window_dev, H, W, Cc, out_dtype, dtype_code, scale, lo, hi):
        out = torch.empty((H, W, Cc), dtype=out_dtype, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(_native.lib().ir_tile_blend(pred.data_ptr(), xy_dev.data_ptr(), T, th, tw, TH, TW,
                                                  window_dev.data_ptr(), window_dev.shape[1], H, W, Cc, out.data_ptr(),
                                                  dtype_code, float(scale), float(lo), float(hi), stream))
        return out


def _run_local(model, input_img, device, patch_size, patch_overlap, pad, tile_batch):
    """Single-rank restore of one image even when a process group exists (frames partitioned over ranks)."""
    return run_model_inference(model, input_img, device, patch_size, patch_overlap, pad, tile_batch, solo=True)[0]


def run_model_inference(model, input_img: np.ndarray, device, patch_size=None, patch_overlap: int = 32, pad: bool = True,
                        tile_batch: int = 16, backend=None, group=None, solo: bool = False):
    """Returns (restored image with the input's dtype, inference time in ms), like the reference.

    pad=True is the Restormer path (reflect-pad each tile to a multiple of 8, crop the prediction back);
    pad=False is the DnCNN path.  `tile_batch` tiles go through the model per forward.  With an initialised
    torch.distributed process group the tiles are split over the ranks."""
    t0 = time.time()
    if input_img.ndim != 3:
        raise ValueError("expected an HWC image")
    if input_img.dtype not in _DTYPES:
        raise ValueError(f"unsupported image dtype {input_img.dtype}")
    be = backend or CudaBackend(device)
    h, w = input_img.shape[:2]
    Cin = input_img.shape[2]                  # every input channel goes to the model (dual-pixel: 6, src/utils.py:405)
    Cc = min(3, Cin)                          # channels of the output image / weight map (src/utils.py:394-395)
    h_idx, w_idx, ps = tile_grid(h, w, patch_size, patch_overlap)
    th, tw = min(ps, h), min(ps, w)
    TH, TW = (padded_extent(th), padded_extent(tw)) if pad else (th, tw)
    xy = np.array([(hi, wi) for hi in h_idx for wi in w_idx], dtype=np.int32)
    T = len(xy)
    code = _DTYPES[input_img.dtype]
    if code == 0:
        divisor, scale, lo, hi_ = 255.0, 255.0, 0.0, 255.0
    elif code == 1:
        divisor, scale, lo, hi_ = 65535.0, 65535.0, 0.0, 65535.0
    else:
        mx, mn = float(np.max(input_img)), float(np.min(input_img))
        divisor = mx if mx > 1.0 else 1.0
        scale, lo, hi_ = mx, mn, mx

    import torch.distributed as dist
    world = dist.get_world_size(group) if (not solo and dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo_t, hi_t = partition(T, world)[rank]

    with torch.no_grad():
        img_dev = be.upload(input_img)
        xy_dev = be.upload(xy)
        window = be.upload(get_gaussian_weights(ps, ps, 1)[:, :, 0])
        mine = xy_dev[lo_t:hi_t].contiguous()
        preds = []
        for s in range(0, hi_t - lo_t, max(1, tile_batch)):
            e = min(hi_t - lo_t, s + max(1, tile_batch))
            tiles = be.gather(img_dev, code, divisor, h, w, Cin, mine[s:e].contiguous(), e - s, th, tw, TH, TW)
            pred = model(tiles)
            if pred.shape[1] != Cc:
                raise ValueError(f"model returned {pred.shape[1]} channels for an output image of {Cc}")
            preds.append(pred)
        local = torch.cat(preds, 0) if preds else torch.empty((0, Cc, TH, TW), dtype=torch.float32,
                                                              device=img_dev.device)
        if world > 1:
            # the one exchange step: every rank needs every tile prediction to blend in the reference order
            counts = [b - a for a, b in partition(T, world)]
            width = max(counts)
            padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            padded[: local.shape[0]] = local
            gathered = [torch.empty_like(padded) for _ in range(world)]
            dist.all_gather(gathered, padded, group=group)
            local = torch.cat([g[:n] for g, n in zip(gathered, counts)], 0)
        out = be.blend(local.contiguous(), xy_dev, T, th, tw, TH, TW, window, h, w, Cc,
                       {0: torch.uint8, 1: torch.uint16, 2: torch.float32}[code], code, scale, lo, hi_)
        result = out.cpu().numpy()
    return result, (time.time() - t0) * 1000.0
