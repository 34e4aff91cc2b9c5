"""Tiled inference with the tile cut, normalisation, synthetic degradation, reflect pad, Gaussian-window blend and
output conversion on the GPU, and all tiles of an image batched through the model.

Mirror of the reference harness ``run_model_inference`` (/root/reference/src/utils.py:353-454) and its helpers
``get_gaussian_weights`` (:314-350), ``pad`` (:174-181), ``normalize`` (:159-171), ``add_gaussian_noise`` (:29-36) for
the Restormer / DnCNN path: same positional and keyword arguments, same tile grid, same result (bit-exact given the same
tile predictions; the reference runs one batch-1 forward per tile with a D2H copy and a numpy accumulate in between).

Two entry points:

* ``run_model_inference``  one image, the reference's signature (``get_model_prediction``, src/utils.py:281-311, can call
  it unchanged);
* ``FramePipeline``        a stream of frames: pinned double-buffered staging, H2D of frame k+1 and D2H of frame k-1 on
  their own streams while frame k computes, no per-frame host synchronisation.

Multi-GPU: tiles and frames are independent, so there is no collective on the data path.  A stream of frames is split by
frame (each rank restores its own).  One frame can also be split by tile (contiguous slices of the tile list per rank);
the tile predictions are then gathered to ONE rank (``dst``), which blends in the reference's tile order, so its result is
bit-identical to the single-GPU one; the other ranks return ``None``.
"""
from __future__ import annotations

import time
from typing import Callable

import numpy as np
import torch

from . import _native

_DTYPES = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}
_TORCH_DTYPES = {0: torch.uint8, 1: torch.uint16, 2: torch.float32}


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


def normalize(img: np.ndarray):
    """The reference's default ``normalize`` (src/utils.py:159-171) on the host.  ``run_model_inference`` does this on the
    device inside the tile gather kernel; the function exists so that ``normalize=tiling.normalize`` is accepted."""
    if img.dtype == np.uint16:
        return (img.astype(np.float32) / 65535.0).astype(np.float32)
    if img.dtype == np.uint8:
        return (img.astype(np.float32) / 255.0).astype(np.float32)
    max_val = np.max(img)
    return (img.astype(np.float32) / max_val if max_val > 1.0 else img.astype(np.float32)).astype(np.float32)


_default_normalize = normalize


def pad(x: torch.Tensor, downscale_factor: int = 8):
    """The reference's ``utils.pad`` (src/utils.py:174-181), exported so that callers can pass ``pad=tiling.pad`` exactly
    as ``get_model_prediction`` passes ``pad=pad``.  Inside ``run_model_inference`` any non-None ``pad`` selects the
    device-side reflect pad of the tile gather kernel, which computes the same thing."""
    h, w = x.shape[-2:]
    padh = padded_extent(h, downscale_factor) - h
    padw = padded_extent(w, downscale_factor) - w
    return torch.nn.functional.pad(x, (0, padw, 0, padh), "reflect")


def noise_field(th: int, tw: int, channels: int, noise_level) -> np.ndarray:
    """The float64 field ``add_gaussian_noise`` (src/utils.py:29-36) adds to a [th, tw, C] patch.  The reference calls
    ``np.random.seed(0)`` before every draw, so the field depends only on the patch shape and sigma: it is a constant of
    the run, generated once with the same numpy generator (legacy MT19937 + polar Gaussian, loc + scale * g in double)
    and uploaded; the device adds it per tile."""
    return np.random.RandomState(0).normal(0, noise_level / 255., (th, tw, channels))


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

    def gather(self, img_dev, dtype_code, divisor, H, W, Cc, xy_dev, T, th, tw, TH, TW, noise_dev=None):
        out = torch.empty((T, Cc, TH, TW), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(_native.lib().ir_tile_gather(img_dev.data_ptr(), dtype_code, float(divisor), H, W, Cc,
                                                   xy_dev.data_ptr(), T, th, tw, TH, TW,
                                                   None if noise_dev is None else noise_dev.data_ptr(),
                                                   out.data_ptr(), stream))
        return out

    def blend(self, pred, xy_dev, T, th, tw, TH, TW, window_dev, H, W, Cc, out_dtype, dtype_code, scale, lo, hi, out=None):
        if out is None:
            out = torch.empty((H, W, Cc), dtype=out_dtype, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(_native.lib().ir_tile_blend(pred.data_ptr(), xy_dev.data_ptr(), T, th, tw, TH, TW,
                                                  window_dev.data_ptr(), window_dev.shape[1], H, W, Cc, out.data_ptr(),
                                                  dtype_code, float(scale), float(lo), float(hi), stream))
        return out


class _Geometry:
    """Everything about a frame's tiling that depends only on its shape, dtype and the patch configuration."""

    def __init__(self, shape, dtype, patch_size, patch_overlap, use_pad):
        if len(shape) != 3:
            raise ValueError("expected an HWC image")
        if np.dtype(dtype) not in _DTYPES:
            raise ValueError(f"unsupported image dtype {dtype}")
        self.h, self.w, self.cin = (int(v) for v in shape)   # every input channel goes to the model (dual-pixel: 6, :405)
        self.cout = min(3, self.cin)                          # channels of the output image / weight map (:394-395)
        h_idx, w_idx, self.ps = tile_grid(self.h, self.w, patch_size, patch_overlap)
        self.th, self.tw = min(self.ps, self.h), min(self.ps, self.w)
        self.TH, self.TW = (padded_extent(self.th), padded_extent(self.tw)) if use_pad else (self.th, self.tw)
        self.xy = np.array([(hi, wi) for hi in h_idx for wi in w_idx], dtype=np.int32)
        self.T = len(self.xy)
        self.code = _DTYPES[np.dtype(dtype)]

    def scaling(self, img: np.ndarray):
        """(divisor of normalize :159-171, scale / lo / hi of the output conversion :443-450)."""
        if self.code == 0:
            return 255.0, 255.0, 0.0, 255.0
        if self.code == 1:
            return 65535.0, 65535.0, 0.0, 65535.0
        mx, mn = float(np.max(img)), float(np.min(img))
        return (mx if mx > 1.0 else 1.0), mx, mn, mx


def _check_hooks(normalize, postprocess):
    if normalize is not None and normalize is not _default_normalize:
        raise NotImplementedError("run_model_inference: only the default normalize (src/utils.py:159-171) runs on the "
                                  "device; the DeblurGANv2 hooks are outside this path")
    if postprocess is not None:
        raise NotImplementedError("run_model_inference: postprocess hooks (DeblurGANv2) are outside this path")


def run_model_inference(model, input_img: np.ndarray, device, normalize: Callable | None = None,
                        patch_size: int | None = None, patch_overlap: int = 32, need_degradation=False,
                        noise_level: int | float | None = None, pad: Callable | bool | None = None,
                        postprocess: Callable | None = None, progress_bar=None, *, tile_batch: int = 16, backend=None,
                        group=None, solo: bool = False, dst: int = 0):
    """Returns (restored image with the input's dtype, inference time in ms), like the reference (src/utils.py:353-454),
    with the reference's parameter order and defaults.

    pad: None (the default; DnCNN path) runs the tiles as cut; a callable (``tiling.pad`` / the reference's ``pad``) or
    True selects the Restormer path: reflect-pad each tile to a multiple of 8 and crop the prediction back.
    need_degradation + noise_level: ``add_gaussian_noise`` on every normalised tile (on the device, bit-exact).
    normalize / postprocess: only the defaults are on this path (anything else raises NotImplementedError).
    progress_bar: accepted; ``.tqdm(None, desc, total)`` / ``.update()`` are driven per tile batch when given.
    Keyword-only extras: `tile_batch` tiles go through the model per forward; with an initialised torch.distributed
    process group (and not `solo`) the tiles are split over the ranks and rank `dst` blends and returns the image (the
    others return ``(None, ms)``)."""
    t0 = time.time()
    _check_hooks(normalize, postprocess)
    if not isinstance(input_img, np.ndarray) or input_img.ndim != 3:
        raise ValueError("expected an HWC image")
    use_pad = pad is not None and pad is not False
    geo = _Geometry(input_img.shape, input_img.dtype, patch_size, patch_overlap, use_pad)
    be = backend or CudaBackend(device)
    divisor, scale, lo, hi_ = geo.scaling(input_img)
    noisy = bool(need_degradation) and noise_level is not None

    import torch.distributed as dist
    world = dist.get_world_size(group) if (not solo and dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo_t, hi_t = partition(geo.T, world)[rank]
    if progress_bar is not None:
        progress_bar = progress_bar.tqdm(None, desc="Processing patches", total=geo.T)

    with torch.no_grad():
        img_dev = be.upload(input_img)
        xy_dev = be.upload(geo.xy)
        window = be.upload(get_gaussian_weights(geo.ps, geo.ps, 1)[:, :, 0])
        noise_dev = be.upload(noise_field(geo.th, geo.tw, geo.cin, noise_level)) if noisy else None
        mine = xy_dev[lo_t:hi_t].contiguous()
        preds = []
        step = max(1, tile_batch)
        for s in range(0, hi_t - lo_t, step):
            e = min(hi_t - lo_t, s + step)
            tiles = be.gather(img_dev, geo.code, divisor, geo.h, geo.w, geo.cin, mine[s:e].contiguous(), e - s, geo.th,
                              geo.tw, geo.TH, geo.TW, noise_dev)
            pred = model(tiles)
            if pred.shape[1] != geo.cout:
                raise ValueError(f"model returned {pred.shape[1]} channels for an output image of {geo.cout}")
            preds.append(pred)
            if progress_bar is not None:
                for _ in range(e - s):
                    progress_bar.update()
        local = torch.cat(preds, 0) if preds else torch.empty((0, geo.cout, geo.TH, geo.TW), dtype=torch.float32,
                                                              device=img_dev.device)
        if world > 1:
            # the one exchange step, and only when ONE frame is split: the blending rank collects the tile predictions
            # (3 MB per 512x512 tile); nobody else receives or blends anything
            counts = [b - a for a, b in partition(geo.T, world)]
            width = max(counts)
            padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            padded[: local.shape[0]] = local
            gathered = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
            dist.gather(padded, gathered, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
            if rank != dst:
                return None, (time.time() - t0) * 1000.0
            local = torch.cat([g[:n] for g, n in zip(gathered, counts)], 0)
        out = be.blend(local.contiguous(), xy_dev, geo.T, geo.th, geo.tw, geo.TH, geo.TW, window, geo.h, geo.w, geo.cout,
                       _TORCH_DTYPES[geo.code], geo.code, scale, lo, hi_)
        result = out.cpu().numpy()
    return result, (time.time() - t0) * 1000.0


def _run_local(model, input_img, device, patch_size, patch_overlap, use_pad, tile_batch):
    """Single-rank restore of one image even when a process group exists (frames partitioned over ranks)."""
    return run_model_inference(model, input_img, device, patch_size=patch_size, patch_overlap=patch_overlap,
                               pad=pad if use_pad else None, tile_batch=tile_batch, solo=True)[0]


class FramePipeline:
    """Restores a stream of same-shaped frames with copies overlapped with compute.

    Per frame k: [copy stream] pinned host -> device image; [compute stream] tile gather -> model forward(s) -> blend;
    [output stream] device image -> pinned host.  Two staging slots per direction, ordered by CUDA events only: the host
    blocks when it reuses a slot (frame k-2 must have left it) and once at the end.  The arithmetic is the single-frame
    path's, so every frame is bit-identical to ``run_model_inference`` on it.

    >>> pipe = FramePipeline(model, device, frames[0].shape, frames[0].dtype, patch_size=512, patch_overlap=96, pad=True)
    >>> outs = pipe.run(frames)
    """

    SLOTS = 2

    def __init__(self, model, device, frame_shape, frame_dtype, patch_size=None, patch_overlap=32, pad=None,
                 need_degradation=False, noise_level=None, tile_batch=16):
        self.model = model
        self.device = torch.device(device)
        self.be = CudaBackend(self.device)
        use_pad = pad is not None and pad is not False
        self.geo = geo = _Geometry(frame_shape, frame_dtype, patch_size, patch_overlap, use_pad)
        if geo.code == 2:
            raise NotImplementedError("FramePipeline: float frames need a per-frame min / max pass; use "
                                      "run_model_inference for them")
        self.tile_batch = max(1, tile_batch)
        tdt = _TORCH_DTYPES[geo.code]
        with torch.cuda.device(self.device):
            self.copy_stream = torch.cuda.Stream()
            self.compute_stream = torch.cuda.Stream()
            self.out_stream = torch.cuda.Stream()
            self.h_in = [torch.empty((geo.h, geo.w, geo.cin), dtype=tdt).pin_memory() for _ in range(self.SLOTS)]
            self.h_out = [torch.empty((geo.h, geo.w, geo.cout), dtype=tdt).pin_memory() for _ in range(self.SLOTS)]
            self.d_in = [torch.empty((geo.h, geo.w, geo.cin), dtype=tdt, device=self.device) for _ in range(self.SLOTS)]
            self.d_out = [torch.empty((geo.h, geo.w, geo.cout), dtype=tdt, device=self.device) for _ in range(self.SLOTS)]
            self.xy = torch.from_numpy(geo.xy).to(self.device)
            self.window = torch.from_numpy(get_gaussian_weights(geo.ps, geo.ps, 1)[:, :, 0].copy()).to(self.device)
            noisy = bool(need_degradation) and noise_level is not None
            self.noise = (torch.from_numpy(noise_field(geo.th, geo.tw, geo.cin, noise_level)).to(self.device)
                          if noisy else None)
            self.preds = torch.empty((geo.T, geo.cout, geo.TH, geo.TW), dtype=torch.float32, device=self.device)
            torch.cuda.synchronize(self.device)
        ev = lambda: [torch.cuda.Event() for _ in range(self.SLOTS)]
        self.ev_h2d, self.ev_gathered, self.ev_blend, self.ev_d2h = ev(), ev(), ev(), ev()
        self.ev_start, self.ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.last_device_ms = None           # device time of the last run(): first H2D issued -> last D2H complete

    def _compute(self, slot):
        geo, be = self.geo, self.be
        divisor, scale, lo, hi_ = geo.scaling(None)
        for s in range(0, geo.T, self.tile_batch):
            e = min(geo.T, s + self.tile_batch)
            tiles = be.gather(self.d_in[slot], geo.code, divisor, geo.h, geo.w, geo.cin, self.xy[s:e], e - s, geo.th,
                              geo.tw, geo.TH, geo.TW, self.noise)
            if e == geo.T:
                self.ev_gathered[slot].record()          # the input slot may be overwritten from here on
            self.preds[s:e] = self.model(tiles)
        be.blend(self.preds, self.xy, geo.T, geo.th, geo.tw, geo.TH, geo.TW, self.window, geo.h, geo.w, geo.cout,
                 _TORCH_DTYPES[geo.code], geo.code, scale, lo, hi_, out=self.d_out[slot])

    def run(self, frames, copy_out: bool = True):
        """Restores every frame; returns the list of restored numpy frames (or None when copy_out is False: the timed
        region of a benchmark still moves every result into pinned host memory, it just skips the numpy copy)."""
        geo = self.geo
        results = [None] * len(frames)
        pending = [None] * self.SLOTS            # frame index whose result sits in h_out[slot]

        def collect(slot):
            k = pending[slot]
            if k is None:
                return
            self.ev_d2h[slot].synchronize()
            if copy_out:
                results[k] = self.h_out[slot].numpy().copy()
            pending[slot] = None

        with torch.no_grad(), torch.cuda.device(self.device):
            self.ev_start.record(self.copy_stream)
            for k, frame in enumerate(frames):
                if frame.shape != (geo.h, geo.w, geo.cin) or _DTYPES.get(frame.dtype) != geo.code:
                    raise ValueError("FramePipeline: every frame must have the shape / dtype given at construction")
                slot = k % self.SLOTS
                if k >= self.SLOTS:
                    # frame k-2 has left this slot's pinned input (its H2D copy is done) and its result is collected
                    self.ev_h2d[slot].synchronize()
                    collect(slot)
                self.h_in[slot].copy_(torch.from_numpy(frame))               # host memcpy into pinned staging
                with torch.cuda.stream(self.copy_stream):
                    if k >= self.SLOTS:
                        self.copy_stream.wait_event(self.ev_gathered[slot])  # frame k-2's tiles have been cut from d_in
                    self.d_in[slot].copy_(self.h_in[slot], non_blocking=True)
                    self.ev_h2d[slot].record()
                with torch.cuda.stream(self.compute_stream):
                    self.compute_stream.wait_event(self.ev_h2d[slot])
                    if k >= self.SLOTS:
                        self.compute_stream.wait_event(self.ev_d2h[slot])    # d_out[slot] has been copied out
                    self._compute(slot)
                    self.ev_blend[slot].record()
                with torch.cuda.stream(self.out_stream):
                    self.out_stream.wait_event(self.ev_blend[slot])
                    self.h_out[slot].copy_(self.d_out[slot], non_blocking=True)
                    self.ev_d2h[slot].record()
                pending[slot] = k
            self.ev_end.record(self.out_stream)
            for slot in range(self.SLOTS):
                collect((len(frames) + slot) % self.SLOTS)
            torch.cuda.synchronize(self.device)
            self.last_device_ms = self.ev_start.elapsed_time(self.ev_end) if len(frames) else 0.0
        return results if copy_out else None
