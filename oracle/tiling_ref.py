"""CPU restatement of the reference's tiled-inference harness (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates /root/reference/src/utils.py: normalize :159-171, pad :174-181, get_gaussian_weights :314-350 and
run_model_inference :353-454 with the noise injection of add_gaussian_noise :29-36 (:408-409); the postprocess hooks are
omitted (DeblurGANv2 only).  ``model`` is any callable NCHW float32 torch tensor -> tensor.
"""
from __future__ import annotations

import numpy as np
import torch


def normalize(img: np.ndarray):
    """utils.normalize :159-171."""
    if img.dtype == np.uint16:
        out = img.astype(np.float32) / 65535.0
    elif img.dtype == np.uint8:
        out = img.astype(np.float32) / 255.0
    else:
        max_val = np.max(img)
        out = img.astype(np.float32) / max_val if max_val > 1.0 else img.astype(np.float32)
    return out.astype(np.float32)


def pad(x: torch.Tensor, downscale_factor: int = 8):
    """utils.pad :174-181: reflect-pad right/bottom up to the next multiple of 8 (only when not already one)."""
    h, w = x.shape[-2:]
    H = ((h + downscale_factor) // downscale_factor) * downscale_factor
    W = ((w + downscale_factor) // downscale_factor) * downscale_factor
    padh = H - h if h % downscale_factor != 0 else 0
    padw = W - w if w % downscale_factor != 0 else 0
    return torch.nn.functional.pad(x, (0, padw, 0, padh), "reflect")


def gaussian_weights(height: int, width: int, n_channels=3, sigma_scale=0.125):
    """utils.get_gaussian_weights :314-350 (centre = size/2.0, sigma = size*0.125, float64 math then float32)."""
    y_grid, x_grid = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    g = np.exp(-((y_grid - height / 2.0) ** 2 / (2 * (height * sigma_scale) ** 2)
                 + (x_grid - width / 2.0) ** 2 / (2 * (width * sigma_scale) ** 2)))
    return np.repeat(g[:, :, np.newaxis], n_channels, axis=2).astype(np.float32)


def tile_grid(h: int, w: int, patch_size, patch_overlap: int):
    """Tile origins of run_model_inference :383-392."""
    if patch_size:
        patch_size = min(patch_size, max(h, w))
        stride = max(patch_size - patch_overlap, 1)
        h_idx = list(range(0, h - patch_size, stride)) + [max(h - patch_size, 0)]
        w_idx = list(range(0, w - patch_size, stride)) + [max(w - patch_size, 0)]
    else:
        patch_size = max(h, w)
        h_idx, w_idx = [0], [0]
    return h_idx, w_idx, patch_size


def add_gaussian_noise(img: np.ndarray, sigma=15):
    """utils.add_gaussian_noise :29-36: global numpy generator reseeded with 0, float64 noise added in place to the
    float32 patch (numpy computes the sum in double and rounds to float32), clip to [0, 1]."""
    if img.dtype != np.float32 and img.dtype != np.float64:
        img = img.astype(np.float32) / 255.
    rs = np.random.RandomState(0)            # == np.random.seed(0) followed by np.random.normal, without the side effect
    img += rs.normal(0, sigma / 255., img.shape)
    img = np.clip(img, 0, 1)
    return img.astype(np.float32)


def run_model_inference(model, input_img: np.ndarray, patch_size=None, patch_overlap: int = 32, use_pad: bool = True,
                        need_degradation=False, noise_level=None):
    """run_model_inference :353-454 (returns the restored image only)."""
    with torch.no_grad():
        img = normalize(input_img)
        h, w = img.shape[:2]
        h_idx, w_idx, ps = tile_grid(h, w, patch_size, patch_overlap)
        c = min(3, img.shape[2])
        out = np.zeros((h, w, c), dtype=np.float32)
        wmap = np.zeros((h, w, c), dtype=np.float32)
        window = gaussian_weights(ps, ps, c)
        for hi in h_idx:
            for wi in w_idx:
                patch = img[hi:hi + ps, wi:wi + ps, :].copy()
                if need_degradation and noise_level is not None:
                    patch = add_gaussian_noise(patch, noise_level)
                x = torch.from_numpy(patch.transpose(2, 0, 1)).unsqueeze(0)
                if use_pad:
                    hp, wp = x.shape[-2:]
                    y = model(pad(x))[:, :, :hp, :wp]
                else:
                    y = model(x)
                pred = y.squeeze(0).cpu().numpy().transpose(1, 2, 0)
                ch, cw = pred.shape[:2]
                cur = window[:ch, :cw]
                out[hi:hi + ch, wi:wi + cw, :] += pred * cur
                wmap[hi:hi + ch, wi:wi + cw, :] += cur
        out /= np.maximum(wmap, 1e-8)
        if input_img.dtype == np.uint16:
            return np.clip(out * 65535.0, 0, 65535).round().astype(np.uint16)
        if input_img.dtype == np.uint8:
            return np.clip(out * 255.0, 0, 255).round().astype(np.uint8)
        return np.clip(out * np.max(input_img), np.min(input_img), np.max(input_img)).astype(input_img.dtype)
