"""Golden outputs of the UNMODIFIED reference harness run_model_inference (TEST INFRASTRUCTURE).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_tiling

Imports /root/reference/src/utils.py with stub modules for the packages that are absent here (skimage, deblurganv2,
mair: none is on the Restormer / DnCNN path) and runs it on synthetic images with a deterministic stand-in model
(element-wise, position dependent, bit-identical on CPU and GPU), so the fixtures pin tile grid, reflect pad,
Gaussian window, blend order and output rounding."""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

from . import synth

REF_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, dtype, H, W, C, patch_size, overlap, use_pad, seed)
CASES = [
    ("tiling_u8_color_70x90_p32", "uint8", 70, 90, 3, 32, 8, True, 1),
    ("tiling_u8_gray_50x37_p24", "uint8", 50, 37, 1, 24, 6, True, 2),
    ("tiling_u16_color_45x64_p40", "uint16", 45, 64, 3, 40, 12, True, 3),
    ("tiling_u8_color_33x47_nopad", "uint8", 33, 47, 3, 20, 4, False, 4),
    ("tiling_f32_color_40x40_p64", "float32", 40, 40, 3, 64, 16, True, 5),
]
# synthetic degradation (need_degradation=True, noise_level=sigma): (name, dtype, H, W, C, patch, overlap, use_pad, seed, sigma)
NOISE_CASES = [
    ("tiling_noise25_u8_color_70x90_p32", "uint8", 70, 90, 3, 32, 8, True, 6, 25),
    ("tiling_noise15_u8_gray_50x37_p24", "uint8", 50, 37, 1, 24, 6, True, 7, 15),
    ("tiling_noise50_u8_gray_33x47_nopad", "uint8", 33, 47, 1, 20, 4, False, 8, 50),
]


class StandIn(torch.nn.Module):
    """y = x * mask[:, :, :H, :W] + 0.05 — element-wise (identical IEEE results on CPU and GPU) and position dependent
    inside the tile, so overlapping tiles disagree and the window weights matter."""

    def __init__(self, size=128, seed=77):
        super().__init__()
        self.register_buffer("mask", synth.synth_tensor((1, 1, size, size), seed, 0.5, 1.0))

    def forward(self, x):
        return x * self.mask[:, :, : x.shape[2], : x.shape[3]] + 0.05


def make_image(dtype, h, w, c, seed):
    u = synth.synth_uniform((h, w, c), 500 + seed)
    if dtype == "uint8":
        return (u * 255.999).astype(np.uint8)
    if dtype == "uint16":
        return (u * 65535.999).astype(np.uint16)
    return (u * 3.0).astype(np.float32)          # max > 1 -> normalised by its max (utils.normalize :165-169)


def import_reference_utils():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_SRC)
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    dummy = type("Dummy", (), {})
    stub("skimage"); stub("skimage.metrics", peak_signal_noise_ratio=None, structural_similarity=None)
    stub("deblurganv2", normalize=None, pad=None, postprocess=None, get_model=None)
    stub("deblurganv2.models"); stub("deblurganv2.models.fpn_inception", FPNInception=dummy)
    stub("deblurganv2.models.fpn_mobilenet", FPNMobileNet=dummy)
    stub("mair", get_model=None); stub("mair.basicsr"); stub("mair.basicsr.archs")
    stub("mair.basicsr.archs.mair_arch", MaIR=dummy)
    stub("mair.realDenoising"); stub("mair.realDenoising.basicsr"); stub("mair.realDenoising.basicsr.models")
    stub("mair.realDenoising.basicsr.models.archs"); stub("mair.realDenoising.basicsr.models.archs.mairunet_arch",
                                                           MaIRUNet=dummy)
    stub("rednet", get_model=None)
    import utils  # noqa: E402
    return utils


def main():
    utils = import_reference_utils()
    model = StandIn().eval()
    for name, dtype, h, w, c, ps, ov, use_pad, seed in CASES:
        img = make_image(dtype, h, w, c, seed)
        out, _ = utils.run_model_inference(model, img, torch.device("cpu"), patch_size=ps, patch_overlap=ov,
                                           pad=utils.pad if use_pad else None)
        meta = dict(kind="tiling", dtype=dtype, shape=[h, w, c], patch_size=ps, patch_overlap=ov, use_pad=use_pad,
                    seed=seed)
        np.savez(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), out=out)
        print(name, out.dtype, out.shape, int(out.astype(np.float64).sum()))
    for name, dtype, h, w, c, ps, ov, use_pad, seed, sigma in NOISE_CASES:
        img = make_image(dtype, h, w, c, seed)
        out, _ = utils.run_model_inference(model, img, torch.device("cpu"), patch_size=ps, patch_overlap=ov,
                                           need_degradation=True, noise_level=sigma,
                                           pad=utils.pad if use_pad else None)
        meta = dict(kind="tiling_noise", dtype=dtype, shape=[h, w, c], patch_size=ps, patch_overlap=ov, use_pad=use_pad,
                    seed=seed, sigma=sigma)
        np.savez(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), out=out)
        print(name, out.dtype, out.shape, int(out.astype(np.float64).sum()))
    np.save(os.path.join(OUT, "gaussian_window_24.npy"), utils.get_gaussian_weights(24, 24, 1))


if __name__ == "__main__":
    main()
