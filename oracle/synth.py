"""Deterministic synthetic weights and images (TEST INFRASTRUCTURE, see oracle/__init__.py).

Weights and inputs are produced by a pure-integer counter hash so that the
golden generator (which runs the unmodified reference in the build container),
the CPU oracle and the GPU tests all see bit-identical tensors without
shipping 100 MB checkpoints.  Nothing here depends on torch's RNG stream.

The parameter schemas restate the reference's module tree
(src/restormer/restormer.py:194-243, src/dncnn/models/network_dncnn.py:41-67,
src/dncnn/models/basicblock.py:61-98); ``tests/test_oracle_golden.py`` checks
names and shapes against the list recorded from the live reference.
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch

_M32 = np.uint64(0xFFFFFFFF)


def _mix(idx: np.ndarray, seed: int) -> np.ndarray:
    """murmur3-style finaliser over a uint64 counter; returns uint32 values in a uint64 array."""
    x = (idx + np.uint64((seed * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF)) & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & _M32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & _M32
    x ^= x >> np.uint64(16)
    return x


def synth_uniform(shape, seed: int) -> np.ndarray:
    """float32 uniform in [0,1) with 24 random bits, bit-exact on every platform."""
    n = int(np.prod(shape)) if len(shape) else 1
    idx = np.arange(n, dtype=np.uint64)
    u = (_mix(idx, seed) >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / (1 << 24))
    return u.reshape(shape)


def synth_tensor(shape, seed: int, lo: float, hi: float) -> torch.Tensor:
    u = synth_uniform(tuple(shape), seed)
    return torch.from_numpy(np.float32(lo) + u * np.float32(hi - lo))


def synth_image(shape, seed: int, sigma: float | None = None) -> torch.Tensor:
    """[0,1] image; optional additive noise of std sigma/255 then clip (shape of src/utils.py:29-36)."""
    img = synth_uniform(tuple(shape), seed)
    # low-pass a little so that neighbouring pixels correlate like a natural image
    if sigma is not None:
        u1 = np.maximum(synth_uniform(tuple(shape), seed + 101), np.float32(1e-7))
        u2 = synth_uniform(tuple(shape), seed + 202)
        gauss = np.sqrt(-2.0 * np.log(u1.astype(np.float64))) * np.cos(2.0 * math.pi * u2.astype(np.float64))
        img = np.clip(img.astype(np.float64) + gauss * (sigma / 255.0), 0.0, 1.0).astype(np.float32)
    return torch.from_numpy(img)


# --------------------------------------------------------------------------------------
# parameter schemas
# --------------------------------------------------------------------------------------

RESTORMER_TASKS = {
    # src/restormer/options/*.yml network_g blocks
    "gray_denoise": dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", dual_pixel_task=False),
    "color_denoise": dict(inp_channels=3, out_channels=3, LayerNorm_type="BiasFree", dual_pixel_task=False),
    "real_denoise": dict(inp_channels=3, out_channels=3, LayerNorm_type="BiasFree", dual_pixel_task=False),
    "motion_deblur": dict(inp_channels=3, out_channels=3, LayerNorm_type="WithBias", dual_pixel_task=False),
    "defocus_single": dict(inp_channels=3, out_channels=3, LayerNorm_type="WithBias", dual_pixel_task=False),
    "defocus_dual": dict(inp_channels=6, out_channels=3, LayerNorm_type="WithBias", dual_pixel_task=True),
}


def _block_schema(prefix, C, heads, ffn, bias, with_bias_ln):
    h = int(C * ffn)
    out = [(f"{prefix}.norm1.body.weight", (C,), "ln_w")]
    if with_bias_ln:
        out.append((f"{prefix}.norm1.body.bias", (C,), "ln_b"))
    out.append((f"{prefix}.attn.temperature", (heads, 1, 1), "temp"))
    for nm, shp in (("qkv", (3 * C, C, 1, 1)), ("qkv_dwconv", (3 * C, 1, 3, 3)), ("project_out", (C, C, 1, 1))):
        out.append((f"{prefix}.attn.{nm}.weight", shp, "conv"))
        if bias:
            out.append((f"{prefix}.attn.{nm}.bias", (shp[0],), "bias"))
    out.append((f"{prefix}.norm2.body.weight", (C,), "ln_w"))
    if with_bias_ln:
        out.append((f"{prefix}.norm2.body.bias", (C,), "ln_b"))
    for nm, shp in (("project_in", (2 * h, C, 1, 1)), ("dwconv", (2 * h, 1, 3, 3)), ("project_out", (C, h, 1, 1))):
        out.append((f"{prefix}.ffn.{nm}.weight", shp, "conv"))
        if bias:
            out.append((f"{prefix}.ffn.{nm}.bias", (shp[0],), "bias"))
    return out


def restormer_schema(inp_channels=3, out_channels=3, dim=48, num_blocks=(4, 6, 6, 8), num_refinement_blocks=4,
                     heads=(1, 2, 4, 8), ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias",
                     dual_pixel_task=False):
    """Ordered (name, shape, kind) list == reference ``Restormer(...).state_dict()`` order.

    Order follows attribute registration in src/restormer/restormer.py:207-243.
    """
    wb = LayerNorm_type != "BiasFree"
    d = dim
    # OverlapPatchEmbed is built with its own default bias=False (restormer.py:158,207)
    out = [("patch_embed.proj.weight", (d, inp_channels, 3, 3), "conv")]

    def stage(name, C, hd, n):
        r = []
        for i in range(n):
            r += _block_schema(f"{name}.{i}", C, hd, ffn_expansion_factor, bias, wb)
        return r

    def conv(name, co, ci, k, b=bias):
        r = [(f"{name}.weight", (co, ci, k, k), "conv")]
        if b:
            r.append((f"{name}.bias", (co,), "bias"))
        return r

    out += stage("encoder_level1", d, heads[0], num_blocks[0])
    out += conv("down1_2.body.0", d // 2, d, 3, False)
    out += stage("encoder_level2", 2 * d, heads[1], num_blocks[1])
    out += conv("down2_3.body.0", d, 2 * d, 3, False)
    out += stage("encoder_level3", 4 * d, heads[2], num_blocks[2])
    out += conv("down3_4.body.0", 2 * d, 4 * d, 3, False)
    out += stage("latent", 8 * d, heads[3], num_blocks[3])
    out += conv("up4_3.body.0", 16 * d, 8 * d, 3, False)
    out += conv("reduce_chan_level3", 4 * d, 8 * d, 1)
    out += stage("decoder_level3", 4 * d, heads[2], num_blocks[2])
    out += conv("up3_2.body.0", 8 * d, 4 * d, 3, False)
    out += conv("reduce_chan_level2", 2 * d, 4 * d, 1)
    out += stage("decoder_level2", 2 * d, heads[1], num_blocks[1])
    out += conv("up2_1.body.0", 4 * d, 2 * d, 3, False)
    out += stage("decoder_level1", 2 * d, heads[0], num_blocks[0])
    out += stage("refinement", 2 * d, heads[0], num_refinement_blocks)
    if dual_pixel_task:
        out += conv("skip_conv", 2 * d, d, 1)
    out += conv("output", out_channels, 2 * d, 3)
    return out


def dncnn_schema(in_nc=1, out_nc=1, nc=64, nb=17, act_mode="BR"):
    """Ordered (name, shape, kind) list == reference ``DnCNN(...).state_dict()`` order.

    ``B.sequential`` flattens the per-layer Sequentials (basicblock.py:15-35), so the
    indices count conv, [BN], ReLU modules in a row (network_dncnn.py:63-67).
    """
    out = []
    idx = 0
    has_bn = "B" in act_mode
    for layer in range(nb):
        ci = in_nc if layer == 0 else nc
        co = out_nc if layer == nb - 1 else nc
        out.append((f"model.{idx}.weight", (co, ci, 3, 3), "conv"))
        out.append((f"model.{idx}.bias", (co,), "bias"))
        idx += 1
        if layer == nb - 1:
            break
        if has_bn and layer > 0:
            out.append((f"model.{idx}.weight", (nc,), "bn_w"))
            out.append((f"model.{idx}.bias", (nc,), "bn_b"))
            out.append((f"model.{idx}.running_mean", (nc,), "bn_mean"))
            out.append((f"model.{idx}.running_var", (nc,), "bn_var"))
            out.append((f"model.{idx}.num_batches_tracked", (), "bn_count"))
            idx += 1
        idx += 1  # activation
    return out


def synth_state_dict(schema, seed: int = 0, gain: float = 1.0 / math.sqrt(3.0)):
    """Deterministic weights: conv ~ U(-b,b) with b = gain*sqrt(3/fan_in); the default gain gives
    b = 1/sqrt(fan_in), the distribution of torch's default Conv2d init (what "random-init" means in
    BASELINE.json), so outputs stay on the [0,1] image scale the 1e-3 tolerance is stated on;
    LN weight ~ U(0.5,1.5), LN bias ~ U(-0.1,0.1), temperature ~ U(0.5,2) (SURVEY.md §8d recipe:
    the reference's all-ones/zeros init would hide per-head and LN-affine indexing bugs)."""
    sd = {}
    for name, shape, kind in schema:
        s = (zlib.crc32(name.encode()) ^ (seed * 7919)) & 0x7FFFFFFF
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            b = gain * math.sqrt(3.0) / math.sqrt(fan_in)
            t = synth_tensor(shape, s, -b, b)
        elif kind == "bias":
            t = synth_tensor(shape, s, -0.05, 0.05)
        elif kind in ("ln_w", "bn_w"):
            t = synth_tensor(shape, s, 0.5, 1.5)
        elif kind in ("ln_b", "bn_b", "bn_mean"):
            t = synth_tensor(shape, s, -0.1, 0.1)
        elif kind == "bn_var":
            t = synth_tensor(shape, s, 0.5, 1.5)
        elif kind == "bn_count":
            t = torch.tensor(100, dtype=torch.long)
        elif kind == "temp":
            t = synth_tensor(shape, s, 0.5, 2.0)
        else:
            raise ValueError(kind)
        sd[name] = t
    return sd
