"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

It imports ``restormer.restormer`` and ``dncnn.models.network_dncnn`` from
/root/reference/src (read-only), loads the deterministic synthetic weights of
oracle/synth.py with ``strict=True`` and stores inputs' seeds + outputs.  The GPU box
has no /root/reference; tests there only read the committed .npz files.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import synth

REF_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (case name, task, batch, H, W, weight seed, image seed, noise sigma or None)
RESTORMER_CASES = [
    ("restormer_gray_64", "gray_denoise", 1, 64, 64, 1, 11, 25.0),
    ("restormer_color_b2_32x48", "color_denoise", 2, 32, 48, 2, 12, 15.0),
    ("restormer_real_128", "real_denoise", 1, 128, 128, 3, 13, None),
    ("restormer_motion_64", "motion_deblur", 1, 64, 64, 4, 14, None),
    ("restormer_dual_32x40", "defocus_dual", 1, 32, 40, 5, 15, None),
    ("restormer_gray_8", "gray_denoise", 1, 8, 8, 6, 16, None),
]
# (case name, in_nc, nb, act_mode, batch, H, W, weight seed, image seed, sigma)
DNCNN_CASES = [
    ("dncnn_s_R_64", 1, 17, "R", 1, 64, 64, 21, 31, 25.0),
    ("dncnn_s_BR_64", 1, 17, "BR", 1, 64, 64, 22, 32, 25.0),
    ("dncnn_gray_blind_R_b2_33x47", 1, 20, "R", 2, 33, 47, 23, 33, 50.0),
    ("dncnn_color_blind_R_40x56", 3, 20, "R", 1, 40, 56, 24, 34, 15.0),
]
# single TransformerBlock cases: (name, C, heads, LN type, batch, h, w, weight seed, x seed)
BLOCK_CASES = [
    ("block_c48_h1_biasfree", 48, 1, "BiasFree", 2, 16, 24, 41, 51),
    ("block_c96_h2_withbias", 96, 2, "WithBias", 1, 16, 16, 42, 52),
    ("block_c96_h1_biasfree", 96, 1, "BiasFree", 1, 24, 16, 43, 53),
    ("block_c192_h4_withbias", 192, 4, "WithBias", 1, 8, 16, 44, 54),
    ("block_c384_h8_biasfree", 384, 8, "BiasFree", 2, 8, 8, 45, 55),
]


def _import_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_SRC)
    from restormer.restormer import Restormer, TransformerBlock  # noqa: E402
    from dncnn.models.network_dncnn import DnCNN  # noqa: E402
    return Restormer, TransformerBlock, DnCNN


def main():
    torch.set_grad_enabled(False)
    Restormer, TransformerBlock, DnCNN = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    schemas = {}

    for name, task, b, h, w, ws, xs, sigma in RESTORMER_CASES:
        kw = synth.RESTORMER_TASKS[task]
        model = Restormer(**kw, bias=False).eval()
        sd = synth.synth_state_dict(synth.restormer_schema(**kw), ws)
        ref_sd = model.state_dict()
        assert list(ref_sd.keys()) == list(sd.keys()), "schema order differs from the live reference"
        model.load_state_dict(sd, strict=True)
        schemas[task] = [[k, list(v.shape)] for k, v in ref_sd.items()]
        x = synth.synth_image((b, kw["inp_channels"], h, w), xs, sigma)
        taps = {}
        hooks = []
        for stage in ("patch_embed", "encoder_level1", "encoder_level2", "encoder_level3", "latent",
                      "decoder_level3", "decoder_level2", "decoder_level1", "refinement"):
            hooks.append(getattr(model, stage).register_forward_hook(
                lambda m, i, o, s=stage: taps.__setitem__(s, o.detach().clone())))
        y = model(x)
        y64 = model.double()(x.double())
        for hk in hooks:
            hk.remove()
        arrays = {"y": y.numpy(), "y64": y64.numpy()}
        for k, v in taps.items():
            arrays["tap_" + k] = v.float().numpy()[:, :, ::4, ::4].copy()
        meta = dict(kind="restormer", task=task, shape=[b, kw["inp_channels"], h, w], wseed=ws, xseed=xs, sigma=sigma)
        np.savez(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **arrays)
        print(name, "max|y|", float(y.abs().max()), "fp32-vs-fp64", float((y.double() - y64).abs().max()))

    for name, in_nc, nb, act, b, h, w, ws, xs, sigma in DNCNN_CASES:
        model = DnCNN(in_nc=in_nc, out_nc=in_nc, nc=64, nb=nb, act_mode=act).eval()
        sd = synth.synth_state_dict(synth.dncnn_schema(in_nc, in_nc, 64, nb, act), ws)
        ref_sd = model.state_dict()
        assert list(ref_sd.keys()) == list(sd.keys()), "schema order differs from the live reference"
        model.load_state_dict(sd, strict=True)
        schemas[f"dncnn_{in_nc}_{nb}_{act}"] = [[k, list(v.shape)] for k, v in ref_sd.items()]
        x = synth.synth_image((b, in_nc, h, w), xs, sigma)
        y = model(x)
        y64 = model.double()(x.double())
        meta = dict(kind="dncnn", in_nc=in_nc, nb=nb, act_mode=act, shape=[b, in_nc, h, w], wseed=ws, xseed=xs, sigma=sigma)
        np.savez(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), y=y.numpy(), y64=y64.numpy())
        print(name, "max|y|", float(y.abs().max()), "fp32-vs-fp64", float((y.double() - y64).abs().max()))

    for name, C, heads, ln, b, h, w, ws, xs in BLOCK_CASES:
        blk = TransformerBlock(dim=C, num_heads=heads, ffn_expansion_factor=2.66, bias=False, LayerNorm_type=ln).eval()
        schema = synth._block_schema("blk", C, heads, 2.66, False, ln != "BiasFree")
        sd = synth.synth_state_dict(schema, ws)
        blk.load_state_dict({k[len("blk."):]: v for k, v in sd.items()}, strict=True)
        x = synth.synth_tensor((b, C, h, w), xs, -1.0, 1.0)
        y_attn = blk.attn(blk.norm1(x))
        y = blk(x)
        y64 = blk.double()(x.double())
        meta = dict(kind="block", C=C, heads=heads, LayerNorm_type=ln, shape=[b, C, h, w], wseed=ws, xseed=xs)
        np.savez(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), y=y.numpy(), y64=y64.numpy(),
                 attn=y_attn.numpy())
        print(name, "max|y|", float(y.abs().max()))

    with open(os.path.join(OUT, "schemas.json"), "w") as f:
        json.dump(schemas, f)


if __name__ == "__main__":
    main()
