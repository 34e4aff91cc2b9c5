#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): one TransformerBlock at each
level width (C = 48 / 96 / 192 / 384, both tensor-core modes) through ir_block_forward, a tiny whole Restormer (gray and
dual-pixel), and a DnCNN, every result checked against the CPU oracle so that a sanitizer-clean run is also a correct one.

    compute-sanitizer --tool racecheck python scripts/sanitize_target.py [blocks|model|all]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import image_restoration_models_b200 as M  # noqa: E402
import oracle  # noqa: E402

torch.set_grad_enabled(False)


def blocks():
    from test_gpu_parity import run_block
    from conftest import golden_names, load_golden
    for name in golden_names("block"):
        meta, z = load_golden(name)
        wb = meta["LayerNorm_type"] != "BiasFree"
        sd = oracle.synth_state_dict(oracle.synth._block_schema("blk", meta["C"], meta["heads"], 2.66, False, wb),
                                     meta["wseed"])
        x = oracle.synth_tensor(meta["shape"], meta["xseed"], -1.0, 1.0)
        for mode in (0, 1, 3):
            y = run_block(meta, sd, x, mode)
            err = float(np.abs(y.astype(np.float64) - z["y64"]).max())
            print(f"block {name} mode {mode}: max-abs {err:.2e}", flush=True)
            assert err <= 1e-3, (name, mode, err)


def model():
    for task, shape in (("gray_denoise", (1, 1, 32, 48)), ("defocus_dual", (1, 6, 32, 32))):
        kw = oracle.RESTORMER_TASKS[task]
        sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 5)
        x = oracle.synth_image(shape, 6, 25.0)
        y_ref = oracle.restormer_forward(sd, x).numpy()
        for mode in ("fp32", "half"):
            m = M.Restormer(**kw, bias=False).eval()
            m.load_state_dict(sd, strict=True)
            y = m.cuda().set_mode(mode)(x.cuda()).cpu().numpy()
            err = float(np.abs(y - y_ref).max())
            print(f"restormer {task} {mode}: max-abs {err:.2e}", flush=True)
            assert err <= 1e-3, (task, mode, err)
    dsd = oracle.synth_state_dict(oracle.dncnn_schema(1, 1, 64, 17, "R"), 8)
    dx = oracle.synth_image((1, 1, 40, 136), 9, 25.0)
    d = M.DnCNN(1, 1, 64, 17, "R").eval()
    d.load_state_dict(dsd, strict=True)
    derr = float(np.abs(d.cuda()(dx.cuda()).cpu().numpy() - oracle.dncnn_forward(dsd, dx).numpy()).max())
    print(f"dncnn: max-abs {derr:.2e}", flush=True)
    assert derr <= 1e-3


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("blocks", "all"):
        blocks()
    if what in ("model", "all"):
        model()
    print("sanitize target OK")
