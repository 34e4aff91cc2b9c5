"""CPU: the oracle restatement reproduces the outputs of the unmodified reference (tests/golden/*.npz,
written by oracle/make_golden.py in the build container) and its schemas match the live state_dict."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN, golden_names, load_golden

torch.set_grad_enabled(False)


@pytest.mark.parametrize("name", golden_names("restormer"))
def test_restormer_oracle_matches_reference(name):
    meta, z = load_golden(name)
    kw = oracle.RESTORMER_TASKS[meta["task"]]
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), meta["wseed"])
    x = oracle.synth_image(meta["shape"], meta["xseed"], meta["sigma"])
    if np.prod(meta["shape"]) > 3 * 64 * 64:   # keep the CPU suite short: big cases in fp32 only
        y = oracle.restormer_forward(sd, x).numpy()
        assert np.abs(y - z["y"]).max() < 2e-5
        return
    taps = {}
    y = oracle.restormer_forward(sd, x, taps).numpy()
    # fp32 vs fp32: only summation-order noise of the same ATen kernels
    assert np.abs(y - z["y"]).max() < 2e-5
    y64 = oracle.restormer_forward({k: v.double() for k, v in sd.items()}, x.double()).numpy()
    assert np.abs(y64 - z["y64"]).max() < 1e-9
    for k, v in taps.items():
        assert np.abs(v.numpy()[:, :, ::4, ::4] - z["tap_" + k]).max() < 5e-5, k


@pytest.mark.parametrize("name", golden_names("dncnn"))
def test_dncnn_oracle_matches_reference(name):
    meta, z = load_golden(name)
    sd = oracle.synth_state_dict(
        oracle.dncnn_schema(meta["in_nc"], meta["in_nc"], 64, meta["nb"], meta["act_mode"]), meta["wseed"])
    x = oracle.synth_image(meta["shape"], meta["xseed"], meta["sigma"])
    y = oracle.dncnn_forward(sd, x).numpy()
    assert np.abs(y - z["y"]).max() < 1e-5
    y64 = oracle.dncnn_forward({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()},
                               x.double()).numpy()
    assert np.abs(y64 - z["y64"]).max() < 1e-9


@pytest.mark.parametrize("name", golden_names("block"))
def test_block_oracle_matches_reference(name):
    meta, z = load_golden(name)
    wb = meta["LayerNorm_type"] != "BiasFree"
    sd = oracle.synth_state_dict(oracle.synth._block_schema("blk", meta["C"], meta["heads"], 2.66, False, wb),
                                 meta["wseed"])
    x = oracle.synth_tensor(meta["shape"], meta["xseed"], -1.0, 1.0)
    y = oracle.transformer_block(sd, "blk", x).numpy()
    assert np.abs(y - z["y"]).max() < 2e-5
    a = oracle.attention(sd, "blk.attn", oracle.layer_norm(sd, "blk.norm1", x)).numpy()
    assert np.abs(a - z["attn"]).max() < 2e-5


def test_schemas_match_live_reference_record():
    rec = json.load(open(os.path.join(GOLDEN, "schemas.json")))
    for task, kw in oracle.RESTORMER_TASKS.items():
        if task in rec:
            mine = [[k, list(s)] for k, s, _ in oracle.restormer_schema(**kw)]
            assert mine == rec[task], task
    for key, want in rec.items():
        if key.startswith("dncnn_"):
            _, n, nb, act = key.split("_")
            mine = [[k, list(s)] for k, s, _ in oracle.dncnn_schema(int(n), int(n), 64, int(nb), act)]
            assert mine == want, key


def test_synth_is_platform_independent():
    # pinned values of the integer hash: any drift would silently invalidate every golden
    u = oracle.synth.synth_uniform((4,), 7)
    assert u.dtype == np.float32
    ref = oracle.synth.synth_uniform((4,), 7)
    assert (u == ref).all()
    assert 0.0 <= float(u.min()) and float(u.max()) < 1.0
    t = oracle.synth_tensor((2, 3), 123, -1.0, 1.0)
    assert abs(float(t.sum()) - float(oracle.synth_tensor((2, 3), 123, -1.0, 1.0).sum())) == 0.0
    pinned = np.load(os.path.join(GOLDEN, "synth_pin.npy"))
    assert (oracle.synth.synth_uniform((16,), 99) == pinned).all()
