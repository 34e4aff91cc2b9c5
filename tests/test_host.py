"""CPU: boundary contract of the mirror classes and the C-ABI library (no compute without a GPU)."""
import ctypes as C
import json
import os
import re
import subprocess

import pytest
import torch

import image_restoration_models_b200 as M
from image_restoration_models_b200 import _native
import oracle
from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    def parse(name):
        hdr = open(os.path.join(ROOT, "include", name)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
        return set(re.findall(r"\b(ir_[a-z0-9_]+)\s*\(", hdr))
    product, testing = parse("irb200.h"), parse("irb200_testing.h")
    assert product and testing, "no declarations parsed"
    # the test hooks and the hardware probe live outside the product ABI
    assert testing == {"ir_test_conv1x1", "ir_test_conv3x3", "ir_probe_shifted_descriptor"} and not (product & testing)
    declared = product | testing
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = _native.lib()                       # raises if any symbol is missing
    for name in declared:
        assert hasattr(lib, name)
    assert lib.ir_abi_version() == _native.ABI_VERSION
    nm = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ir_[a-z0-9_]+)", nm))
    assert declared <= exported


@pytest.mark.parametrize("task", list(oracle.RESTORMER_TASKS))
def test_restormer_state_dict_contract(task):
    kw = oracle.RESTORMER_TASKS[task]
    m = M.Restormer(**kw, bias=False)
    want = [(k, tuple(s)) for k, s, _ in oracle.restormer_schema(**kw)]
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == want
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 1)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])
    # the native plan agrees with the module tree, tensor by tensor
    lib = _native.lib()
    assert lib.ir_restormer_param_count(C.byref(m._cfg)) == len(got)
    for i, (_, shp) in enumerate(got):
        n = 1
        for d in shp:
            n *= d
        assert lib.ir_restormer_param_numel(C.byref(m._cfg), i) == n
    assert lib.ir_restormer_packed_bytes(C.byref(m._cfg), 0) >= 4 * sum(p.numel() for p in m.parameters())


def test_restormer_bias_and_custom_widths_schema():
    kw = dict(inp_channels=3, out_channels=3, dim=32, num_blocks=[1, 2, 2, 3], num_refinement_blocks=2,
              heads=[1, 2, 2, 4], ffn_expansion_factor=2.0, bias=True, LayerNorm_type="WithBias")
    m = M.Restormer(**kw)
    want = [(k, tuple(s)) for k, s, _ in oracle.restormer_schema(**kw)]
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == want
    assert _native.lib().ir_restormer_param_count(C.byref(m._cfg)) == len(want)


@pytest.mark.parametrize("n,nb,act", [(1, 17, "R"), (1, 17, "BR"), (1, 20, "R"), (3, 20, "R")])
def test_dncnn_state_dict_contract(n, nb, act):
    rec = json.load(open(os.path.join(GOLDEN, "schemas.json")))[f"dncnn_{n}_{nb}_{act}"]
    m = M.DnCNN(in_nc=n, out_nc=n, nc=64, nb=nb, act_mode=act)
    got = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    assert got == rec
    sd = oracle.synth_state_dict(oracle.dncnn_schema(n, n, 64, nb, act), 3)
    m.load_state_dict(sd, strict=True)
    lib = _native.lib()
    assert lib.ir_dncnn_param_count(C.byref(m._cfg)) == len(got)
    for i, (_, shp) in enumerate(got):
        numel = 1
        for d in shp:
            numel *= d
        assert lib.ir_dncnn_param_numel(C.byref(m._cfg), i) == numel


def test_param_counts_match_survey():
    count = lambda m: sum(p.numel() for p in m.parameters())
    assert count(M.Restormer(1, 1, LayerNorm_type="BiasFree")) == 26109076
    assert count(M.Restormer(3, 3, LayerNorm_type="BiasFree")) == 26111668
    assert count(M.Restormer(3, 3, LayerNorm_type="WithBias")) == 26126644
    assert count(M.Restormer(6, 3, LayerNorm_type="WithBias", dual_pixel_task=True)) == 26132548
    assert count(M.DnCNN(1, 1, 64, 17, "R")) == 555137
    assert count(M.DnCNN(1, 1, 64, 17, "BR")) == 557057


def test_no_cpu_fallback_and_argument_errors():
    m = M.Restormer(3, 3).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64))
    d = M.DnCNN(1, 1, 64, 17, "R").eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d(torch.zeros(1, 1, 32, 32))
    with pytest.raises(NotImplementedError):
        M.DnCNN(act_mode="L")
    lib = _native.lib()
    bad = _native.IrRestormerCfg(3, 3, 20, (C.c_int32 * 4)(1, 1, 1, 1), 1, (C.c_int32 * 4)(1, 2, 4, 8), 2.66, 0, 1, 0)
    assert lib.ir_restormer_param_count(C.byref(bad)) == -1
    assert "invalid argument" in _native.last_error()
    with pytest.raises(ValueError):
        _native.check(_native.IR_ERR_INVALID)
    # non-dual model with inp != out cannot add the image residual (restormer.py:281)
    bad2 = _native.IrRestormerCfg(6, 3, 48, (C.c_int32 * 4)(4, 6, 6, 8), 4, (C.c_int32 * 4)(1, 2, 4, 8), 2.66, 0, 1, 0)
    assert lib.ir_restormer_param_count(C.byref(bad2)) == -1


def test_hidden_width_rounds_like_python():
    """int(dim * ffn_expansion_factor) in double precision (restormer.py:80): 80 * 2.1 = 168.00000000000003 -> 168, while
    the factor as a C float (2.0999999) gives 167.  The factor crosses the ABI as a double."""
    kw = dict(inp_channels=3, out_channels=3, dim=80, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1,
              heads=[5, 10, 10, 20], ffn_expansion_factor=2.1, bias=False, LayerNorm_type="WithBias")
    assert int(80 * 2.1) == 168 and int(80 * float(torch.tensor(2.1, dtype=torch.float32))) == 167
    m = M.Restormer(**kw)
    lib = _native.lib()
    tensors = list(m.state_dict().values())
    assert lib.ir_restormer_param_count(C.byref(m._cfg)) == len(tensors)
    for i, t in enumerate(tensors):
        assert lib.ir_restormer_param_numel(C.byref(m._cfg), i) == t.numel(), i


def test_fp16_range_guard_selects_strict_mode_from_the_weights():
    kw = oracle.RESTORMER_TASKS["gray_denoise"]
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 5)
    m = M.Restormer(**kw, bias=False).eval()
    m.load_state_dict(sd)
    b0 = m.fp16_range_bound()
    assert 0 < b0 < 0.25 * 65504 and m.resolved_mode() == "fp32"
    big = {k: (v * 50 if (".norm1." in k or ".norm2." in k) and k.endswith("weight") else v) for k, v in sd.items()}
    m.load_state_dict(big)
    assert m.fp16_range_bound() > 100 * b0 and m.resolved_mode() == "fp32_strict"
    assert m.set_range_guard(False).resolved_mode() == "fp32"
    assert m.set_range_guard(True).set_mode("half").resolved_mode() == "half"       # explicit modes are never overridden
    # the strict mode keeps the GDFN hidden tensor in HBM: larger workspace, more launches are fine, same parameters
    m.set_mode("fp32").load_state_dict(sd)
    fast = m.workspace_bytes(1, 256, 256)
    m.load_state_dict(big)
    assert m.workspace_bytes(1, 256, 256) > fast
    # conv biases disable every fused kernel: nothing is fp16 in fp32 mode
    assert M.Restormer(3, 3, bias=True).fp16_range_bound() == 0.0


def test_workspace_and_launch_queries():
    m = M.Restormer(3, 3)
    b1 = m.workspace_bytes(1, 256, 256)
    b4 = m.workspace_bytes(4, 256, 256)
    assert 0 < b1 < b4 <= 4 * b1 + (1 << 24)
    assert m.workspace_bytes(1, 512, 512) >= 3 * b1
    assert m.launches_per_forward() > 44 * 6
    assert M.DnCNN(1, 1, 64, 17, "R").launches_per_forward() == 17
    # the launch sequence DESIGN.md describes: 5 launches per block at C = 48 / 96 (norm1, fused MDTA front, fold,
    # attention output + norm2, fused GDFN), 8 at C = 192, 9 at C = 384, plus the 11 stand-alone convs / copies
    # (IRB_NORM1_CHAIN=1 would move norm1 into the previous block's GDFN epilogue: 4 per block + one pass per chain)
    g = M.Restormer(1, 1, LayerNorm_type="BiasFree")
    assert g.launches_per_forward() == (24 * 5 if not os.environ.get("IRB_NORM1_CHAIN") else 24 * 4 + 4) + 12 * 8 + 8 * 9 + 11
    # the GDFN hidden tensor never exists in HBM at the high-resolution levels: the workspace of the bench workload
    # is dominated by qkv (fp32 mode) and stays under 9 GB / 6 GB
    # neither does qkv there (fused front): the workspace of the bench workload stays under 4.5 GB in both modes
    assert g.workspace_bytes(8, 512, 512) < 4.5 * (1 << 30)
    assert g.set_mode("half").workspace_bytes(8, 512, 512) < 4.5 * (1 << 30)


def test_get_model_factories_roundtrip(tmp_path):
    # restormer.get_model: YAML network_g + checkpoint['params'] (src/restormer/__init__.py:8-20)
    from image_restoration_models_b200 import restormer as rpkg, dncnn as dpkg
    kw = dict(inp_channels=1, out_channels=1, dim=48, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1,
              heads=[1, 2, 4, 8], ffn_expansion_factor=2.66, bias=False, LayerNorm_type="BiasFree",
              dual_pixel_task=False)
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 9)
    ck = tmp_path / "net_g.pth"
    torch.save({"params": sd}, ck)
    yml = tmp_path / "opt.yml"
    yml.write_text("network_g:\n  type: Restormer\n" + "".join(f"  {k}: {v}\n" for k, v in kw.items())
                   + f"path:\n  pretrain_network_g: {ck}\n")
    m = rpkg.get_model(str(yml), torch.device("cpu"))
    assert not m.training and torch.equal(m.state_dict()["output.weight"], sd["output.weight"])
    dsd = oracle.synth_state_dict(oracle.dncnn_schema(1, 1, 64, 17, "R"), 4)
    dk = tmp_path / "dncnn_25.pth"
    torch.save(dsd, dk)
    dm = dpkg.get_model(str(dk), 1, 17, torch.device("cpu"))
    assert not dm.training and torch.equal(dm.state_dict()["model.0.weight"], dsd["model.0.weight"])
