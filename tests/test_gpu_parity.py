"""GPU parity: the CUDA path (through the C-ABI, via the mirror classes) against the committed golden
outputs of the unmodified reference and against the CPU oracle on fresh seeded inputs.

Tolerance (BASELINE.json north_star): fp32 mode max-abs <= 1e-3 on [0,1] images and |PSNR delta| <= 0.01 dB."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

import image_restoration_models_b200 as M
from image_restoration_models_b200 import _native
import oracle
from conftest import golden_names, load_golden, record

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

TOL_MAXABS = 1e-3
TOL_PSNR_DB = 0.01


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 10.0 * math.log10(1.0 / max(mse, 1e-30))


def check_parity(case, y, y_ref, clean):
    err = float(np.abs(y.astype(np.float64) - y_ref.astype(np.float64)).max())
    d_psnr = abs(psnr(y, clean) - psnr(y_ref, clean))
    record(case, max_abs=err, psnr_delta_db=d_psnr, psnr_vs_ref_db=psnr(y, y_ref))
    assert np.isfinite(y).all(), case
    assert err <= TOL_MAXABS, (case, err)
    assert d_psnr <= TOL_PSNR_DB, (case, d_psnr)


MODES = ["fp32", "half"]   # tf32 operands + fp32 intermediates / fp16 operands + fp16 intermediates: same parity bar


def build_restormer(kw, wseed, mode="fp32"):
    m = M.Restormer(**kw, bias=False).eval()
    m.load_state_dict(oracle.synth_state_dict(oracle.restormer_schema(**kw), wseed), strict=True)
    return m.cuda().set_mode(mode)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", golden_names("restormer"))
def test_restormer_vs_reference_golden(name, mode):
    meta, z = load_golden(name)
    kw = oracle.RESTORMER_TASKS[meta["task"]]
    m = build_restormer(kw, meta["wseed"], mode)
    name = f"{name}[{mode}]"
    x = oracle.synth_image(meta["shape"], meta["xseed"], meta["sigma"])
    clean = oracle.synth_image(meta["shape"], meta["xseed"], None).numpy()
    y = m(x.cuda()).cpu().numpy()
    # dual-pixel (6 -> 3 channels): score PSNR against the first view of the synthetic clean input
    check_parity(name, y, z["y64"], clean[:, : y.shape[1]])


# bf16 mode (BASELINE config 3 "fp32 and bf16"; north_star: "bf16 mode reported separately"): 8 mantissa bits are outside the
# 1e-3 bar by construction.  The numbers land in the parity report; the test pins that the mode runs, is deterministic, stays
# within what bf16 rounding of every intermediate can cause (<= 2e-2 max-abs, <= 0.05 dB) and differs from the fp16 mode.
BF16_MAXABS, BF16_PSNR_DB = 2e-2, 0.05


@pytest.mark.parametrize("name", golden_names("restormer"))
def test_restormer_bf16_mode_reported_separately(name):
    meta, z = load_golden(name)
    kw = oracle.RESTORMER_TASKS[meta["task"]]
    m = build_restormer(kw, meta["wseed"], "bf16")
    x = oracle.synth_image(meta["shape"], meta["xseed"], meta["sigma"])
    clean = oracle.synth_image(meta["shape"], meta["xseed"], None).numpy()
    y = m(x.cuda()).cpu().numpy()
    y2 = m(x.cuda()).cpu().numpy()
    assert np.array_equal(y, y2)
    err = float(np.abs(y.astype(np.float64) - z["y64"]).max())
    d_psnr = abs(psnr(y, clean[:, : y.shape[1]]) - psnr(z["y64"], clean[:, : y.shape[1]]))
    record(f"{name}[bf16]", max_abs=err, psnr_delta_db=d_psnr, psnr_vs_ref_db=psnr(y, z["y64"]))
    assert np.isfinite(y).all() and err <= BF16_MAXABS and d_psnr <= BF16_PSNR_DB, (name, err, d_psnr)
    yh = m.set_mode("half")(x.cuda()).cpu().numpy()
    assert not np.array_equal(y, yh)                      # a real second flavour, not an alias of the fp16 mode
    assert float(np.abs(yh.astype(np.float64) - z["y64"]).max()) < err


def test_dncnn_bf16_mode_reported_separately():
    name = golden_names("dncnn")[0]
    meta, z = load_golden(name)
    n = meta["in_nc"]
    m = M.DnCNN(n, n, 64, meta["nb"], meta["act_mode"]).eval()
    m.load_state_dict(oracle.synth_state_dict(oracle.dncnn_schema(n, n, 64, meta["nb"], meta["act_mode"]),
                                              meta["wseed"]), strict=True)
    m = m.cuda().set_mode("bf16")
    x = oracle.synth_image(meta["shape"], meta["xseed"], meta["sigma"])
    y = m(x.cuda()).cpu().numpy()
    err = float(np.abs(y.astype(np.float64) - z["y64"]).max())
    record(f"{name}[bf16]", max_abs=err)
    assert np.isfinite(y).all() and err <= BF16_MAXABS, err


@pytest.mark.parametrize("name", golden_names("dncnn"))
def test_dncnn_vs_reference_golden(name):
    meta, z = load_golden(name)
    n = meta["in_nc"]
    m = M.DnCNN(n, n, 64, meta["nb"], meta["act_mode"]).eval()
    m.load_state_dict(oracle.synth_state_dict(oracle.dncnn_schema(n, n, 64, meta["nb"], meta["act_mode"]),
                                              meta["wseed"]), strict=True)
    m = m.cuda()
    x = oracle.synth_image(meta["shape"], meta["xseed"], meta["sigma"])
    clean = oracle.synth_image(meta["shape"], meta["xseed"], None).numpy()
    y = m(x.cuda()).cpu().numpy()
    check_parity(name, y, z["y64"], clean)


def run_block(meta, sd, x_nchw, mode=0):
    """One TransformerBlock through ir_block_* (channels-last in place)."""
    lib = _native.lib()
    Cc, heads = meta["C"], meta["heads"]
    wb = int(meta["LayerNorm_type"] != "BiasFree")
    B, _, H, W = x_nchw.shape
    params = [v.cuda().contiguous() for v in sd.values()]
    nbytes = lib.ir_block_packed_bytes(Cc, heads, 2.66, 0, wb, mode)
    packed = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    _native.check(lib.ir_block_pack_weights(Cc, heads, 2.66, 0, wb, _native.ptr_array(params), len(params),
                                            packed.data_ptr(), nbytes, mode, stream))
    ws = torch.empty(lib.ir_block_workspace_bytes(Cc, heads, 2.66, B, H, W, mode), dtype=torch.uint8, device="cuda")
    xg = x_nchw.cuda().contiguous()
    xl = torch.empty(B * H * W * Cc, dtype=torch.float32, device="cuda")
    _native.check(lib.ir_nchw_to_nhwc(xg.data_ptr(), xl.data_ptr(), B, Cc, H, W, stream))
    # the layout helper must agree with torch's permute
    assert torch.equal(xl.view(B, H, W, Cc), xg.permute(0, 2, 3, 1))
    _native.check(lib.ir_block_forward(Cc, heads, 2.66, 0, wb, packed.data_ptr(), xl.data_ptr(), B, H, W,
                                       ws.data_ptr(), ws.numel(), mode, stream))
    out = torch.empty_like(xg)
    _native.check(lib.ir_nhwc_to_nchw(xl.data_ptr(), out.data_ptr(), B, Cc, H, W, stream))
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])   # IR_MODE_FP32 (tf32), IR_MODE_HALF, IR_MODE_FP32_SIMT (exact fp32), IR_MODE_FP32_STRICT
@pytest.mark.parametrize("name", golden_names("block"))
def test_transformer_block_vs_reference_golden(name, mode):
    meta, z = load_golden(name)
    wb = meta["LayerNorm_type"] != "BiasFree"
    sd = oracle.synth_state_dict(oracle.synth._block_schema("blk", meta["C"], meta["heads"], 2.66, False, wb),
                                 meta["wseed"])
    x = oracle.synth_tensor(meta["shape"], meta["xseed"], -1.0, 1.0)
    y = run_block(meta, sd, x, mode)
    err = float(np.abs(y.astype(np.float64) - z["y64"]).max())
    record(f"{name}[mode{mode}]", max_abs=err)
    assert err <= (2e-5 if mode == 2 else TOL_MAXABS), (name, err)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("task,shape,wseed,xseed", [
    ("color_denoise", (1, 3, 40, 72), 61, 71),
    ("motion_deblur", (2, 3, 24, 16), 62, 72),
    ("defocus_dual", (1, 6, 16, 48), 63, 73),
])
def test_restormer_vs_oracle_fresh_inputs(task, shape, wseed, xseed, mode):
    kw = oracle.RESTORMER_TASKS[task]
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), wseed)
    x = oracle.synth_image(shape, xseed, 25.0)
    y_ref = oracle.restormer_forward({k: v.double() for k, v in sd.items()}, x.double()).numpy()
    m = build_restormer(kw, wseed, mode)
    y = m(x.cuda()).cpu().numpy()
    clean = oracle.synth_image(shape, xseed, None).numpy()[:, : y.shape[1]]
    check_parity(f"oracle_{task}_{shape[2]}x{shape[3]}[{mode}]", y, y_ref, clean)


def test_restormer_with_conv_bias_and_custom_widths_vs_oracle():
    kw = dict(inp_channels=3, out_channels=3, dim=32, num_blocks=[1, 2, 1, 2], num_refinement_blocks=1,
              heads=[1, 2, 2, 4], ffn_expansion_factor=2.0, bias=True, LayerNorm_type="WithBias")
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 64)
    x = oracle.synth_image((2, 3, 32, 24), 74, None)
    y_ref = oracle.restormer_forward({k: v.double() for k, v in sd.items()}, x.double()).numpy()
    m = M.Restormer(**kw).eval()
    m.load_state_dict(sd, strict=True)
    y = m.cuda()(x.cuda()).cpu().numpy()
    check_parity("oracle_bias_dim32", y, y_ref, x.numpy())


def test_dncnn_vs_oracle_odd_sizes():
    sd = oracle.synth_state_dict(oracle.dncnn_schema(3, 3, 64, 20, "R"), 65)
    x = oracle.synth_image((2, 3, 19, 31), 75, 50.0)
    y_ref = oracle.dncnn_forward({k: v.double() for k, v in sd.items()}, x.double()).numpy()
    m = M.DnCNN(3, 3, 64, 20, "R").eval()
    m.load_state_dict(sd, strict=True)
    y = m.cuda()(x.cuda()).cpu().numpy()
    check_parity("oracle_dncnn_color_19x31", y, y_ref, oracle.synth_image((2, 3, 19, 31), 75, None).numpy())


def test_shape_validation_raises_before_launch():
    m = build_restormer(oracle.RESTORMER_TASKS["color_denoise"], 1)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 60, 64, device="cuda"))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 1, 64, 64, device="cuda"))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 64, 64, device="cuda", dtype=torch.float16))


@pytest.mark.parametrize("mode", MODES)
def test_full_size_properties_config2_gray_denoise(mode):
    """BASELINE config 2 (batch 8 of 512x512 gray): properties that hold at any size.
    (a) run-to-run determinism, bit-exact; (b) permuting the batch permutes the output, bit-exact;
    (c) an image's result does not depend on its batch-mates: only the pixel split of the Gram reduction differs,
        which moves the folded attention matrix by fp32 noise before its tf32 rounding -> within the parity bar."""
    kw = oracle.RESTORMER_TASKS["gray_denoise"]
    m = build_restormer(kw, 81, mode)
    x = oracle.synth_image((8, 1, 512, 512), 91, 25.0).cuda()
    y1 = m(x)
    y2 = m(x)
    assert torch.equal(y1, y2)
    assert torch.isfinite(y1).all()
    perm = torch.tensor([3, 0, 7, 1, 6, 2, 5, 4], device="cuda")
    yp = m(x[perm].contiguous())
    assert torch.equal(yp, y1[perm])
    ys = m(x[2:3].contiguous())
    err = float((ys - y1[2:3]).abs().max())
    record(f"config2_single_vs_batched[{mode}]", max_abs=err)
    assert err <= TOL_MAXABS


@pytest.mark.parametrize("mode", MODES)
def test_full_size_properties_config3_real_denoise_and_oracle_crop(mode):
    """BASELINE config 3 shape (256x256 colour patches; batch reduced to 4 to bound test time) and a
    256x256 single image against the CPU oracle (5 s on the host)."""
    kw = oracle.RESTORMER_TASKS["real_denoise"]
    m = build_restormer(kw, 82, mode)
    x = oracle.synth_image((4, 3, 256, 256), 92, None)
    y = m(x.cuda())
    assert torch.equal(y, m(x.cuda()))
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 82)
    y_ref = oracle.restormer_forward(sd, x[1:2]).numpy()      # fp32 oracle (self-noise ~3e-7)
    check_parity(f"config3_image1_vs_oracle[{mode}]", y[1:2].cpu().numpy(), y_ref, x[1:2].numpy())


def test_dncnn_config1_full_size():
    """BASELINE config 1: DnCNN-S gray sigma=25, one 256x256 image, both act modes."""
    for act, seed in (("BR", 83), ("R", 84)):
        sd = oracle.synth_state_dict(oracle.dncnn_schema(1, 1, 64, 17, act), seed)
        x = oracle.synth_image((1, 1, 256, 256), 93, 25.0)
        y_ref = oracle.dncnn_forward({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()},
                                     x.double()).numpy()
        m = M.DnCNN(1, 1, 64, 17, act).eval()
        m.load_state_dict(sd, strict=True)
        y = m.cuda()(x.cuda()).cpu().numpy()
        check_parity(f"config1_dncnn_{act}", y, y_ref, oracle.synth_image((1, 1, 256, 256), 93, None).numpy())


# ---------------------------------------------------------------------------------------------------------------
# Parity AT the benchmarked sizes: MDTA's Gram is a reduction over all H*W pixels
# (restormer.py:121-125), so the rounding error of the tensor-core operands could grow with the image; these cases
# compare a full 512x512 forward with the CPU oracle (fp32, ~7 s each on the host; its own noise is ~3e-7).
# ---------------------------------------------------------------------------------------------------------------
_ORACLE_CACHE = {}


def oracle_512(task, wseed, xseed, sigma, batch_index=None):
    key = (task, wseed, xseed, sigma, batch_index)
    if key not in _ORACLE_CACHE:
        kw = oracle.RESTORMER_TASKS[task]
        sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), wseed)
        n = kw["inp_channels"]
        if batch_index is None:
            x = oracle.synth_image((1, n, 512, 512), xseed, sigma)
        else:
            x = oracle.synth_image((8, n, 512, 512), xseed, sigma)[batch_index:batch_index + 1]
        _ORACLE_CACHE[key] = oracle.restormer_forward(sd, x).numpy()
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("task,wseed,xseed,sigma", [
    ("gray_denoise", 85, 95, 25.0),      # BASELINE config 2 element: 1 x 1 x 512 x 512, BiasFree
    ("motion_deblur", 86, 96, None),     # config 4 tile: 1 x 3 x 512 x 512, WithBias
    ("defocus_dual", 87, 97, None),      # config 5 tile: 1 x 6 x 512 x 512 dual-pixel
])
def test_restormer_512_vs_oracle(task, wseed, xseed, sigma, mode):
    kw = oracle.RESTORMER_TASKS[task]
    y_ref = oracle_512(task, wseed, xseed, sigma)
    m = build_restormer(kw, wseed, mode)
    x = oracle.synth_image((1, kw["inp_channels"], 512, 512), xseed, sigma)
    y = m(x.cuda()).cpu().numpy()
    clean = oracle.synth_image((1, kw["inp_channels"], 512, 512), xseed, None).numpy()[:, : y.shape[1]]
    check_parity(f"size512_{task}[{mode}]", y, y_ref, clean)


@pytest.mark.parametrize("mode", MODES)
def test_config2_batch_element_vs_oracle(mode):
    """One image taken out of a real 8 x 512 x 512 batch (the bench workload) against the oracle on that image."""
    kw = oracle.RESTORMER_TASKS["gray_denoise"]
    y_ref = oracle_512("gray_denoise", 81, 91, 25.0, batch_index=5)
    m = build_restormer(kw, 81, mode)
    x = oracle.synth_image((8, 1, 512, 512), 91, 25.0)
    y = m(x.cuda())[5:6].cpu().numpy()
    clean = oracle.synth_image((8, 1, 512, 512), 91, None).numpy()[5:6]
    check_parity(f"config2_batch8_element5[{mode}]", y, y_ref, clean)


# ---------------------------------------------------------------------------------------------------------------
# fp16 range (VERDICT r01 "What's weak"): IR_MODE_FP32 holds operand-only tensors as fp16.  Scaled LayerNorm gains
# push v / the GDFN hidden tensor towards fp16's limits; the pack-time guard must move such a model to
# IR_MODE_FP32_STRICT, the device conversions must saturate instead of producing inf, and below the guard's threshold
# the fast mode must still meet the bar (relative to the output's scale).
# ---------------------------------------------------------------------------------------------------------------
def scaled_ln_state_dict(kw, wseed, gain):
    """LayerNorm gains x g with the block outputs kept at their scale: norm1.weight x g and attn.project_out / g (q, k
    are L2-normalised, v grows x g); norm2.weight x g and ffn.project_out / g^2 (hidden grows x g, the gated product
    x g^2).  The network stays as well conditioned as the unscaled one -- the output remains an O(1) image and the 1e-3
    bar keeps its meaning -- while v / hidden / gated sweep fp16's range.  (Scaling the gains alone makes every block's
    update dwarf the residual stream: the forward then amplifies ANY rounding, tf32's included, and says nothing about
    range.)"""
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), wseed)
    out = {}
    for k, v in sd.items():
        if (".norm1." in k or ".norm2." in k) and k.endswith("weight"):
            v = v * gain
        elif k.endswith("attn.project_out.weight"):
            v = v / gain
        elif k.endswith("ffn.project_out.weight"):
            v = v / (gain * gain)
        out[k] = v
    return out


@pytest.mark.parametrize("task,gain,expect", [
    ("gray_denoise", 2.0, "fp32"),            # BiasFree, below the guard: fast mode
    ("motion_deblur", 12.0, "fp32"),          # WithBias, just below the guard: gated products reach ~1e3
    ("gray_denoise", 50.0, "fp32_strict"),    # the guard trips: tf32 operands / fp32 tensors everywhere
    ("motion_deblur", 50.0, "fp32_strict"),
    ("motion_deblur", 400.0, "fp32_strict"),  # hidden ~1e3, gated products ~1e6: really outside fp16's range (> 65504)
])
def test_fp16_range_guard_and_strict_mode(task, gain, expect):
    kw = oracle.RESTORMER_TASKS[task]
    sd = scaled_ln_state_dict(kw, 88, gain)
    x = oracle.synth_image((1, kw["inp_channels"], 64, 64), 98, 25.0)
    y_ref = oracle.restormer_forward({k: v.double() for k, v in sd.items()}, x.double()).numpy()
    m = M.Restormer(**kw, bias=False).eval()
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    assert m.resolved_mode() == expect, (m.fp16_range_bound(), expect)
    y = m(x.cuda()).cpu().numpy()
    scale = max(1.0, float(np.abs(y_ref).max()))
    rel = float(np.abs(y - y_ref).max()) / scale
    record(f"range_{task}_gain{gain:g}", mode=m.resolved_mode(), bound=m.fp16_range_bound(), rel_err=rel, out_scale=scale)
    assert np.isfinite(y).all()
    assert rel <= TOL_MAXABS, (task, gain, rel)
    if expect == "fp32_strict":
        # the fast mode on the same weights (guard off): saturating conversions keep every value finite
        yf = m.set_range_guard(False)(x.cuda()).cpu().numpy()
        assert m.resolved_mode() == "fp32"
        record(f"range_{task}_gain{gain:g}_guard_off", rel_err=float(np.abs(yf - y_ref).max()) / scale)
        assert np.isfinite(yf).all()


def test_strict_mode_parity_on_goldens():
    """IR_MODE_FP32_STRICT (no fp16 tensor anywhere) against the reference goldens, same bar."""
    for name in golden_names("restormer")[:3]:
        meta, z = load_golden(name)
        kw = oracle.RESTORMER_TASKS[meta["task"]]
        m = build_restormer(kw, meta["wseed"], "fp32_strict")
        x = oracle.synth_image(meta["shape"], meta["xseed"], meta["sigma"])
        clean = oracle.synth_image(meta["shape"], meta["xseed"], None).numpy()
        y = m(x.cuda()).cpu().numpy()
        check_parity(f"{name}[fp32_strict]", y, z["y64"], clean[:, : y.shape[1]])


# ---------------------------------------------------------------------------------------------------------------
# Own memory check.  compute-sanitizer is closed on this GPU pool (gpurun_out/sanitizer_*: "runs under it have left GPUs
# needing a reset"), so out-of-bounds writes are hunted with guard zones instead: input, output, packed weights and
# workspace sit between canary regions that must come back untouched, the workspace is exactly the size the library
# asked for, and the forward must still match the oracle.  Shapes with partial 8 x 16 tiles and odd batch counts.
# ---------------------------------------------------------------------------------------------------------------
GUARD = 1 << 20


def _guarded(nbytes, device="cuda"):
    buf = torch.full((GUARD + nbytes + GUARD,), 0xA5, dtype=torch.uint8, device=device)
    return buf, buf[GUARD:GUARD + nbytes]


def _guards_intact(buf, nbytes):
    return bool((buf[:GUARD] == 0xA5).all()) and bool((buf[GUARD + nbytes:] == 0xA5).all())


@pytest.mark.parametrize("mode", ["fp32", "half", "fp32_strict"])
@pytest.mark.parametrize("task,shape", [
    ("color_denoise", (1, 3, 40, 72)),      # partial tiles at every level
    ("motion_deblur", (3, 3, 24, 16)),      # narrower than one tile, odd batch
    ("gray_denoise", (2, 1, 64, 64)),
    ("defocus_dual", (1, 6, 16, 48)),
])
def test_guard_zones_stay_intact(task, shape, mode):
    kw = oracle.RESTORMER_TASKS[task]
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 66)
    m = build_restormer(kw, 66, mode)
    x = oracle.synth_image(shape, 76, 25.0)
    y_plain = m(x.cuda())                                  # packs the weights
    lib = _native.lib()
    B, _, H, W = shape
    native_mode = m._native_mode
    packed_src = m._packed[2]
    pbuf, packed = _guarded(packed_src.numel())
    packed.copy_(packed_src)
    ws_bytes = lib.ir_restormer_workspace_bytes(C.byref(m._cfg), B, H, W, native_mode)
    wbuf, ws = _guarded(ws_bytes)
    xbuf, xv = _guarded(x.numel() * 4)
    xv.view(torch.float32).copy_(x.flatten())
    ybytes = B * m.out_channels * H * W * 4
    ybuf, yv = _guarded(ybytes)
    stream = torch.cuda.current_stream().cuda_stream
    _native.check(lib.ir_restormer_forward(C.byref(m._cfg), packed.data_ptr(), xv.data_ptr(), yv.data_ptr(), B, H, W,
                                           ws.data_ptr(), ws_bytes, native_mode, stream))
    torch.cuda.synchronize()
    assert _guards_intact(pbuf, packed_src.numel()), "write outside the packed weights"
    assert _guards_intact(wbuf, ws_bytes), "write outside the workspace"
    assert _guards_intact(xbuf, x.numel() * 4), "write outside the input"
    assert _guards_intact(ybuf, ybytes), "write outside the output"
    assert torch.equal(packed, packed_src), "the forward modified the packed weights"
    assert torch.equal(xv.view(torch.float32), x.flatten().cuda()), "the forward modified its input"
    y = yv.view(torch.float32).view(B, m.out_channels, H, W)
    assert torch.equal(y, y_plain)                         # and the run is bit-identical to the module's own call
    # one byte less than the library asked for must be refused, not overrun
    assert lib.ir_restormer_forward(C.byref(m._cfg), packed.data_ptr(), xv.data_ptr(), yv.data_ptr(), B, H, W,
                                    ws.data_ptr(), ws_bytes - 1, native_mode, stream) == _native.IR_ERR_WORKSPACE


def test_cuda_graph_replay_is_bit_identical_and_reused():
    """ir_*_forward_graph: plain launches on the first call of a key, capture on the second, replay afterwards; inputs
    change between calls (staging slots), results must equal the un-captured path bit for bit."""
    lib = _native.lib()
    lib.ir_graph_cache_clear()
    kw = oracle.RESTORMER_TASKS["color_denoise"]
    m = build_restormer(kw, 67, "fp32").set_cuda_graphs(True)
    ref = build_restormer(kw, 67, "fp32").set_cuda_graphs(False)
    xs = [oracle.synth_image((1, 3, 64, 80), 300 + i, 25.0).cuda() for i in range(4)]
    for x in xs:
        assert torch.equal(m(x), ref(x))
    e, c, r = (C.c_longlong(), C.c_longlong(), C.c_longlong())
    c0 = c.value
    lib.ir_graph_cache_stats(C.byref(e), C.byref(c), C.byref(r))
    assert e.value == 1 and r.value >= 3          # one key: call 1 plain, call 2 capture + launch, calls 3-4 replays
    # a second shape is a second key; "auto" picks the graph path for tile-sized inputs
    auto = build_restormer(kw, 67, "fp32")
    assert auto._use_graph(1, 512, 512) and not auto._use_graph(8, 512, 512)
    x2 = oracle.synth_image((2, 3, 32, 32), 310, 25.0).cuda()
    for _ in range(3):
        assert torch.equal(m(x2), ref(x2))
    d = M.DnCNN(1, 1, 64, 17, "R").eval()
    d.load_state_dict(oracle.synth_state_dict(oracle.dncnn_schema(1, 1, 64, 17, "R"), 8), strict=True)
    d = d.cuda()
    dx = oracle.synth_image((1, 1, 48, 56), 9, 25.0).cuda()
    y_plain = d.set_cuda_graphs(False)(dx)
    d.set_cuda_graphs(True)
    for _ in range(3):
        assert torch.equal(d(dx), y_plain)
    assert lib.ir_graph_cache_clear() == 0


def test_native_library_is_the_loaded_code():
    """The forward must run from the in-tree .so (no silent PyTorch path)."""
    maps = open("/proc/self/maps").read()
    assert "libirb200.so" in maps
