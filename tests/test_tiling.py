"""Tiled-inference harness: oracle vs the reference's golden outputs, host logic of the
package (tile grid, window, partition) and the multi-rank path on CPU with gloo (world size 2).

The CPU tests drive image_restoration_models_b200.tiling.run_model_inference with a numpy backend built from the
oracle (test infrastructure) in place of the CUDA kernels; the GPU test uses the real kernels and must be bit-exact."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch

import oracle
from oracle import tiling_ref
from oracle.make_golden_tiling import CASES, NOISE_CASES, StandIn, make_image
from image_restoration_models_b200 import tiling
from conftest import GOLDEN, ROOT, load_golden


class NumpyBackend:
    """CPU stand-in for the CUDA gather/blend kernels, written with the oracle's arithmetic (tests only)."""

    def upload(self, a):
        return torch.from_numpy(np.ascontiguousarray(a))

    def gather(self, img, code, divisor, H, W, C, xy, T, th, tw, TH, TW, noise=None):
        x = img.numpy().astype(np.float32)
        if divisor != 1.0:
            x = x / np.float32(divisor)
        out = []
        for h0, w0 in xy.numpy():
            patch = x[h0:h0 + th, w0:w0 + tw].copy()
            if noise is not None:
                patch += noise.numpy()
                patch = np.clip(patch, 0, 1).astype(np.float32)
            t = torch.from_numpy(patch.transpose(2, 0, 1).copy()).unsqueeze(0)
            out.append(torch.nn.functional.pad(t, (0, TW - tw, 0, TH - th), "reflect") if (TH > th or TW > tw) else t)
        return torch.cat(out, 0)

    def blend(self, pred, xy, T, th, tw, TH, TW, window, H, W, C, out_dtype, code, scale, lo, hi, out=None):
        out = np.zeros((H, W, C), np.float32)
        wm = np.zeros((H, W, C), np.float32)
        win = np.repeat(window.numpy()[:th, :tw, None], C, axis=2)
        for t, (h0, w0) in enumerate(xy.numpy()):
            p = pred[t, :, :th, :tw].numpy().transpose(1, 2, 0)
            out[h0:h0 + th, w0:w0 + tw] += p * win
            wm[h0:h0 + th, w0:w0 + tw] += win
        out /= np.maximum(wm, 1e-8)
        v = np.clip(out * np.float32(scale), np.float32(lo), np.float32(hi))
        if code in (0, 1):
            v = v.round()
        return torch.from_numpy(v.astype({0: np.uint8, 1: np.uint16, 2: np.float32}[code]))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tiling_oracle_matches_reference_harness(case):
    name, dtype, h, w, c, ps, ov, use_pad, seed = case
    meta, z = load_golden(name)
    img = make_image(dtype, h, w, c, seed)
    out = tiling_ref.run_model_inference(StandIn().eval(), img, ps, ov, use_pad)
    assert out.dtype == z["out"].dtype and np.array_equal(out, z["out"])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tiling_host_logic_matches_reference_harness(case):
    name, dtype, h, w, c, ps, ov, use_pad, seed = case
    meta, z = load_golden(name)
    img = make_image(dtype, h, w, c, seed)
    out, ms = tiling.run_model_inference(StandIn().eval(), img, "cpu", patch_size=ps, patch_overlap=ov,
                                         pad=tiling.pad if use_pad else None, tile_batch=3, backend=NumpyBackend())
    assert out.dtype == z["out"].dtype and np.array_equal(out, z["out"])
    assert ms >= 0.0


@pytest.mark.parametrize("case", NOISE_CASES, ids=[c[0] for c in NOISE_CASES])
def test_noise_injection_oracle_and_host_logic_match_reference_harness(case):
    """need_degradation=True: add_gaussian_noise (src/utils.py:29-36, :408-409) reseeds numpy per tile, so one float64
    field per tile shape reproduces it; goldens from the unmodified reference harness."""
    name, dtype, h, w, c, ps, ov, use_pad, seed, sigma = case
    meta, z = load_golden(name)
    img = make_image(dtype, h, w, c, seed)
    ref = tiling_ref.run_model_inference(StandIn().eval(), img, ps, ov, use_pad, need_degradation=True, noise_level=sigma)
    assert np.array_equal(ref, z["out"])
    before = np.random.get_state()[1].copy()
    out, _ = tiling.run_model_inference(StandIn().eval(), img, "cpu", patch_size=ps, patch_overlap=ov,
                                        need_degradation=True, noise_level=sigma,
                                        pad=tiling.pad if use_pad else None, tile_batch=2, backend=NumpyBackend())
    assert np.array_equal(out, z["out"])
    assert np.array_equal(before, np.random.get_state()[1])          # the package does not reseed the global generator
    # need_degradation without a level, or a level without the flag, adds nothing (:408)
    clean, _ = tiling.run_model_inference(StandIn().eval(), img, "cpu", patch_size=ps, patch_overlap=ov,
                                          need_degradation=False, noise_level=sigma,
                                          pad=tiling.pad if use_pad else None, backend=NumpyBackend())
    assert not np.array_equal(clean, z["out"])


def test_called_exactly_like_get_model_prediction():
    """The reference's caller (src/utils.py:294-302) passes pad=pad, patch_size=, patch_overlap=, need_degradation=,
    noise_level= and progress_bar= by keyword after three positionals; the DnCNN branch (:303-310) passes no pad."""
    class Bar:
        def __init__(self): self.total, self.n = None, 0
        def tqdm(self, it, desc=None, total=None): self.total = total; return self
        def update(self): self.n += 1
    name, dtype, h, w, c, ps, ov, use_pad, seed = CASES[0]
    meta, z = load_golden(name)
    img = make_image(dtype, h, w, c, seed)
    bar = Bar()
    restored, ms = tiling.run_model_inference(StandIn().eval(), img, "cpu", pad=tiling.pad, patch_size=ps,
                                              patch_overlap=ov, need_degradation=False, noise_level=None,
                                              progress_bar=bar, backend=NumpyBackend())
    assert np.array_equal(restored, z["out"]) and bar.total == bar.n == 12
    # DnCNN branch: no pad argument -> pad=None -> tiles are NOT padded (H, W need not be multiples of 8)
    name, dtype, h, w, c, ps, ov, use_pad, seed = CASES[3]
    meta, z = load_golden(name)
    restored, _ = tiling.run_model_inference(StandIn().eval(), make_image(dtype, h, w, c, seed), "cpu", patch_size=ps,
                                             patch_overlap=ov, need_degradation=False, noise_level=None,
                                             progress_bar=None, backend=NumpyBackend())
    assert np.array_equal(restored, z["out"])
    # the fourth positional argument is `normalize`, as in the reference; only the default one is on this path
    tiling.run_model_inference(StandIn().eval(), make_image(dtype, h, w, c, seed), "cpu", tiling.normalize, ps, ov,
                               backend=NumpyBackend())
    with pytest.raises(NotImplementedError):
        tiling.run_model_inference(StandIn().eval(), img, "cpu", lambda a: a, backend=NumpyBackend())
    with pytest.raises(NotImplementedError):
        tiling.run_model_inference(StandIn().eval(), img, "cpu", postprocess=lambda a: a, backend=NumpyBackend())


def test_tile_grid_window_and_partition():
    # BASELINE configs 4 and 5: GoPro 720x1280 -> 6 tiles, DPDD 1120x1680 -> 12 tiles of 512
    assert tiling.tile_grid(720, 1280, 512, 96) == ([0, 208], [0, 416, 768], 512)
    assert tiling.tile_grid(1120, 1680, 512, 96) == ([0, 416, 608], [0, 416, 832, 1168], 512)
    assert tiling.tile_grid(512, 512, 256, 48) == ([0, 208, 256], [0, 208, 256], 256)     # heavy overlap of the last tile
    assert tiling.tile_grid(100, 60, 256, 48) == ([0], [0], 100)                          # patch clipped to max(h, w)
    assert tiling.tile_grid(33, 47, None, 8) == ([0], [0], 47)
    assert [tiling.padded_extent(n) for n in (8, 9, 15, 16, 100)] == [8, 16, 16, 16, 104]
    w = tiling.get_gaussian_weights(24, 24, 1)
    assert np.array_equal(w, np.load(os.path.join(GOLDEN, "gaussian_window_24.npy")))
    for n, world in ((6, 4), (12, 8), (3, 8), (192, 8)):
        parts = tiling.partition(n, world)
        assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        tiling.run_model_inference(lambda x: x, np.zeros((4, 4), np.uint8), "cpu", backend=NumpyBackend())
    np.testing.assert_array_equal(tiling.normalize(np.array([[[0, 255]]], np.uint8)), np.array([[[0.0, 1.0]]], np.float32))


def _worker(rank, world, port, case, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    name, dtype, h, w, c, ps, ov, use_pad, seed = case
    img = make_image(dtype, h, w, c, seed)
    out, _ = tiling.run_model_inference(StandIn().eval(), img, "cpu", patch_size=ps, patch_overlap=ov,
                                        pad=tiling.pad if use_pad else None, tile_batch=2, backend=NumpyBackend(), dst=1)
    if out is not None:
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), out)
    dist.destroy_process_group()


@pytest.mark.parametrize("case", [CASES[0], CASES[2]], ids=[CASES[0][0], CASES[2][0]])
def test_two_ranks_split_tiles_and_agree_with_single_rank(case, tmp_path):
    """World size 2 over gloo: ranks take contiguous tile slices, the predictions are gathered to ONE rank (dst = 1 here)
    which blends in the reference order -> its image is bit-identical to the single-rank (and the reference's) result;
    the other rank returns None and blends nothing."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, case, str(tmp_path)), nprocs=2, join=True)
    meta, z = load_golden(case[0])
    assert not (tmp_path / "rank0.npy").exists()
    assert np.array_equal(np.load(tmp_path / "rank1.npy"), z["out"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tiling_cuda_kernels_bit_exact_vs_reference_harness(case):
    name, dtype, h, w, c, ps, ov, use_pad, seed = case
    meta, z = load_golden(name)
    img = make_image(dtype, h, w, c, seed)
    out, _ = tiling.run_model_inference(StandIn().eval().cuda(), img, "cuda", patch_size=ps, patch_overlap=ov,
                                        pad=tiling.pad if use_pad else None, tile_batch=4)
    assert out.dtype == z["out"].dtype and np.array_equal(out, z["out"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", NOISE_CASES, ids=[c[0] for c in NOISE_CASES])
def test_noise_injection_cuda_kernel_bit_exact_vs_reference_harness(case):
    name, dtype, h, w, c, ps, ov, use_pad, seed, sigma = case
    meta, z = load_golden(name)
    img = make_image(dtype, h, w, c, seed)
    out, _ = tiling.run_model_inference(StandIn().eval().cuda(), img, "cuda", patch_size=ps, patch_overlap=ov,
                                        need_degradation=True, noise_level=sigma, pad=tiling.pad if use_pad else None,
                                        tile_batch=4)
    assert np.array_equal(out, z["out"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", [CASES[0], CASES[2]], ids=[CASES[0][0], CASES[2][0]])
def test_frame_pipeline_is_bit_identical_to_single_frame_path(case):
    """FramePipeline (streams, pinned double buffers, events) over 5 distinct frames == run_model_inference per frame,
    and frame 0 == the reference golden."""
    name, dtype, h, w, c, ps, ov, use_pad, seed = case
    meta, z = load_golden(name)
    frames = [make_image(dtype, h, w, c, seed + 40 * i) for i in range(5)]
    model = StandIn().eval().cuda()
    pipe = tiling.FramePipeline(model, "cuda", frames[0].shape, frames[0].dtype, patch_size=ps, patch_overlap=ov,
                                pad=tiling.pad if use_pad else None, tile_batch=5)
    outs = pipe.run(frames)
    assert np.array_equal(outs[0], z["out"])
    for f, o in zip(frames, outs):
        single, _ = tiling.run_model_inference(model, f, "cuda", patch_size=ps, patch_overlap=ov,
                                               pad=tiling.pad if use_pad else None)
        assert np.array_equal(o, single)
    assert pipe.run(frames[:3], copy_out=False) is None


@pytest.mark.gpu
def test_tiled_restormer_matches_oracle_harness():
    """End to end: uint8 colour image, 64x64 tiles with overlap 16 through the CUDA Restormer vs the oracle harness
    with the oracle forward.  The model parity bar is 1e-3 on [0,1] -> at most 1 LSB on uint8 after rounding."""
    import image_restoration_models_b200 as M
    kw = oracle.RESTORMER_TASKS["motion_deblur"]
    sd = oracle.synth_state_dict(oracle.restormer_schema(**kw), 91)
    img = make_image("uint8", 100, 148, 3, 9)
    ref = tiling_ref.run_model_inference(lambda x: oracle.restormer_forward(sd, x), img, 64, 16, True)
    m = M.Restormer(**kw, bias=False).eval()
    m.load_state_dict(sd, strict=True)
    out, _ = tiling.run_model_inference(m.cuda(), img, "cuda", patch_size=64, patch_overlap=16, pad=tiling.pad, tile_batch=8)
    diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.02


class DualStandIn(torch.nn.Module):
    """6 -> 3 channel stand-in for the dual-pixel model (left / right views in, one image out)."""

    def forward(self, x):
        return 0.5 * (x[:, :3] + x[:, 3:]) * 0.9 + 0.03


def test_dual_pixel_six_channel_input_three_channel_output_host_logic():
    """BASELINE config 5 shape family: the harness feeds all 6 channels to the model and blends min(3, C) = 3 output
    channels (src/utils.py:394-395,405); uint16 in, uint16 out."""
    img = make_image("uint16", 45, 64, 6, 77)
    ref = tiling_ref.run_model_inference(DualStandIn().eval(), img, 40, 12, True)
    out, _ = tiling.run_model_inference(DualStandIn().eval(), img, "cpu", patch_size=40, patch_overlap=12, pad=True,
                                        tile_batch=2, backend=NumpyBackend())
    assert out.shape == (45, 64, 3) and out.dtype == np.uint16 and np.array_equal(out, ref)


@pytest.mark.gpu
def test_dual_pixel_six_channel_input_cuda_kernels_bit_exact():
    img = make_image("uint16", 45, 64, 6, 77)
    ref = tiling_ref.run_model_inference(DualStandIn().eval(), img, 40, 12, True)
    out, _ = tiling.run_model_inference(DualStandIn().eval().cuda(), img, "cuda", patch_size=40, patch_overlap=12, pad=True,
                                        tile_batch=4)
    assert out.shape == (45, 64, 3) and np.array_equal(out, ref)
