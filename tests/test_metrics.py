"""PSNR / SSIM (calculate_metrics src/utils.py:134-156): the oracle against closed forms and an
independent brute-force restatement (CPU), and ir_image_metrics against the oracle (GPU)."""
import math

import numpy as np
import pytest
import torch

from oracle import metrics_ref
from oracle.make_golden_tiling import make_image
from conftest import record


def brute_ssim(a, b, R):
    """Definition of scikit-image's default SSIM written with explicit 7x7 windows in exact rational-free float64."""
    a, b = a.astype(np.float64), b.astype(np.float64)
    H, W = a.shape
    c1, c2 = (0.01 * R) ** 2, (0.03 * R) ** 2
    tot, n = 0.0, 0
    for y in range(3, H - 3):
        for x in range(3, W - 3):
            p, q = a[y - 3:y + 4, x - 3:x + 4], b[y - 3:y + 4, x - 3:x + 4]
            ux, uy = p.mean(), q.mean()
            vx = ((p - ux) ** 2).sum() / 48.0
            vy = ((q - uy) ** 2).sum() / 48.0
            vxy = ((p - ux) * (q - uy)).sum() / 48.0
            tot += ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
            n += 1
    return tot / n


def pair(dtype, h, w, c, seed, sigma):
    """A clean image and a degraded copy of the same dtype."""
    clean = make_image(dtype, h, w, c, seed)
    if clean.dtype == np.float32:
        clean = (clean / 3.0).astype(np.float32)          # make_image's float images span [0, 3)
    rng = np.random.RandomState(seed + 1)
    if clean.dtype == np.float32:
        noisy = np.clip(clean + rng.normal(0, sigma, clean.shape), 0, 1).astype(np.float32)
    else:
        top = np.iinfo(clean.dtype).max
        noisy = np.clip(np.round(clean.astype(np.float64) + rng.normal(0, sigma * top, clean.shape)), 0, top).astype(clean.dtype)
    return noisy, clean


def test_oracle_closed_forms():
    img = make_image("uint8", 20, 24, 3, 3)
    p, s = metrics_ref.calculate_metrics(img, img)
    assert math.isinf(p) and p > 0 and abs(s - 1.0) < 1e-15
    a = np.full((16, 16, 1), 100, np.uint8)
    b = np.full((16, 16, 1), 110, np.uint8)
    p, s = metrics_ref.calculate_metrics(b, a)
    assert abs(p - 10 * math.log10(255.0 ** 2 / 100.0)) < 1e-12
    c1 = (0.01 * 255) ** 2
    assert abs(s - (2 * 100 * 110 + c1) / (100 ** 2 + 110 ** 2 + c1)) < 1e-12
    with pytest.raises(ValueError):
        metrics_ref.ssim2d(np.zeros((6, 9), np.uint8), np.zeros((6, 9), np.uint8), 255)


@pytest.mark.parametrize("dtype,R", [("uint8", 255), ("uint16", 65535), ("float32", 1.0)])
def test_oracle_vs_brute_force(dtype, R):
    noisy, clean = pair(dtype, 15, 18, 1, 11, 0.08)
    want = brute_ssim(clean[:, :, 0], noisy[:, :, 0], R)
    _, got = metrics_ref.calculate_metrics(noisy, clean)
    assert abs(got - want) < (2e-6 if dtype == "float32" else 1e-11)       # float32 images are filtered in float32
    mse = np.mean((clean.astype(np.float64) - noisy.astype(np.float64)) ** 2)
    p, _ = metrics_ref.calculate_metrics(noisy, clean)
    assert abs(p - 10 * math.log10(R * R / mse)) < (1e-4 if dtype == "float32" else 1e-12)


def test_oracle_colour_is_channel_mean():
    noisy, clean = pair("uint8", 24, 31, 3, 5, 0.1)
    _, s = metrics_ref.calculate_metrics(noisy, clean)
    per = [metrics_ref.ssim2d(clean[:, :, c], noisy[:, :, c], 255) for c in range(3)]
    assert abs(s - np.mean(per)) < 1e-15


def test_package_metrics_refuses_cpu():
    from image_restoration_models_b200 import metrics
    img = make_image("uint8", 16, 16, 3, 1)
    with pytest.raises(RuntimeError):
        metrics.calculate_metrics(img, img, device="cpu")


GPU_CASES = [("uint8", 70, 90, 3, 0.1), ("uint8", 50, 37, 1, 0.06), ("uint16", 45, 64, 3, 0.05), ("float32", 40, 40, 3, 0.1),
             ("uint8", 7, 7, 1, 0.2), ("uint8", 720, 1280, 3, 0.1), ("uint16", 1120, 1680, 3, 0.02)]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,h,w,c,sigma", GPU_CASES)
def test_device_metrics_match_oracle(dtype, h, w, c, sigma):
    from image_restoration_models_b200 import metrics
    noisy, clean = pair(dtype, h, w, c, 21, sigma)
    want_p, want_s = metrics_ref.calculate_metrics(noisy, clean)
    got_p, got_s = metrics.calculate_metrics(noisy, clean)
    again = metrics.calculate_metrics(torch.from_numpy(noisy).cuda(), torch.from_numpy(clean).cuda())
    assert again == (got_p, got_s)                       # fixed reduction order: bit-reproducible
    tol_p, tol_s = (1e-4, 2e-6) if dtype == "float32" else (1e-11, 1e-11)
    record(f"metrics_{dtype}_{h}x{w}x{c}", psnr=got_p, psnr_err=abs(got_p - want_p), ssim=got_s, ssim_err=abs(got_s - want_s))
    assert abs(got_p - want_p) <= tol_p and abs(got_s - want_s) <= tol_s, (got_p, want_p, got_s, want_s)
    if c == 1:                                           # HW and HWx1 take the same path (utils.py:151-154)
        assert metrics.calculate_metrics(noisy[:, :, 0], clean[:, :, 0]) == (got_p, got_s)


@pytest.mark.gpu
def test_device_metrics_edges():
    from image_restoration_models_b200 import metrics
    img = make_image("uint8", 33, 47, 3, 2)
    p, s = metrics.calculate_metrics(img, img)
    assert math.isinf(p) and p > 0 and s == 1.0
    noisy, clean = pair("float32", 32, 32, 1, 4, 0.1)
    want = metrics_ref.calculate_metrics(noisy, clean, data_range=2.0)
    got = metrics.calculate_metrics(noisy, clean, data_range=2.0)
    assert abs(got[0] - want[0]) < 1e-4 and abs(got[1] - want[1]) < 2e-6
    with pytest.raises(ValueError):
        metrics.calculate_metrics(img[:6], img[:6])                  # win_size exceeds the image extent
    with pytest.raises(ValueError):
        metrics.calculate_metrics(img, img[:, :40])                  # shape mismatch
    with pytest.raises(ValueError):
        metrics.calculate_metrics(np.zeros((16, 16, 2), np.uint8), np.zeros((16, 16, 2), np.uint8))
