"""CPU restatement of ``calculate_metrics`` (/root/reference/src/utils.py:134-156) -- TEST INFRASTRUCTURE ONLY.

The arithmetic lives in scikit-image (``requirements.txt``: ``scikit-image>=0.18.1``, unpinned; not installed in this
image and not vendored under /root/reference), so this file restates the published algorithm of
``skimage.metrics.peak_signal_noise_ratio`` and ``skimage.metrics.structural_similarity`` (Wang et al. 2004 with
scikit-image's defaults) on top of ``scipy.ndimage.uniform_filter`` -- the very filter scikit-image calls.  PARITY
UNPINNED for this row: the reference holds no golden PSNR / SSIM values and scikit-image cannot be run here; the pin is the
closed-form cases in tests/test_oracle_golden.py (identical images, constant offset, a hand-computed 7x7 window).
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter


def _float_type(dtype):
    # skimage._shared.utils._supported_float_type: float32 / float16 stay float32, everything else is float64
    return np.float32 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.float16)) else np.float64


def psnr(image_true: np.ndarray, image_test: np.ndarray, data_range) -> float:
    """skimage.metrics.peak_signal_noise_ratio: 10 log10(R^2 / mean((a - b)^2)), images converted to float first."""
    ft = _float_type(np.result_type(image_true.dtype, image_test.dtype))
    a, b = image_true.astype(ft), image_test.astype(ft)
    err = np.mean((a - b) ** 2, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return float(10 * np.log10((float(data_range) ** 2) / err))


def ssim2d(im1: np.ndarray, im2: np.ndarray, data_range, win_size=7, K1=0.01, K2=0.03) -> float:
    """skimage.metrics.structural_similarity on one 2-D channel: uniform window, sample covariance, cropped mean."""
    if min(im1.shape) < win_size:
        raise ValueError("win_size exceeds image extent.")
    ft = _float_type(im1.dtype)
    im1, im2 = im1.astype(ft, copy=False), im2.astype(ft, copy=False)
    NP = win_size ** im1.ndim
    cov_norm = NP / (NP - 1)
    ux, uy = uniform_filter(im1, size=win_size), uniform_filter(im2, size=win_size)
    uxx, uyy, uxy = (uniform_filter(im1 * im1, size=win_size), uniform_filter(im2 * im2, size=win_size),
                     uniform_filter(im1 * im2, size=win_size))
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))


def calculate_metrics(pred: np.ndarray, target: np.ndarray, data_range=None):
    """utils.py:134-156: data range from the dtype, PSNR over everything, SSIM per channel (channel_axis=2) or 2-D."""
    if data_range is None:
        data_range = 255 if pred.dtype == np.uint8 else 65535 if pred.dtype == np.uint16 else 1.0
    p = psnr(target, pred, data_range)
    if pred.ndim == 3 and pred.shape[2] == 3:
        s = float(np.mean([ssim2d(target[:, :, c], pred[:, :, c], data_range) for c in range(3)]))
    elif pred.ndim == 3 and pred.shape[2] == 1:
        s = ssim2d(target[:, :, 0], pred[:, :, 0], data_range)
    else:
        s = ssim2d(target, pred, data_range)
    return p, s
