"""Image-quality indexes of the reference's result tables (assess_psnr / assess_ssim of
ImageQualityIndexes.jl, called at /root/reference/src/BPLDenoising.jl:201-204, :400-403).

Host-side post-processing, off the hot path.  ImageQualityIndexes is not vendored with the
reference, so these follow its documented defaults: PSNR with peak value 1 for floating-point
images; SSIM of Wang et al. (2004) with an 11×11 Gaussian window (σ = 1.5), K = (0.01, 0.03),
peak 1, the window applied with symmetric boundary padding, mean over all pixels.
"""
from __future__ import annotations

import numpy as np


def psnr_from_sqerr(sqerr: float, npix: int, peakval: float = 1.0) -> float:
    mse = sqerr / npix
    return float("inf") if mse == 0 else float(20.0 * np.log10(peakval) - 10.0 * np.log10(mse))


def assess_psnr(x, ref, peakval: float = 1.0) -> float:
    d = np.asarray(x, dtype=np.float64) - np.asarray(ref, dtype=np.float64)
    return psnr_from_sqerr(float(np.vdot(d, d)), d.size, peakval)


def _gauss_filter(a: np.ndarray, w: np.ndarray) -> np.ndarray:
    r = len(w) // 2
    p = np.pad(a, r, mode="symmetric")
    t = sum(w[k] * p[k:k + a.shape[0], :] for k in range(len(w)))          # separable: rows …
    return sum(w[k] * t[:, k:k + a.shape[1]] for k in range(len(w)))       # … then columns


def assess_ssim(x, ref, peakval: float = 1.0, K=(0.01, 0.03), sigma: float = 1.5, size: int = 11) -> float:
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(ref, dtype=np.float64)
    k = np.arange(size) - size // 2
    w = np.exp(-(k * k) / (2.0 * sigma * sigma))
    w /= w.sum()
    C1, C2 = (K[0] * peakval) ** 2, (K[1] * peakval) ** 2
    mx, my = _gauss_filter(x, w), _gauss_filter(y, w)
    sxx = _gauss_filter(x * x, w) - mx * mx
    syy = _gauss_filter(y * y, w) - my * my
    sxy = _gauss_filter(x * y, w) - mx * my
    ssim = ((2 * mx * my + C1) * (2 * sxy + C2)) / ((mx * mx + my * my + C1) * (sxx + syy + C2))
    return float(ssim.mean())
