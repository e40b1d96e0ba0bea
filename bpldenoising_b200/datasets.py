"""Synthetic datasets of BASELINE.md §5 (configs 4 and 5).

Frozen generator: piecewise-constant phantom (8 random discs /
rectangles, levels U[0,1]) plus a smooth ramp, clipped to [0,1]; noisy =
clip(truth + σ·N(0,1), 0, 1) quantised to k/255 like the reference's 8-bit PNG
datasets (/root/reference/src/Datasets.jl:54-65).
"""
from __future__ import annotations

import numpy as np


def synthetic_dataset(M: int, N: int, O: int, seed: int = 20240601, noise: float = 0.1):
    """Returns (truth, noisy), each M×N×O float64, Fortran order."""
    rng = np.random.default_rng(seed)
    ii, jj = np.meshgrid(np.arange(M), np.arange(N), indexing="ij")
    truth = np.zeros((M, N, O), order="F")
    noisy = np.zeros((M, N, O), order="F")
    for o in range(O):
        img = 0.25 * (ii / max(M - 1, 1) + jj / max(N - 1, 1)) * rng.uniform(0.0, 1.0)
        for _ in range(8):
            level = rng.uniform(0.0, 1.0)
            if rng.uniform() < 0.5:
                ci, cj = rng.uniform(0, M), rng.uniform(0, N)
                r = rng.uniform(0.05, 0.3) * min(M, N)
                mask = (ii - ci) ** 2 + (jj - cj) ** 2 <= r * r
            else:
                i0, j0 = rng.uniform(0, M), rng.uniform(0, N)
                h, w = rng.uniform(0.05, 0.4) * M, rng.uniform(0.05, 0.4) * N
                mask = (ii >= i0) & (ii < i0 + h) & (jj >= j0) & (jj < j0 + w)
            img = np.where(mask, level, img)
        img = np.clip(img, 0.0, 1.0)
        nz = np.clip(img + noise * rng.standard_normal((M, N)), 0.0, 1.0)
        truth[:, :, o] = np.round(img * 255.0) / 255.0
        noisy[:, :, o] = np.round(nz * 255.0) / 255.0
    return truth, noisy
