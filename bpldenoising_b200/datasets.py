"""Synthetic datasets of BASELINE.md §5 (configs 4 and 5).

Frozen generator: piecewise-constant phantom (8 random discs /
rectangles, levels U[0,1]) plus a smooth ramp, clipped to [0,1]; noisy =
clip(truth + σ·N(0,1), 0, 1) quantised to k/255 like the reference's 8-bit PNG
datasets (/root/reference/src/Datasets.jl:54-65).
"""
from __future__ import annotations

import os

import numpy as np

# /root/reference/src/Datasets.jl:11-17
remotedatasets = ["cameraman_128_5", "cameraman_128_10", "faces_train_128_10", "faces_val_128_10", "circle_128_10"]


def full_datasetname(datasetname: str) -> str:
    """First known dataset whose name starts with `datasetname` (Datasets.jl:27-49); the
    reference falls back to a fuzzy match and otherwise throws ArgumentError — here ValueError."""
    for name in remotedatasets:
        if name.startswith(datasetname):
            return name
    raise ValueError(f'"{datasetname}" not found in remotedatasets {remotedatasets}')


def load_dataset(datasetfiles: str):
    """load_dataset (Datasets.jl:54-65): `filelist.txt` holds one `true.png,data.png` pair per
    line; 8-bit grey PNGs become k/255 in M×N×O Float64 stacks (column-major).  → (true, data)."""
    from PIL import Image
    with open(os.path.join(datasetfiles, "filelist.txt")) as fh:
        pairs = [ln.strip() for ln in fh if ln.strip()]

    def read(name):
        im = Image.open(os.path.join(datasetfiles, name))
        if im.mode == "1":                      # circle truth: 1-bit → {0, 1}
            return np.asarray(im, dtype=np.float64)
        return np.asarray(im.convert("L"), dtype=np.float64) / 255.0

    first = read(pairs[0].split(",")[0])
    M, N = first.shape
    true_images = np.zeros((M, N, len(pairs)), order="F")
    data_images = np.zeros((M, N, len(pairs)), order="F")
    for i, pair in enumerate(pairs):
        t, d = pair.split(",")[:2]
        true_images[:, :, i] = read(t)
        data_images[:, :, i] = read(d)
    return true_images, data_images


def testdataset(datasetname: str, dataset_dir: str = "BPLDenoising/datasets/"):
    """testdataset(datasetname) (Datasets.jl:19-25); the reference resolves `dataset_dir`
    relative to the current working directory (Datasets.jl:9)."""
    return load_dataset(os.path.join(dataset_dir, full_datasetname(datasetname)))


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
