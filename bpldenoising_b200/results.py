"""On-disk artefacts of an experiment: `save_results`
(/root/reference/src/BPLDenoising.jl:185-299) — the optimiser log, the per-image quality table
(`img_num / orig_ssim / orig_psnr / out_ssim / out_psnr` + means) and the true / data / reco PNGs,
plus the up-sampled, linearly stretched parameter map for patch parameters (:252-257).

Host-side IO around the hot path; nothing here touches the GPU.  `write_log` belongs to the
un-vendored AlgTools package: the log is written as a `#` comment line followed by tab-separated
`iter time function_value gradient_norm radius residual` rows (the fields of `BilevelLogEntry`,
/root/reference/src/BilevelVisualise.jl:39-46).
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import quality

default_save_prefix = "output"   # BPLDenoising.jl:39


def linear_stretch(a: np.ndarray) -> np.ndarray:
    """adjust_histogram!(img, LinearStretching()) (:337-339): affine map of [min, max] onto [0, 1]."""
    a = np.asarray(a, dtype=np.float64)
    lo, hi = float(a.min()), float(a.max())
    return np.zeros_like(a) if hi == lo else (a - lo) / (hi - lo)


def _save_png(path: str, img: np.ndarray) -> None:
    from PIL import Image
    Image.fromarray(np.round(np.clip(img, 0.0, 1.0) * 255.0).astype(np.uint8)).save(path)   # grayimg → 8-bit


def write_log(path: str, log, comment: str) -> None:
    with open(path, "w") as io:
        io.write(comment if comment.endswith("\n") else comment + "\n")
        io.write("iter\ttime\tfunction_value\tgradient_norm\tradius\tresidual\n")
        for e in log:
            io.write(f"{e.iter}\t{e.time}\t{e.function_value}\t{e.gradient_norm}\t{e.radius}\t{e.residual}\n")


def quality_table(b, b_data, opt_img):
    """Rows of the quality file and the means of the output columns (:196-214)."""
    O = b.shape[2]
    rows = []
    for i in range(O):
        rows.append((i + 1,
                     quality.assess_ssim(b[:, :, i], b_data[:, :, i]), quality.assess_psnr(b[:, :, i], b_data[:, :, i]),
                     quality.assess_ssim(b[:, :, i], opt_img[:, :, i]), quality.assess_psnr(b[:, :, i], opt_img[:, :, i])))
    return rows, float(np.mean([r[3] for r in rows])), float(np.mean([r[4] for r in rows]))


def save_results(params: dict, b, b_data, x, opt_img, log, out_root: Optional[str] = None) -> dict:
    """save_results(params, b, b_data, x, opt_img, st) for scalar (:185-217), patch (:220-258) and
    m×n×3 sum-of-regularisers (:260-299) parameters.  `params` needs `dataset_name` and `save_prefix`; nothing is written when
    `params["save_results"]` is false (:186).  Returns the paths it wrote."""
    if not params.get("save_results", True):
        return {}
    b, b_data, opt_img = (np.asarray(a, dtype=np.float64) for a in (b, b_data, opt_img))
    out_path = os.path.join(out_root or default_save_prefix, params["dataset_name"])
    os.makedirs(out_path, exist_ok=True)
    stem = os.path.join(out_path, params["save_prefix"])
    written = {"log": stem + ".txt", "quality": stem + "_quality.txt", "png": []}
    write_log(written["log"], log, f"# params = {params}, x = {np.asarray(x).tolist()}")
    rows, mean_ssim, mean_psnr = quality_table(b, b_data, opt_img)
    with open(written["quality"], "w") as io:
        io.write("img_num \t orig_ssim \t orig_psnr \t out_ssim \t out_psnr\n")
        for i, ns, npn, os_, op in rows:
            io.write(f"{i}\t {ns} \t {npn} \t {os_} \t {op}\n")
            for tag, img in (("true", b), ("data", b_data), ("reco", opt_img)):
                p = f"{stem}_{tag}_{i}.png"
                _save_png(p, img[:, :, i - 1])
                written["png"].append(p)
        # the m×n×3 method accumulates `mean_psnr += mean_psnr` (:282): the file's last PSNR field is 0.0 there
        file_psnr = 0.0 if np.ndim(x) == 3 else mean_psnr
        io.write(f"\t\t\t\t\t {mean_ssim}\t {file_psnr}\n")
    xa = np.asarray(x, dtype=np.float64)
    if xa.ndim == 3:   # three patch parameters: up-sampled, stretched JOINTLY (adjust_histogram! on the M×N×3 array), one PNG each (:291-297)
        M, N = b.shape[:2]
        ii = (np.arange(M) * xa.shape[0]) // M
        jj = (np.arange(N) * xa.shape[1]) // N
        xbar = linear_stretch(np.stack([xa[:, :, k][np.ix_(ii, jj)] for k in range(xa.shape[2])], axis=2))
        for k in range(xa.shape[2]):
            p = f"{stem}_par_{k + 1}.png"
            _save_png(p, xbar[:, :, k])
            written["png"].append(p)
    if xa.ndim == 2:   # patch parameter: block-constant up-sampling (PatchOp, S7), stretched, as PNG (:252-257)
        M, N = b.shape[:2]
        ii = (np.arange(M) * xa.shape[0]) // M
        jj = (np.arange(N) * xa.shape[1]) // N
        p = stem + "_par.png"
        _save_png(p, linear_stretch(xa[np.ix_(ii, jj)]))
        written["png"].append(p)
    written["mean_ssim"], written["mean_psnr"] = mean_ssim, mean_psnr
    return written


# ------------------------------------------------------------------------------
# cost curves (/root/reference/src/BPLDenoising.jl:106-110, :153-157)
# ------------------------------------------------------------------------------
def save_cost_curve(dataset_name: str, parameter_range, costs, parameter_range_2=None, out_root: Optional[str] = None) -> dict:
    """What `generate_cost` / `generate_2d_cost` `@save` (:110 `<name>_cost.jld2` with `parameter_range costs`; :157
    `<name>_cost_2d.jld2` with `parameter_range_1 parameter_range_2 costs`), under `output/<name>/` like the reference.
    JLD2 is Julia's own container (an HDF5 dialect written by JLD2.jl) and cannot be produced without Julia, so the
    variables are written under the reference's names as `.npz` (for this mirror) and as raw little-endian Float64
    `.bin` files with a `.json` index — the form `julia/cost_curves_to_jld2.jl` turns into the `.jld2` file
    `generate_cost_plot` / `generate_2d_cost_plot` (:113-126, :160-174) `@load`.  In the Julia deployment the reference's
    own `@save` line stays and none of this is needed."""
    import json
    root = os.path.join(out_root or default_save_prefix, dataset_name)
    os.makedirs(root, exist_ok=True)
    two_d = parameter_range_2 is not None
    stem = os.path.join(root, dataset_name + ("_cost_2d" if two_d else "_cost"))
    names = (("parameter_range_1", parameter_range), ("parameter_range_2", parameter_range_2), ("costs", costs)) if two_d \
        else (("parameter_range", parameter_range), ("costs", costs))
    arrays = {k: np.asarray(v, dtype=np.float64) for k, v in names}
    if two_d:
        assert arrays["costs"].shape == (arrays["parameter_range_1"].size, arrays["parameter_range_2"].size)
    else:
        assert arrays["costs"].shape == arrays["parameter_range"].shape
    np.savez(stem + ".npz", **arrays)
    index = {"jld2": os.path.basename(stem) + ".jld2", "variables": []}
    for k, a in arrays.items():
        fn = f"{os.path.basename(stem)}.{k}.bin"
        a.ravel(order="F").astype("<f8").tofile(os.path.join(root, fn))      # column-major, as Julia reads it
        index["variables"].append({"name": k, "file": fn, "shape": list(a.shape)})
    with open(stem + ".json", "w") as fh:
        json.dump(index, fh, indent=1)
    return {"npz": stem + ".npz", "index": stem + ".json"}
