"""Host-side mirror of the reference interface: argument handling that needs no GPU."""
import numpy as np
import pytest


def test_lambda_and_stack_normalisation(bp):
    from bpldenoising_b200 import learning as L
    lam, scalar = L._lam(0.1)
    assert scalar and lam.shape == (1, 1)
    lam, scalar = L._lam(np.ones((2, 3)))
    assert not scalar and lam.flags.f_contiguous
    with pytest.raises(ValueError):
        L._lam(np.ones(3))
    s = L._stack(np.zeros((4, 5)))
    assert s.shape == (4, 5, 1) and s.flags.f_contiguous
    with pytest.raises(ValueError):
        L._stack(np.zeros(4))
    assert L.L2CostFunction(np.ones((2, 2)), np.zeros((2, 2))) == 2.0


def test_synthetic_dataset_is_frozen(bp):
    t, f = bp.synthetic_dataset(32, 24, 3, seed=20240601)
    t2, f2 = bp.synthetic_dataset(32, 24, 3, seed=20240601)
    assert np.array_equal(t, t2) and np.array_equal(f, f2)
    assert t.shape == (32, 24, 3) and f.flags.f_contiguous
    assert np.allclose(f * 255, np.round(f * 255)) and f.min() >= 0 and f.max() <= 1
    assert 0.05 < np.std(f - t) < 0.15


def test_dataset_loader_follows_the_reference_format(bp, datasets, tmp_path):
    """filelist.txt + 8-bit / 1-bit PNG pairs → k/255 stacks (/root/reference/src/Datasets.jl:54-65)."""
    from PIL import Image
    for name in ("faces_val_128_10", "circle_128_10"):
        t, d = datasets[name]
        root = tmp_path / "BPLDenoising" / "datasets" / name
        root.mkdir(parents=True)
        lines = []
        for i in range(min(3, t.shape[2])):
            if set(np.unique(t[:, :, i])) <= {0.0, 1.0}:   # the circle truth is a 1-bit PNG
                Image.fromarray(t[:, :, i].astype(bool)).save(root / f"t{i}.png")
            else:
                Image.fromarray(np.round(t[:, :, i] * 255).astype(np.uint8)).save(root / f"t{i}.png")
            Image.fromarray(np.round(d[:, :, i] * 255).astype(np.uint8)).save(root / f"d{i}.png")
            lines.append(f"t{i}.png,d{i}.png")
        (root / "filelist.txt").write_text("\n".join(lines) + "\n")
        tt, dd = bp.testdataset(name[:8], dataset_dir=str(tmp_path / "BPLDenoising" / "datasets"))
        k = len(lines)
        assert tt.flags.f_contiguous and tt.shape == (128, 128, k)
        assert np.array_equal(tt, t[:, :, :k]) and np.array_equal(dd, d[:, :, :k])
    with pytest.raises(ValueError):
        bp.datasets.full_datasetname("no_such_dataset")
    assert bp.datasets.full_datasetname("cameraman") == "cameraman_128_5"   # first prefix match, like findfirst


def test_quality_indexes(bp, datasets):
    q = bp.quality
    t, d = datasets["cameraman_128_5"]
    a, b = t[:, :, 0], d[:, :, 0]
    mse = np.mean((a - b) ** 2)
    assert abs(q.assess_psnr(b, a) - (-10 * np.log10(mse))) < 1e-12
    assert abs(q.psnr_from_sqerr(np.sum((a - b) ** 2), a.size) - q.assess_psnr(a, b)) < 1e-12
    assert q.assess_psnr(a, a) == float("inf")
    assert abs(q.assess_ssim(a, a) - 1.0) < 1e-12
    s = q.assess_ssim(a, b)
    assert 0.0 < s < 1.0 and abs(s - q.assess_ssim(b, a)) < 1e-12        # symmetric, degraded by noise
    # noisier input scores lower on both indexes
    rng = np.random.default_rng(0)
    c = np.clip(a + 0.2 * rng.standard_normal(a.shape), 0, 1)
    assert q.assess_ssim(a, c) < s and q.assess_psnr(a, c) < q.assess_psnr(a, b)


def test_save_results_writes_the_reference_artefacts(bp, datasets, tmp_path):
    """save_results (/root/reference/src/BPLDenoising.jl:185-258): log, quality table, PNGs, parameter map."""
    from PIL import Image
    from bpldenoising_b200 import results, trbox
    t, d = datasets["faces_val_128_10"]
    t, d = t[:, :, :2], d[:, :, :2]
    reco = 0.5 * (t + d)
    log = [trbox.LogEntry(1, 0.1, 3.0, 2.0, 0.1, 0.0), trbox.LogEntry(2, 0.2, 2.5, 1.0, 0.05, 0.0)]
    prm = dict(dataset_name="faces_val_128_10", save_prefix="tv_optimal_parameter_scalar_faces_val_128_10", save_results=True)
    w = results.save_results(prm, t, d, 0.07, reco, log, out_root=str(tmp_path))
    lines = open(w["quality"]).read().splitlines()
    assert lines[0].split() == ["img_num", "orig_ssim", "orig_psnr", "out_ssim", "out_psnr"] and len(lines) == 4
    r1 = [float(v) for v in lines[1].split()]
    assert r1[0] == 1 and abs(r1[2] - bp.quality.assess_psnr(t[:, :, 0], d[:, :, 0])) < 1e-12 and r1[4] > r1[2]
    assert abs(float(lines[3].split()[1]) - w["mean_psnr"]) < 1e-12
    assert len(w["png"]) == 6 and open(w["log"]).read().count("\n") == 4
    back = np.asarray(Image.open([p for p in w["png"] if p.endswith("_true_2.png")][0]), dtype=np.float64) / 255.0
    assert np.array_equal(back, t[:, :, 1])                      # 8-bit round trip of k/255 data
    w2 = results.save_results(dict(prm, save_prefix="patch"), t, d, np.array([[0.1, 0.2], [0.3, 0.5]]), reco, log,
                              out_root=str(tmp_path))
    par = np.asarray(Image.open([p for p in w2["png"] if p.endswith("_par.png")][0]))
    assert par.shape == (128, 128) and par[0, 0] == 0 and par[127, 127] == 255 and par[0, 127] == 64
    assert results.save_results(dict(prm, save_results=False), t, d, 0.07, reco, log, out_root=str(tmp_path)) == {}
    # m×n×3 parameter (:260-299): three jointly stretched maps; the file's mean PSNR is the reference's 0.0 (:282)
    x3 = np.stack([np.array([[0.1, 0.2], [0.3, 0.5]]) * s for s in (1.0, 0.5, 2.0)], axis=2)
    w3 = results.save_results(dict(prm, save_prefix="sumregs_patch"), t, d, x3, reco, log, out_root=str(tmp_path))
    pars = [np.asarray(Image.open(p)) for p in w3["png"] if "_par_" in p]
    assert len(pars) == 3 and pars[2][127, 127] == 255 and pars[1][0, 0] == 0 and pars[0][127, 127] < 255
    assert float(open(w3["quality"]).read().splitlines()[3].split()[1]) == 0.0 and w3["mean_psnr"] > 0


def test_cost_curves_are_saved_under_the_reference_names(tmp_path):
    """generate_cost / generate_2d_cost `@save` (/root/reference/src/BPLDenoising.jl:110, :157): the same variables under
    the same names, as .npz plus the raw column-major Float64 files julia/cost_curves_to_jld2.jl converts to .jld2."""
    import json
    from bpldenoising_b200 import results
    pr = np.geomspace(1e-3, 1.0, 7)
    costs = 1.0 / pr
    w = results.save_cost_curve("cameraman_128_5", pr, costs, out_root=str(tmp_path))
    z = np.load(w["npz"])
    assert sorted(z.files) == ["costs", "parameter_range"] and np.array_equal(z["costs"], costs)
    idx = json.load(open(w["index"]))
    assert idx["jld2"] == "cameraman_128_5_cost.jld2" and [v["name"] for v in idx["variables"]] == ["parameter_range", "costs"]
    raw = np.fromfile(tmp_path / "cameraman_128_5" / idx["variables"][1]["file"], dtype="<f8")
    assert np.array_equal(raw, costs)
    p1, p2 = np.array([0.1, 0.2, 0.3]), np.array([1.0, 2.0])
    c2 = np.arange(6.0).reshape(3, 2)
    w2 = results.save_cost_curve("circle_128_10", p1, c2, parameter_range_2=p2, out_root=str(tmp_path))
    idx2 = json.load(open(w2["index"]))
    assert idx2["jld2"] == "circle_128_10_cost_2d.jld2"
    raw2 = np.fromfile(tmp_path / "circle_128_10" / idx2["variables"][2]["file"], dtype="<f8")
    assert np.array_equal(raw2.reshape((3, 2), order="F"), c2)          # column-major on disk
