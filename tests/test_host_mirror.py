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
