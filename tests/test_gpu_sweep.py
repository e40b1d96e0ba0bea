"""λ-sweeps and validation through the C ABI (`bpltv_sweep`): all parameter sets × images as one
batch, against the oracle's one-solve-at-a-time loop (the reference's generate_cost /
generate_2d_cost / validate_tv_parameter, /root/reference/src/BPLDenoising.jl:92-158, :381-415)."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _ref_costs(oracle, t, f, params, maxiter, upsample=None):
    us, costs = [], []
    for p in params:
        alpha = p if np.ndim(p) == 0 else oracle.patch_upsample(np.asarray(p, dtype=float), *f.shape[:2])
        u = oracle.pdps(f, alpha, maxiter=maxiter)
        us.append(u)
        costs.append(oracle.cost(u, t))
    return np.array(costs), us


@pytest.mark.parametrize("kernel", ["auto", "generic", "march", "tblock", "resident"])
def test_scalar_sweep_matches_the_loop(bp, ctx, oracle, datasets, kernel):
    t, f = datasets["faces_train_128_10"]
    t, f = np.asfortranarray(t[:, :, :3]), np.asfortranarray(f[:, :, :3])
    ctx.set_dataset((t, f))
    params = [0.02, 0.05, 0.1, 0.2, 0.0]
    kid = dict(auto=bp.KERNEL_AUTO, generic=bp.KERNEL_GENERIC, march=bp.KERNEL_MARCH, tblock=bp.KERNEL_TBLOCK,
               resident=bp.KERNEL_RESIDENT)[kernel]
    costs, sq, u = ctx.sweep(params, bp.pdps_opts(maxiter=150, kernel=kid), return_u=True, return_sqerr=True)
    rc, rus = _ref_costs(oracle, t, f, params, 150)
    for l in range(len(params)):
        assert np.array_equal(u[:, :, :, l], rus[l]), (kernel, l)
    assert np.allclose(costs, rc, rtol=1e-12, atol=0)
    assert np.allclose(0.5 * sq.sum(axis=0), costs, rtol=1e-14)
    assert sq.shape == (3, len(params))
    assert np.allclose(sq[1, 2], np.sum((rus[2][:, :, 1] - t[:, :, 1]) ** 2), rtol=1e-12)


def test_patch_sweep_and_2d_cost(bp, ctx, oracle, datasets):
    t, f = datasets["cameraman_128_5"]
    r1, r2 = [0.03, 0.1], [0.05, 0.12, 0.2]
    costs = bp.generate_2d_tv_cost((t, f), r1, r2, num_samples=1, ctx=ctx, maxiter=120)
    assert costs.shape == (2, 3)
    for i, a in enumerate(r1):
        for j, b in enumerate(r2):
            alpha = oracle.patch_upsample(np.array([[a], [b]]), 128, 128)
            ref = oracle.cost(oracle.pdps(f, alpha, maxiter=120), t)
            assert abs(costs[i, j] - ref) <= 1e-12 * ref, (i, j)
    # general patch grids, ragged image, streaming kernels
    rng = np.random.default_rng(5)
    tt = np.asfortranarray(np.round(rng.uniform(0, 1, (66, 40, 2)) * 255) / 255)
    ff = np.asfortranarray(np.clip(tt + 0.1 * rng.standard_normal(tt.shape), 0, 1))
    ctx.set_dataset((tt, ff))
    grids = [rng.uniform(0.01, 0.2, (3, 2)) for _ in range(4)]
    for kid in (bp.KERNEL_GENERIC, bp.KERNEL_MARCH, bp.KERNEL_TBLOCK):
        c, u = ctx.sweep(grids, bp.pdps_opts(maxiter=60, kernel=kid), return_u=True)
        for l, gr in enumerate(grids):
            ref = oracle.pdps(ff, oracle.patch_upsample(gr, 66, 40), maxiter=60)
            assert np.array_equal(u[:, :, :, l], ref), (kid, l)
            assert abs(c[l] - oracle.cost(ref, tt)) <= 1e-12 * c[l]


def test_scalar_cost_curve_full_length(bp, ctx, oracle, datasets):
    # generate_scalar_tv_cost: TVDenoise = 10000 iterations (/root/reference/src/BPLDenoising.jl:51)
    t, f = datasets["cameraman_128_5"]
    rng_ = np.geomspace(0.005, 0.5, 24)
    costs = bp.generate_scalar_tv_cost((t, f), rng_, num_samples=1, ctx=ctx)
    assert ctx.stats()["pdps_iterations"] == 10000 and ctx.stats()["kernel_launches"] <= 4
    for l in (0, 11, 23):
        ref = oracle.cost(oracle.pdps(f, float(rng_[l]), maxiter=10000), t)
        assert abs(costs[l] - ref) <= 1e-12 * ref
    k = int(np.argmin(costs))
    assert 0 < k < 23, "the cost curve has an interior minimum on the reference's dataset"


def test_sweep_init_mode_fp32_and_errors(bp, ctx, ctx32, oracle, datasets):
    t, f = datasets["faces_val_128_10"]
    t, f = np.asfortranarray(t[:64, :48, :2]), np.asfortranarray(f[:64, :48, :2])
    ctx.set_dataset((t, f))
    for kid in (bp.KERNEL_MARCH, bp.KERNEL_RESIDENT):
        c, u = ctx.sweep([0.04, 0.09, 0.15], bp.pdps_opts(maxiter=70, init_mode=1, kernel=kid), return_u=True)
        for l, lam in enumerate([0.04, 0.09, 0.15]):
            assert np.array_equal(u[:, :, :, l], oracle.pdps(f, lam, maxiter=70, init_mode=1)), (kid, l)
    ctx32.set_dataset((t, f))
    c32, u32 = ctx32.sweep([0.04, 0.09], bp.pdps_opts(maxiter=70), return_u=True)
    for l, lam in enumerate([0.04, 0.09]):
        assert np.array_equal(u32[:, :, :, l].astype(np.float32), oracle.pdps(f, lam, maxiter=70, dtype=np.float32))
    with pytest.raises(bp.BpltvError):
        ctx.sweep([0.1, -0.1])
    with pytest.raises(ValueError):
        ctx.sweep([])
    with pytest.raises(ValueError):
        ctx.sweep([0.1, np.ones((2, 2))])
    fresh = bp.Context([0], 64)
    with pytest.raises(bp.BpltvError):
        fresh.sweep([0.1])
    fresh.close()


def test_validate_tv_parameter_table(bp, ctx, oracle, datasets):
    t, f = datasets["faces_val_128_10"]
    t, f = np.asfortranarray(t[:, :, :3]), np.asfortranarray(f[:, :, :3])
    res = bp.validate_tv_parameter(0.07, (t, f), ctx=ctx, maxiter=400)
    ref = oracle.pdps(f, 0.07, maxiter=400)
    assert np.array_equal(res["u"], ref)
    assert abs(res["cost"] - oracle.cost(ref, t)) <= 1e-12 * res["cost"]
    assert len(res["table"]) == 3
    for i, row in enumerate(res["table"]):
        assert abs(row["out_psnr"] - bp.quality.assess_psnr(t[:, :, i], ref[:, :, i])) < 1e-9
        assert row["out_psnr"] > row["orig_psnr"] and row["out_ssim"] > row["orig_ssim"]   # denoising helps
    assert abs(res["mean_psnr"] - np.mean([r["out_psnr"] for r in res["table"]])) < 1e-12


def test_sweep_on_all_devices(bp, oracle, datasets):
    try:
        c2 = bp.Context([0, 1], 64)
    except bp.BpltvError:
        pytest.skip("needs 2 GPUs")
    t, f = datasets["faces_train_128_10"]
    t, f = np.asfortranarray(t[:, :, :5]), np.asfortranarray(f[:, :, :5])
    c2.set_dataset((t, f))
    params = [0.03, 0.1, 0.3]
    costs, sq, u = c2.sweep(params, bp.pdps_opts(maxiter=90), return_u=True, return_sqerr=True)
    for l, lam in enumerate(params):
        ref = oracle.pdps(f, lam, maxiter=90)
        assert np.array_equal(u[:, :, :, l], ref)
        assert abs(costs[l] - oracle.cost(ref, t)) <= 1e-12 * costs[l]
    c2.close()


def test_sweep_through_the_streaming_kernels_in_fast_and_fp32_modes(bp, ctx, ctx32, oracle):
    """The λ-sweep instantiations of the temporally blocked kernel (T = 2 and T = 4, BATCH = true), in fast
    arithmetic and in fp32; T = 3 is not built for sweeps and says so."""
    rng = np.random.default_rng(11)
    t = np.asfortranarray(np.round(rng.uniform(0, 1, (64, 50, 3)) * 255) / 255)
    f = np.asfortranarray(np.clip(t + 0.1 * rng.standard_normal(t.shape), 0, 1))
    params = [0.03, 0.08, 0.2]
    refs = [oracle.pdps(f, p, maxiter=81) for p in params]
    ctx.set_dataset((t, f)); ctx32.set_dataset((t, f))
    for depth in (2, 4):
        for arith, tol in ((bp.STRICT, 0.0), (bp.FAST, 1e-10)):
            c, u = ctx.sweep(params, bp.pdps_opts(maxiter=81, kernel=bp.KERNEL_TBLOCK, tblock=depth, arith=arith), return_u=True)
            for l in range(3):
                assert rel_l2(u[:, :, :, l], refs[l]) <= tol, (depth, arith, l)
    c32, u32 = ctx32.sweep(params, bp.pdps_opts(maxiter=81, kernel=bp.KERNEL_TBLOCK), return_u=True)
    for l, p in enumerate(params):
        assert np.array_equal(u32[:, :, :, l].astype(np.float32), oracle.pdps(f, p, maxiter=81, dtype=np.float32))
    with pytest.raises(bp.BpltvError) as ei:
        ctx.sweep(params, bp.pdps_opts(maxiter=9, kernel=bp.KERNEL_TBLOCK, tblock=3))
    assert "depths 2 and 4" in str(ei.value)
