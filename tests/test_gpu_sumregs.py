"""Sum-of-regularisers path (/root/reference/src/SumRegsLearningFunction.jl) through the C ABI against
oracle/sumregs.py.  Tolerances as for the TV path: bit-identical iterates in strict arithmetic, relative
L2 ≤ 1e-10 (fp64) / 1e-5 (fp32) otherwise."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sr():
    from oracle import sumregs
    return sumregs


@pytest.mark.parametrize("shape", [(64, 64, 2), (33, 17, 3), (1, 9, 1), (9, 1, 2), (2, 2, 1), (3, 5, 1), (128, 128, 1)])
def test_sumregs_denoise_is_bit_identical_to_the_oracle(bp, ctx, ctx32, sr, shape):
    M, N, O = shape
    rng = np.random.default_rng(M * 31 + N)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    x = np.array([0.03, 0.012, 0.05])
    ref = sr.sumregs_pdps(f, list(x), maxiter=60)
    u = ctx.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=60))
    assert np.array_equal(u, ref), (shape, np.abs(u - ref).max())
    uf = ctx.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=60, arith=bp.FAST, init_mode=1))
    assert rel_l2(uf, sr.sumregs_pdps(f, list(x), maxiter=60, init_mode=1)) <= 1e-10
    u32 = ctx32.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=60))
    assert np.array_equal(u32.astype(np.float32), sr.sumregs_pdps(f, list(x), maxiter=60, dtype=np.float32))
    # patch parameter m×n×3 (:62-85): three up-sampled maps
    xp = rng.uniform(0.005, 0.08, (2, 3, 3)) if M >= 2 and N >= 3 else np.tile(x, (1, 1, 1))
    from oracle import oracle as orc
    maps = [orc.patch_upsample(xp[:, :, k], M, N) for k in range(3)]
    up = ctx.sumregs_denoise(f, xp, bp.sumregs_pdps_opts(maxiter=40))
    assert np.array_equal(up, sr.sumregs_pdps(f, maps, maxiter=40)), shape


def test_sumregs_resident_and_streaming_kernels_agree(bp, ctx, ctx32, sr, monkeypatch):
    """The cluster-resident solve (one launch, one image per thread-block cluster, halo columns pushed through
    distributed shared memory) against the streaming pair (two launches per iteration) and the oracle: bit-identical
    for every cluster size, with ragged column splits, a λ-map, fp32 and fast arithmetic."""
    from oracle import oracle as orc
    M, N, O = 40, 37, 3
    rng = np.random.default_rng(5)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, (M, N, O)) * 255) / 255)
    x = np.array([0.03, 0.012, 0.05])
    xp = rng.uniform(0.005, 0.08, (2, 3, 3))
    maps = [orc.patch_upsample(xp[:, :, k], M, N) for k in range(3)]
    ref = sr.sumregs_pdps(f, list(x), maxiter=50)
    refp = sr.sumregs_pdps(f, maps, maxiter=50)
    ref32 = sr.sumregs_pdps(f, list(x), maxiter=50, dtype=np.float32)
    stream = ctx.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=50, kernel=bp.KERNEL_GENERIC))
    assert np.array_equal(stream, ref) and ctx.stats()["kernel_launches"] == 2 * 50
    for cs in (1, 2, 4, 8, 16):
        monkeypatch.setenv("BPLTV_RESIDENT_CS", str(cs))
        bp.reload_env()
        u = ctx.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=50, kernel=bp.KERNEL_RESIDENT))
        assert np.array_equal(u, ref) and ctx.stats()["kernel_launches"] == 1, cs
        up = ctx.sumregs_denoise(f, xp, bp.sumregs_pdps_opts(maxiter=50, kernel=bp.KERNEL_RESIDENT))
        assert np.array_equal(up, refp), cs
        u32 = ctx32.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=50, kernel=bp.KERNEL_RESIDENT))
        assert np.array_equal(u32.astype(np.float32), ref32), cs
        uf = ctx.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=50, kernel=bp.KERNEL_RESIDENT, arith=bp.FAST))
        assert rel_l2(uf, ref) <= 1e-10, cs
    monkeypatch.delenv("BPLTV_RESIDENT_CS")
    # a shape that does not fit on chip streams under AUTO and is refused under RESIDENT
    big = np.asfortranarray(np.round(rng.uniform(0, 1, (600, 600, 1)) * 255) / 255)
    ub = ctx.sumregs_denoise(big, x, bp.sumregs_pdps_opts(maxiter=3))
    assert ctx.stats()["kernel_launches"] == 6 and np.array_equal(ub, sr.sumregs_pdps(big, list(x), maxiter=3))
    with pytest.raises(bp.BpltvError):
        ctx.sumregs_denoise(big, x, bp.sumregs_pdps_opts(maxiter=3, kernel=bp.KERNEL_RESIDENT))


def test_sumregs_denoise_on_the_reference_dataset(bp, ctx, sr, datasets):
    t, f = datasets["cameraman_128_5"]
    x0 = np.array([0.001, 0.001, 0.001])       # α₀ of scalar_bilevel_sumregs_learn (BPLDenoising.jl:429)
    u = bp.sumregs_denoise(f, x0, ctx=ctx, maxiter=400)
    assert np.array_equal(u, sr.sumregs_pdps(f, list(x0), maxiter=400))
    # one regulariser switched on alone is the corresponding TV model: forward-only = the TV path
    # (same recursion, R_K = √18 on both sides)
    v = ctx.sumregs_denoise(f, np.array([0.08, 0.0, 0.0]), bp.sumregs_pdps_opts(maxiter=300))
    w = ctx.denoise(f, 0.08, bp.pdps_opts(maxiter=300, opnorm=18 ** 0.5))
    assert rel_l2(v, w) <= 1e-12
    # properties: denoising reduces every one of the three total variations
    big = ctx.sumregs_denoise(f, np.array([0.03, 0.03, 0.03]), bp.sumregs_pdps_opts(maxiter=1500))
    for kind in sr.KINDS:
        def tv(a):
            d1, d2 = sr._grad(kind, a)
            return np.sqrt(d1 * d1 + d2 * d2).sum()
        assert tv(big[:, :, 0]) < 0.6 * tv(f[:, :, 0]), kind
    with pytest.raises(ValueError):
        ctx.sumregs_denoise(f, np.array([0.1, 0.1]))
    with pytest.raises(bp.BpltvError):
        ctx.sumregs_denoise(f, np.array([0.1, -0.1, 0.1]))
    with pytest.raises(bp.BpltvError):
        ctx.sumregs_denoise(f, x0, bp.sumregs_pdps_opts(rho=0.5, maxiter=2))


# ---- adjoint solve / λ-gradient -------------------------------------------------------------
def _crop(datasets, name, n, k=1, off=40):
    t, f = datasets[name]
    return (np.asfortranarray(t[off:off + n, off:off + n, :k]), np.asfortranarray(f[off:off + n, off:off + n, :k]))


@pytest.mark.parametrize("variant", ["reg", "nonreg"])
@pytest.mark.parametrize("n", [16, 33, 48])
def test_sumregs_gradient_on_the_oracles_u(bp, ctx, sr, datasets, variant, n):
    """Scalar sumregs_gradient (:264-327) / sumregs_gradient_reg (:112-167) on the ORACLE's u: ≤ 1e-10 against
    the compliance-form CPU checker (same algorithm), ≤ 1e-9 (reg) / 1e-6 (non-reg) against the refined literal
    sparse solve — the bars of the TV path."""
    t, f = _crop(datasets, "faces_train_128_10", n, k=2, off=30)
    x = np.array([0.03, 0.02, 0.04])
    u = sr.sumregs_pdps(f, list(x), maxiter=400)
    ctx.set_dataset((t, f))
    g = ctx.sumregs_gradient(x, u, regularised=(variant == "reg"))
    assert g.shape == (3,)
    du = sum(sr.sumregs_gradient_dual(variant, x, u[:, :, i], t[:, :, i]) for i in range(2))
    assert np.all(np.abs(g - du) <= 1e-10 * np.abs(du)), (g, du)
    if variant == "reg":
        lit = sum(sr.sumregs_gradient_reg(x, u[:, :, i], t[:, :, i], refine=3) for i in range(2))
        assert np.all(np.abs(g - lit) <= 1e-9 * np.abs(lit)), (g, lit)
    else:
        lit = sum(sr.sumregs_gradient(x, u[:, :, i], t[:, :, i], refine=3) for i in range(2))
        assert np.all(np.abs(g - lit) <= 1e-6 * np.abs(lit)), (g, lit)


def test_sumregs_gradient_with_flat_regions_and_patch_parameters(bp, ctx, sr, datasets):
    from oracle import oracle as orc
    # the circle truth is piecewise constant: its denoised image has exactly flat pixels (active sets)
    t, f = _crop(datasets, "circle_128_10", 32, off=48)
    x = np.array([0.05, 0.04, 0.06])
    u = sr.sumregs_pdps(f, list(x), maxiter=600)
    u[:8, :8, 0] = u[0, 0, 0]                         # force an exactly flat block (|∇_k u| = 0 for all k)
    ctx.set_dataset((t, f))
    for variant in ("reg", "nonreg"):
        g = ctx.sumregs_gradient(x, u, regularised=(variant == "reg"))
        du = sr.sumregs_gradient_dual(variant, x, u[:, :, 0], t[:, :, 0])
        assert np.all(np.abs(g - du) <= 1e-9 * np.abs(du).max()), (variant, g, du)
    # patch parameter, non-regularised branch (:330-407)
    xp = np.stack([np.array([[0.03, 0.05], [0.02, 0.04]]) * s for s in (1.0, 0.7, 1.3)], axis=2)
    maps = [orc.patch_upsample(xp[:, :, k], 32, 32) for k in range(3)]
    gp = ctx.sumregs_gradient(xp, u, regularised=False)
    dp = sr.sumregs_gradient_dual("nonreg", maps, u[:, :, 0], t[:, :, 0], grid_shape=(2, 2))
    assert gp.shape == (2, 2, 3) and np.all(np.abs(gp - dp) <= 1e-9 * np.abs(dp).max()), (gp, dp)
    # patch regularised branch (:195-262, γ = 1e8): the row-scaled, non-symmetric system, solved by the
    # node-space band LU (lu_band.cuh); against the refined literal sparse solve, the 1e-9 bar of the other
    # regularised variants
    gr = ctx.sumregs_gradient(xp, u, regularised=True)
    lr = sr.sumregs_gradient_reg(maps, u[:, :, 0], t[:, :, 0], grid_shape=(2, 2), refine=3)
    assert gr.shape == (2, 2, 3) and np.all(np.abs(gr - lr) <= 1e-9 * np.abs(lr).max()), (gr, lr)


def test_patch_sumregs_gradient_reg_band_lu(bp, ctx, sr, datasets):
    """Patch variant of sumregs_gradient_reg (:195-262) through the band LU: several images per call (one
    CTA each, summed in image order, :183-191), odd size (short last block, ragged tiles), a 3×2 grid whose
    patches do not divide the image, γ override, fp32 context, and the full 128×128 dataset size."""
    from oracle import oracle as orc
    # three images, 40×40, 2×2 grid
    t, f = _crop(datasets, "faces_train_128_10", 40, k=3, off=30)
    xp = np.stack([np.array([[0.03, 0.05], [0.02, 0.04]]) * s for s in (1.0, 0.7, 1.3)], axis=2)
    maps = [orc.patch_upsample(xp[:, :, k], 40, 40) for k in range(3)]
    u = sr.sumregs_pdps(f, maps, maxiter=300)
    ctx.set_dataset((t, f))
    g = ctx.sumregs_gradient(xp, u, regularised=True)
    lit = sum(sr.sumregs_gradient_reg(maps, u[:, :, i], t[:, :, i], grid_shape=(2, 2), refine=3) for i in range(3))
    assert np.all(np.abs(g - lit) <= 1e-9 * np.abs(lit).max()), (g, lit)
    # the learning function takes this branch for Δ ≤ Δt = 1e-3 (:29-33)
    u2, cost2, g2 = bp.sumregs_learning_function(xp, (t, f), 5e-4, ctx=ctx, maxiter=300)
    assert np.array_equal(u2, u) and np.allclose(g2, g, rtol=1e-13, atol=0)
    # γ through the options (1e3 makes the system benign: 1e-12)
    g3 = ctx.sumregs_gradient(xp, u, regularised=True, opts=bp.sumregs_eval_opts(gamma_patch=1e3))
    lit3 = sum(sr.sumregs_gradient_reg(maps, u[:, :, i], t[:, :, i], grid_shape=(2, 2), gamma=1e3, refine=3)
               for i in range(3))
    assert np.all(np.abs(g3 - lit3) <= 1e-12 * np.abs(lit3).max()), (g3, lit3)
    # odd size, 3×2 grid
    t, f = _crop(datasets, "cameraman_128_5", 37, off=40)
    xq = np.stack([np.array([[0.03, 0.05], [0.02, 0.04], [0.06, 0.01]]) * s for s in (1.0, 0.7, 1.3)], axis=2)
    mq = [orc.patch_upsample(xq[:, :, k], 37, 37) for k in range(3)]
    uq = sr.sumregs_pdps(f, mq, maxiter=300)
    ctx.set_dataset((t, f))
    gq = ctx.sumregs_gradient(xq, uq, regularised=True)
    lq = sr.sumregs_gradient_reg(mq, uq[:, :, 0], t[:, :, 0], grid_shape=(3, 2), refine=3)
    assert gq.shape == (3, 2, 3) and np.all(np.abs(gq - lq) <= 1e-9 * np.abs(lq).max()), (gq, lq)
    # fp32 context: the solve runs in fp32, the gradient in fp64 on the fp32 u (its own oracle input)
    with bp.Context([0], 32) as c32:
        c32.set_dataset((t, f))
        u32 = c32.sumregs_denoise(f, xq, bp.sumregs_pdps_opts(maxiter=300))
        g32 = c32.sumregs_gradient(xq, u32, regularised=True)
        t32 = t.astype(np.float32).astype(np.float64)
        m32 = [m.astype(np.float32).astype(np.float64) for m in mq]
        l32 = sr.sumregs_gradient_reg(m32, u32[:, :, 0].astype(np.float64), t32[:, :, 0], grid_shape=(3, 2), refine=3)
        assert np.all(np.abs(g32 - l32) <= 1e-9 * np.abs(l32).max()), (g32, l32)


def test_patch_sumregs_gradient_reg_at_dataset_size(bp, ctx, sr, datasets):
    """128×128 (the reference's dataset size; 16 384 unknowns, half-bandwidth 256) at α₀ (BPLDenoising.jl:462)."""
    from oracle import oracle as orc
    t, f = datasets["circle_128_10"]
    x0 = 0.001 * np.ones((2, 2, 3)); x0[1, 0, :] *= 1.5; x0[0, 1, 2] *= 0.5
    maps = [orc.patch_upsample(x0[:, :, k], 128, 128) for k in range(3)]
    ctx.set_dataset((t, f))
    u = ctx.sumregs_denoise(f, x0, bp.sumregs_pdps_opts(maxiter=500))
    g = ctx.sumregs_gradient(x0, u, regularised=True)
    st = ctx.stats()
    lit = sr.sumregs_gradient_reg(maps, u[:, :, 0], t[:, :, 0], grid_shape=(2, 2), refine=3)
    assert np.all(np.abs(g - lit) <= 1e-9 * np.abs(lit).max()), (g, lit)
    assert st["ms_gradient"] > 0
    print("patch sumregs_gradient_reg 128x128: %.1f ms" % st["ms_gradient"])


def test_sumregs_learning_function_end_to_end(bp, ctx, sr, datasets):
    """sumregs_learning_function(x, data, Δ) (:8-20) on the reference's cameraman dataset at α₀ (BPLDenoising.jl:429)."""
    from oracle import oracle as orc
    t, f = datasets["cameraman_128_5"]
    x0 = np.array([0.001, 0.001, 0.001])
    u, cost, g = bp.sumregs_learning_function(x0, (t, f), 0.01, ctx=ctx, maxiter=300)     # Δ₀ = 0.01 > Δt
    ref = sr.sumregs_pdps(f, list(x0), maxiter=300)
    assert np.array_equal(u, ref) and abs(cost - orc.cost(ref, t)) <= 1e-12 * cost
    assert g.shape == (3,) and np.all(np.isfinite(g))
    st = ctx.stats()
    # resident solve, cost, gradient (nested dissection: classification, sizes, stencil, one launch per tree level for
    # the factorisation and two per solve, residuals, functional)
    assert st["ms_gradient"] > 0 and 1 + 2 + 5 < st["kernel_launches"] <= 120
    assert 0 < st["solver_max_relres"] < 1.0                                 # |r|/|b| of the banded adjoint solve (informational)
    u2, cost2, g2 = bp.sumregs_learning_function(x0, (t, f), 1e-4, ctx=ctx, maxiter=300)  # Δ ≤ Δt: regularised
    assert np.array_equal(u2, u) and cost2 == cost and not np.allclose(g, g2)
    assert 0 < ctx.stats()["solver_max_relres"] < 1.0
    lit = sr.sumregs_gradient_reg(x0, ref[:, :, 0], t[:, :, 0], refine=3)
    assert np.all(np.abs(g2 - lit) <= 1e-9 * np.abs(lit)), (g2, lit)
    with pytest.raises(bp.BpltvError):
        ctx.sumregs_learn_eval(np.array([0.1, 0.0, 0.1]), 0.1)          # λ must be > 0 for the gradient


def test_scalar_bilevel_sumregs_learn_run(bp, ctx, datasets):
    """scalar_bilevel_sumregs_learn (BPLDenoising.jl:432-451) through the host restatement of the trust-region
    driver: the learned 3-vector stays positive and lowers the upper-level cost."""
    from bpldenoising_b200 import trbox
    t, f = _crop(datasets, "cameraman_128_5", 64, off=32)
    res = trbox.scalar_bilevel_sumregs_learn((t, f), ctx=ctx, maxiter=6)
    assert np.shape(res.x) == (3,) and np.all(np.asarray(res.x) > 0)
    assert res.evaluations == len(res.log) + 1
    c0 = bp.sumregs_learning_function(np.array([0.001, 0.001, 0.001]), (t, f), 0.01, ctx=ctx)[1]
    assert res.log[-1].function_value <= c0


def test_scalar_sumregs_gradient_reg_nested_dissection_vs_band_lu(bp, ctx, ctx32, sr, datasets):
    """Scalar sumregs_gradient_reg (:112-167, γ = 1e3) takes the nested-dissection multifrontal Cholesky at coupling
    radius 2 (nd_sumregs.cuh; eval_opts.solver 0) — against the node-space band LU (solver 1), an independent second
    implementation, and against the refined literal sparse solve: 1e-9 (the bar of the regularised variants); the
    solver's backward error is reported.  Several images, an odd size, the reference's 128², an fp32 context."""
    import os
    x = np.array([0.03, 0.02, 0.04])
    for n, k, off in ((37, 3, 30), (128, 2, 0)):
        t, f = _crop(datasets, "faces_train_128_10", n, k=k, off=off)
        u = np.asfortranarray(ctx.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=300)))
        u[2:9, 3:8, 0] = u[2, 3, 0]                     # exactly flat pixels: the γ·I tensors
        ctx.set_dataset((t, f))
        g_nd = ctx.sumregs_gradient(x, u, regularised=True)
        st = ctx.stats()
        assert st["solver_max_relres"] <= 1e-14, st
        launches_nd = st["kernel_launches"]
        os.environ["BPLTV_GRAD_SOLVER"] = "1"
        bp.reload_env()
        try:
            g_lu = ctx.sumregs_gradient(x, u, regularised=True)
            launches_lu = ctx.stats()["kernel_launches"]
        finally:
            del os.environ["BPLTV_GRAD_SOLVER"]
            bp.reload_env()
        assert launches_nd != launches_lu                # two different paths did run
        assert np.all(np.abs(g_nd - g_lu) <= 1e-10 * np.abs(g_lu).max()), (n, g_nd, g_lu)
        if n <= 48:
            lit = sum(sr.sumregs_gradient_reg(x, u[:, :, i], t[:, :, i], refine=3) for i in range(k))
            assert np.all(np.abs(g_nd - lit) <= 1e-9 * np.abs(lit).max()), (g_nd, lit)
        # fp32 context: the fp32 u, adjoint system in fp64 — north_star's 1e-5 on the same u
        ctx32.set_dataset((t, f))
        g32 = ctx32.sumregs_gradient(x, u.astype(np.float32).astype(np.float64), regularised=True)
        u32 = np.asfortranarray(u.astype(np.float32).astype(np.float64))
        ctx.set_dataset((t.astype(np.float32).astype(np.float64), f))
        g64 = ctx.sumregs_gradient(x, u32, regularised=True)
        assert np.all(np.abs(g32 - g64) <= 1e-5 * np.abs(g64).max()), (g32, g64)


def test_sumregs_gradient_nested_dissection_vs_band_cholesky(bp, ctx, sr, datasets):
    """sumregs_gradient (non-regularised: scalar :264-327, patch :330-407) takes the nested-dissection solver in
    multiplier space (3-6 unknowns per pixel, nd_sumregs.cuh MULT3; eval_opts.solver 0) — against the band Cholesky of
    round 1 (solver 1): 1e-9, both being ≤ 1e-10 from the CPU compliance-form checker at the sizes it finishes."""
    import os
    from oracle import oracle as orc
    x = np.array([0.03, 0.02, 0.04])
    xp = np.stack([np.array([[0.03, 0.05], [0.02, 0.04]]) * s for s in (1.0, 0.7, 1.3)], axis=2)
    for n, k, off in ((37, 3, 30), (128, 2, 0)):
        t, f = _crop(datasets, "faces_train_128_10", n, k=k, off=off)
        u = np.asfortranarray(ctx.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=500)))
        u[2:9, 3:8, 0] = u[2, 3, 0]                     # exactly flat pixels: two modes per operator
        ctx.set_dataset((t, f))
        g_nd = ctx.sumregs_gradient(x, u, regularised=False)
        st = ctx.stats()
        gp_nd = ctx.sumregs_gradient(xp, u, regularised=False)
        os.environ["BPLTV_GRAD_SOLVER"] = "1"
        bp.reload_env()
        try:
            g_b = ctx.sumregs_gradient(x, u, regularised=False)
            stb = ctx.stats()
            gp_b = ctx.sumregs_gradient(xp, u, regularised=False)
        finally:
            del os.environ["BPLTV_GRAD_SOLVER"]
            bp.reload_env()
        assert st["kernel_launches"] != stb["kernel_launches"]          # two different paths did run
        assert st["solver_max_relres"] <= 1e-9, st
        assert np.all(np.abs(g_nd - g_b) <= 1e-9 * np.abs(g_b).max()), (n, g_nd, g_b)
        assert gp_nd.shape == (2, 2, 3) and np.all(np.abs(gp_nd - gp_b) <= 1e-9 * np.abs(gp_b).max()), (n, gp_nd, gp_b)
        if n <= 48:
            du = sum(sr.sumregs_gradient_dual("nonreg", x, u[:, :, i], t[:, :, i]) for i in range(k))
            assert np.all(np.abs(g_nd - du) <= 1e-10 * np.abs(du).max()), (g_nd, du)
            maps = [orc.patch_upsample(xp[:, :, kk], n, n) for kk in range(3)]
            dp = sum(sr.sumregs_gradient_dual("nonreg", maps, u[:, :, i], t[:, :, i], grid_shape=(2, 2)) for i in range(k))
            assert np.all(np.abs(gp_nd - dp) <= 1e-10 * np.abs(dp).max()), (gp_nd, dp)


def test_sumregs_gradient_fallback_and_wave_retry(bp, sr, datasets):
    """Two data-dependent branches of the multiplier-form driver (gradient_nd.cuh, run_gradient3_nd_mult): a front beyond
    what the path takes sends the call to the band Cholesky (BPLTV_ND3_MAXF forces it), and a wave whose measured pools
    exceed the memory budget is repeated with fewer images (BPLTV_ND3_BUDGET_MB forces it) — same gradients either way."""
    import os
    x = np.array([0.03, 0.02, 0.04])
    t, f = _crop(datasets, "faces_train_128_10", 40, k=5, off=20)
    with bp.Context([0], 64) as c:
        c.set_dataset((t, f))
        u = np.asfortranarray(c.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=300)))
        ref = c.sumregs_gradient(x, u, regularised=False)
        launches = c.stats()["kernel_launches"]
        os.environ["BPLTV_GRAD_SOLVER"] = "1"
        bp.reload_env()
        band = c.sumregs_gradient(x, u, regularised=False)
        del os.environ["BPLTV_GRAD_SOLVER"]
        os.environ["BPLTV_ND3_MAXF"] = "100"
        bp.reload_env()
        try:
            fb = c.sumregs_gradient(x, u, regularised=False)
            assert np.array_equal(fb, band) and c.stats()["kernel_launches"] < launches
        finally:
            del os.environ["BPLTV_ND3_MAXF"]
            bp.reload_env()
    os.environ["BPLTV_ND3_BUDGET_MB"] = "40"       # room for one or two 40×40 images per wave, not five
    bp.reload_env()
    try:
        with bp.Context([0], 64) as c:
            c.set_dataset((t, f))
            small = c.sumregs_gradient(x, u, regularised=False)
            assert c.stats()["kernel_launches"] > launches          # several waves
    finally:
        del os.environ["BPLTV_ND3_BUDGET_MB"]
        bp.reload_env()
    assert np.all(np.abs(small - ref) <= 1e-12 * np.abs(ref).max()), (small, ref)


def test_sumregs_gradient_beyond_the_sixteen_column_panel(bp, sr, datasets):
    """Fronts whose 16-column panel does not fit in shared memory take 8-column block steps (nd_factor8_*): forced on a 48²
    crop (BPLTV_ND3_SMEM_KB) — the default factorisation's gradient to 1e-10 — and met for real on a 256×256 image, which the
    banded solver refuses (n ≤ 136): finite, backward error at rounding level, the nested-dissection launch count."""
    import os
    x = np.array([0.03, 0.02, 0.04])
    t, f = _crop(datasets, "faces_train_128_10", 48, k=2, off=20)
    with bp.Context([0], 64) as c:
        c.set_dataset((t, f))
        u = np.asfortranarray(c.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=300)))
        ref = c.sumregs_gradient(x, u, regularised=False)
        os.environ["BPLTV_ND3_SMEM_KB"] = "40"
        bp.reload_env()
        try:
            g8 = c.sumregs_gradient(x, u, regularised=False)
        finally:
            del os.environ["BPLTV_ND3_SMEM_KB"]
            bp.reload_env()
        assert np.all(np.abs(g8 - ref) <= 1e-10 * np.abs(ref).max()), (g8, ref)
    t, f = bp.synthetic_dataset(256, 256, 1, seed=3)
    with bp.Context([0], 64) as c:
        c.set_dataset((t, f))
        u = np.asfortranarray(c.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=500)))
        g = c.sumregs_gradient(x, u, regularised=False)
        st = c.stats()
        assert np.all(np.isfinite(g)) and st["solver_max_relres"] <= 1e-9 and st["kernel_launches"] > 40, (g, st)
        gr = c.sumregs_gradient(x, u, regularised=True)
        assert np.all(np.isfinite(gr)) and np.all(np.sign(gr) == np.sign(g))


def test_cluster_factorisation_is_invisible(bp, ctx, sr, datasets):
    """The banded Cholesky shared by a thread-block cluster (2, 4, 8 CTAs per image) gives bit-identical
    gradients to the single-CTA factorisation, for the TV and the sum-of-regularisers systems."""
    import os
    t, f = _crop(datasets, "faces_train_128_10", 40, k=3, off=20)
    ctx.set_dataset((t, f))
    x3 = np.array([0.03, 0.02, 0.04])
    u3 = sr.sumregs_pdps(f, list(x3), maxiter=200)
    utv = ctx.denoise(f, 0.07, bp.pdps_opts(maxiter=300))
    ref = {}
    for C in ("1", "2", "4", "8", "16"):
        os.environ["BPLTV_GRAD_CLUSTER"] = C
        bp.reload_env()
        try:
            got = (ctx.gradient(0.07, utv, False), ctx.gradient(0.07, utv, True),
                   ctx.sumregs_gradient(x3, u3, False), ctx.sumregs_gradient(x3, u3, True))
        finally:
            del os.environ["BPLTV_GRAD_CLUSTER"]
            bp.reload_env()
        if C == "1":
            ref = got
        else:
            assert got[0] == ref[0] and got[1] == ref[1], C
            # the sum-of-regularisers assembly is deterministic too (its atomics add at most two terms per entry)
            assert np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3]), C


def test_sumregs_single_process_multi_device_context(bp, datasets):
    """The sum-of-regularisers evaluation shards over the devices of a multi-device context like the TV one
    (images in contiguous blocks, [cost, grad] summed on the host in device order).  Needs ≥ 2 GPUs."""
    import ctypes
    cuda = ctypes.CDLL("libcuda.so.1")
    n = ctypes.c_int(0)
    cuda.cuInit(0); cuda.cuDeviceGetCount(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    t, f = _crop(datasets, "faces_train_128_10", 32, k=3, off=24)
    x = np.array([0.02, 0.03, 0.01])
    with bp.Context([0], 64) as c1, bp.Context([0, 1], 64) as c2:
        c1.set_dataset((t, f)); c2.set_dataset((t, f))
        eo = bp.sumregs_eval_opts(bp.sumregs_pdps_opts(maxiter=200))
        for Delta in (0.01, 1e-4):
            u1, cost1, g1 = c1.sumregs_learn_eval(x, Delta, eo)
            u2, cost2, g2 = c2.sumregs_learn_eval(x, Delta, eo)
            assert np.array_equal(u1, u2) and abs(cost1 - cost2) <= 1e-13 * cost1
            assert np.allclose(g1, g2, rtol=1e-12, atol=0)   # host sum over devices vs one in-order device sum
        assert c2.stats()["n_devices"] == 2
        assert np.array_equal(c2.sumregs_denoise(f, x, eo.pdps), u1)


def test_patch_bilevel_sumregs_learn_runs_through_both_branches(bp, ctx, datasets):
    """patch_bilevel_sumregs_learn (BPLDenoising.jl:464-481): m×n×3 parameter through the L-BFGS path of the
    driver, in the non-regularised branch (Δ₀ = 0.1 > Δt) and, started below Δt = 1e-3, in the regularised
    patch branch (band LU)."""
    from bpldenoising_b200 import trbox
    t, f = _crop(datasets, "circle_128_10", 32, off=48)
    res = trbox.patch_bilevel_sumregs_learn((t, f), ctx=ctx, maxiter=3)
    assert np.shape(res.x) == (2, 2, 3) and np.all(np.asarray(res.x) > 0) and res.evaluations == 4
    res2 = trbox.patch_bilevel_sumregs_learn((t, f), ctx=ctx, maxiter=2, Delta0=5e-4)
    assert np.shape(res2.x) == (2, 2, 3) and np.all(np.asarray(res2.x) > 0) and np.all(np.isfinite(res2.x))
