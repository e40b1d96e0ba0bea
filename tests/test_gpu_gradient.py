"""Parity of the CUDA adjoint solve / λ-gradient and of the whole learning function
with the CPU oracle, through the C ABI.

Tolerances.  The gradient kernels are graded on the oracle's u:
  * against the oracle's dual-form solve (same formulation, CPU): ≤ 1e-10 relative;
  * against the literal sparse system of the reference (scipy SuperLU + extended
    precision refinement): ≤ 1e-9 for gradient_reg (well posed) and ≤ 1e-6 for the
    non-regularised gradient, whose literal system is reproducible only to
    1e-6…3e-5 across LU orderings.
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return rel_l2(np.atleast_1d(a), np.atleast_1d(b))


@pytest.mark.parametrize("variant", ["reg", "nonreg"])
def test_scalar_gradient_small_vs_dual_and_literal(bp, ctx, oracle, datasets, variant):
    t, f = (a[:48, :48, :2].copy(order="F") for a in datasets["faces_train_128_10"])
    u = oracle.pdps(f, 0.06, maxiter=2500)
    ctx.set_dataset((t, f))
    g = ctx.gradient(0.06, u, regularised=(variant == "reg"))
    dual = sum(oracle.gradient_dual(variant, 0.06, u[:, :, i], t[:, :, i]) for i in range(2))
    assert _rel(g, dual) <= 1e-10
    if variant == "reg":
        lit = sum(oracle.gradient_reg_scalar(0.06, u[:, :, i], t[:, :, i], refine=3) for i in range(2))
        assert _rel(g, lit) <= 1e-9
    else:
        lit = sum(oracle.gradient_scalar(0.06, u[:, :, i], t[:, :, i], refine=4) for i in range(2))
        assert _rel(g, lit) <= 1e-6


@pytest.mark.parametrize("variant", ["reg", "nonreg"])
def test_patch_gradient_small_vs_dual_and_literal(bp, ctx, oracle, datasets, variant):
    t, f = (a[:48, :48, :1].copy(order="F") for a in datasets["circle_128_10"])
    x = np.array([[0.02, 0.05, 0.03], [0.04, 0.01, 0.06]])
    am = oracle.patch_upsample(x, 48, 48)
    u = oracle.pdps(f, am, maxiter=2500)
    ctx.set_dataset((t, f))
    g = ctx.gradient(x, u, regularised=(variant == "reg"))
    assert g.shape == x.shape
    dual = oracle.gradient_dual(variant, am, u[:, :, 0], t[:, :, 0], grid_shape=x.shape)
    assert _rel(g, dual) <= 1e-10
    if variant == "reg":
        lit = oracle.gradient_reg_patch(am, x.shape, u[:, :, 0], t[:, :, 0], refine=3)
        assert _rel(g, lit) <= 1e-9
    else:
        lit = oracle.gradient_patch(am, x.shape, u[:, :, 0], t[:, :, 0], refine=4)
        assert _rel(g, lit) <= 1e-6


def test_gradient_reg_by_both_solvers(bp, ctx, oracle, datasets, monkeypatch):
    """gradient_reg has two implementations — the multiplier-space banded Cholesky and the node-space band LU
    (default up to 128×128, BPLTV_GRAD_REG_LU overrides): both within the bars of the dual-form checker and the
    refined literal solve, scalar and patch (the patch system is row-scaled as the reference writes it, :210)."""
    t, f = (a[:40, :40, :3].copy(order="F") for a in datasets["faces_train_128_10"])
    x = np.array([[0.02, 0.05], [0.04, 0.06]])
    am = oracle.patch_upsample(x, 40, 40)
    us = oracle.pdps(f, 0.06, maxiter=1500)
    up = oracle.pdps(f, am, maxiter=1500)
    ctx.set_dataset((t, f))
    dual_s = sum(oracle.gradient_dual("reg", 0.06, us[:, :, i], t[:, :, i]) for i in range(3))
    dual_p = sum(oracle.gradient_dual("reg", am, up[:, :, i], t[:, :, i], grid_shape=x.shape) for i in range(3))
    lit_s = sum(oracle.gradient_reg_scalar(0.06, us[:, :, i], t[:, :, i], refine=3) for i in range(3))
    lit_p = sum(oracle.gradient_reg_patch(am, x.shape, up[:, :, i], t[:, :, i], refine=3) for i in range(3))
    got = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("BPLTV_GRAD_REG_LU", mode)
        bp.reload_env()
        gs = ctx.gradient(0.06, us, regularised=True)
        gp = ctx.gradient(x, up, regularised=True)
        assert _rel(gs, dual_s) <= 1e-10 and _rel(gp, dual_p) <= 1e-10, mode
        assert _rel(gs, lit_s) <= 1e-9 and _rel(gp, lit_p) <= 1e-9, mode
        got[mode] = (gs, gp)
    assert _rel(got["0"][0], got["1"][0]) <= 1e-11 and _rel(got["0"][1], got["1"][1]) <= 1e-11
    # odd image size: the forward-only band has half-width n, which the LU rounds up to even for its 16-byte accesses
    t, f = (a[10:47, 20:57, :1] for a in datasets["cameraman_128_5"])
    t, f = (np.concatenate([a, a[::-1, :, :]], axis=2).copy(order="F") for a in (t, f))     # two 37×37 images
    u = oracle.pdps(f, 0.08, maxiter=800)
    ctx.set_dataset((t, f))
    monkeypatch.setenv("BPLTV_GRAD_REG_LU", "1")
    bp.reload_env()
    g = ctx.gradient(0.08, u, regularised=True)
    dual = sum(oracle.gradient_dual("reg", 0.08, u[:, :, i], t[:, :, i]) for i in range(2))
    assert _rel(g, dual) <= 1e-10
    monkeypatch.setenv("BPLTV_GRAD_REG_LU", "0")
    bp.reload_env()
    assert _rel(ctx.gradient(0.08, u, regularised=True), dual) <= 1e-10           # the Cholesky at the odd size
    dual_n = sum(oracle.gradient_dual("nonreg", 0.08, u[:, :, i], t[:, :, i]) for i in range(2))
    assert _rel(ctx.gradient(0.08, u, regularised=False), dual_n) <= 1e-10        # and the non-regularised branch


@pytest.mark.parametrize("name,lam", [("cameraman_128_5", 0.1), ("faces_train_128_10", 0.05), ("circle_128_10", 0.02)])
def test_scalar_gradient_full_size_vs_literal(bp, ctx, oracle, datasets, name, lam):
    t, f = (a[:, :, :2].copy(order="F") for a in datasets[name])
    u = oracle.pdps(f, lam, maxiter=5000)
    ctx.set_dataset((t, f))
    O = u.shape[2]
    g = ctx.gradient(lam, u, regularised=True)
    lit = sum(oracle.gradient_reg_scalar(lam, u[:, :, i], t[:, :, i], refine=3) for i in range(O))
    assert _rel(g, lit) <= 1e-9, (g, lit)
    g = ctx.gradient(lam, u, regularised=False)
    lit = sum(oracle.gradient_scalar(lam, u[:, :, i], t[:, :, i], refine=4) for i in range(O))
    assert _rel(g, lit) <= 1e-6, (g, lit)


def test_learning_function_end_to_end(bp, ctx, oracle, datasets):
    # BASELINE config 1 (scalar λ, cameraman) through the reference-named entry point
    data = datasets["cameraman_128_5"]
    u, cost, grad = bp.tv_op_learning_function(0.1, data, 0.1, ctx=ctx)
    ou, ocost, ograd = oracle.tv_op_learning_function(0.1, data, 0.1, refine=4)
    assert np.array_equal(u, ou)                       # strict arithmetic: identical image
    assert abs(cost - ocost) <= 1e-12 * ocost
    assert _rel(grad, ograd) <= 1e-6
    # Δ ≤ Δt switches to gradient_reg (:21-25)
    _, _, g2 = bp.tv_op_learning_function(0.1, data, 1e-7, ctx=ctx)
    _, _, og2 = oracle.tv_op_learning_function(0.1, data, 1e-7, refine=3, u=ou)
    assert _rel(g2, og2) <= 1e-9
    st = ctx.stats()
    assert st["pdps_iterations"] == 5000 and st["kernel_launches"] > 0


def test_learning_function_patch_config3(bp, ctx, oracle, datasets):
    # BASELINE config 3: patch λ = 1e-4·ones(2,2) on circle_128_10, Δ₀ = 1e-4
    data = datasets["circle_128_10"]
    x = 1e-4 * np.ones((2, 2))
    u, cost, grad = bp.tv_op_learning_function(x, data, 1e-4, ctx=ctx)
    ou, ocost, ograd = oracle.tv_op_learning_function(x, data, 1e-4, refine=4)
    assert np.array_equal(u, ou)
    assert abs(cost - ocost) <= 1e-12 * ocost
    assert grad.shape == (2, 2) and _rel(grad, ograd) <= 1e-6
    _, _, g2 = bp.tv_op_learning_function(x, data, 1e-7, ctx=ctx)
    _, _, og2 = oracle.tv_op_learning_function(x, data, 1e-7, refine=3, u=ou)
    assert _rel(g2, og2) <= 1e-9


def test_gradient_errors(bp, ctx, datasets):
    t, f = datasets["cameraman_128_5"]
    ctx.set_dataset((t, f))
    with pytest.raises(bp.BpltvError):
        ctx.learn_eval(0.0, 0.1)            # λ must stay > 0 (get_bounds, TRBox.jl:160-164)
    ctx.set_dataset((t[:, :100], f[:, :100]))
    with pytest.raises(bp.BpltvError) as ei:
        ctx.learn_eval(0.1, 0.1)
    assert "square" in str(ei.value)        # reference precondition (:102)


def test_gradient_256_config5_shape(bp, ctx, oracle):
    # BASELINE config 5 image size (256×256 synthetic): band solver with the solve vector in
    # global memory (it no longer fits in shared memory), both branches, against the literal solve
    t, f = bp.synthetic_dataset(256, 256, 2, seed=20240602)
    u = ctx.denoise(f, 0.1, bp.pdps_opts(maxiter=1500))
    assert np.array_equal(u, oracle.pdps(f, 0.1, maxiter=1500))
    ctx.set_dataset((t, f))
    g = ctx.gradient(0.1, u, regularised=True)
    lit = sum(oracle.gradient_reg_scalar(0.1, u[:, :, i], t[:, :, i], refine=3) for i in range(2))
    assert _rel(g, lit) <= 1e-9, (g, lit)
    g = ctx.gradient(0.1, u, regularised=False)
    lit = sum(oracle.gradient_scalar(0.1, u[:, :, i], t[:, :, i], refine=4) for i in range(2))
    assert _rel(g, lit) <= 1e-6, (g, lit)


def test_single_process_multi_device_context(bp, oracle, datasets):
    """bpltv_create with several device ids shards the images inside the library and sums
    [cost, grad] on the host (INTEGRATION.md §2 set_devices!).  Needs ≥ 2 GPUs."""
    import ctypes
    cuda = ctypes.CDLL("libcuda.so.1")
    n = ctypes.c_int(0)
    cuda.cuInit(0); cuda.cuDeviceGetCount(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    t, f = (a[:, :, :5].copy(order="F") for a in datasets["faces_train_128_10"])
    with bp.Context([0], 64) as c1, bp.Context([0, 1], 64) as c2:
        c1.set_dataset((t, f)); c2.set_dataset((t, f))
        eo = bp.eval_opts(bp.pdps_opts(maxiter=600))
        u1, cost1, g1 = c1.learn_eval(0.07, 0.1, eo)
        u2, cost2, g2 = c2.learn_eval(0.07, 0.1, eo)
        assert np.array_equal(u1, u2)
        assert abs(cost1 - cost2) <= 1e-13 * cost1 and abs(g1 - g2) <= 1e-12 * abs(g1)
        assert c2.stats()["n_devices"] == 2


@pytest.mark.parametrize("variant", ["reg", "nonreg"])
def test_nested_dissection_vs_band_solvers(bp, ctx, oracle, datasets, variant):
    """Two independent implementations of every adjoint solve: the nested-dissection multifrontal Cholesky (default,
    eval_opts.solver = 0/2) and round 1's banded factorisations (solver = 1).  Same formulation per branch, different
    elimination order and kernels: they must agree far inside the oracle bars, scalar and patch, odd sizes included."""
    reg = variant == "reg"
    for n, O, name in ((128, 3, "faces_train_128_10"), (37, 2, "cameraman_128_5"), (64, 1, "circle_128_10")):
        t, f = (a[:n, :n, :O].copy(order="F") for a in datasets[name])
        if t.shape[2] < O:
            t, f = (np.concatenate([a, a[::-1]], axis=2).copy(order="F") for a in (t, f))
        x = np.array([[0.02, 0.05, 0.03], [0.04, 0.01, 0.06]])
        am = oracle.patch_upsample(x, n, n)
        us = oracle.pdps(f, 0.06, maxiter=1500)
        up = oracle.pdps(f, am, maxiter=1500)
        ctx.set_dataset((t, f))
        nd, band = bp.eval_opts(solver=2), bp.eval_opts(solver=1)
        gs_nd, gs_b = ctx.gradient(0.06, us, reg, nd), ctx.gradient(0.06, us, reg, band)
        st = ctx.stats()
        gp_nd, gp_b = ctx.gradient(x, up, reg, nd), ctx.gradient(x, up, reg, band)
        assert _rel(gs_nd, gs_b) <= 1e-11, (n, gs_nd, gs_b)
        assert _rel(gp_nd, gp_b) <= 1e-11, (n, gp_nd, gp_b)
        dual = sum(oracle.gradient_dual(variant, 0.06, us[:, :, i], t[:, :, i]) for i in range(t.shape[2]))
        assert _rel(gs_nd, dual) <= 1e-10
    ctx.gradient(0.06, us, reg, nd)
    assert 0 <= ctx.stats()["solver_max_relres"] <= 1e-9
    # the banded factorisations report a residual as well: |r|/|b| of their last refinement step (include/bpltv.h)
    ctx.gradient(0.06, us, reg, band)
    assert 0 < ctx.stats()["solver_max_relres"] <= 1e-8


def test_adjoint_solver_reports_failure(bp, ctx, oracle, datasets):
    """A backward error above eval_opts.solver_tol is an error, not a silently inaccurate gradient."""
    t, f = (a[:48, :48, :1].copy(order="F") for a in datasets["cameraman_128_5"])
    u = oracle.pdps(f, 0.1, maxiter=1000)
    ctx.set_dataset((t, f))
    with pytest.raises(bp.BpltvError) as ei:
        ctx.gradient(0.1, u, False, bp.eval_opts(solver_tol=1e-30))
    assert "backward error" in str(ei.value)
    assert np.isfinite(ctx.gradient(0.1, u, False, bp.eval_opts(solver_tol=0.0)))


def test_gradient_256_many_images_nested_dissection(bp, ctx, oracle):
    """BASELINE config 5's image size in a wave of many images (one grid dimension of every kernel)."""
    t, f = bp.synthetic_dataset(256, 256, 5, seed=20240602)
    u = ctx.denoise(f, 0.1, bp.pdps_opts(maxiter=1000))
    ctx.set_dataset((t, f))
    for reg in (True, False):
        g = ctx.gradient(0.1, u, regularised=reg)
        parts = [oracle.gradient_dual("reg" if reg else "nonreg", 0.1, u[:, :, i], t[:, :, i]) for i in range(5)]
        assert _rel(g, sum(parts)) <= 1e-10, (reg, g, sum(parts))


@pytest.mark.parametrize("name,lam", [("cameraman_128_5", 0.1), ("faces_train_128_10", 0.05), ("circle_128_10", 0.02)])
def test_gradient_vs_binary128(bp, ctx, oracle, datasets, name, lam):
    """The parity bar of north_star (1e-10) against what the reference's systems mean: the adjoint systems solved in
    binary128 (oracle/quad_adjoint.c; tests/test_oracle_quad.py shows the literal and the compliance form coincide there and
    that the reference's own double-rounded assembly moves `gradient` by 1e-10 … 1e-7).  Scalar and patch λ, both branches."""
    from oracle import quad
    n = 44
    t, f = (a[30:30 + n, 40:40 + n, :2].copy(order="F") for a in datasets[name])
    if t.shape[2] < 2:
        t, f = (np.concatenate([a, a[::-1]], axis=2).copy(order="F") for a in (t, f))
    u = oracle.pdps(f, lam, maxiter=2500)
    ctx.set_dataset((t, f))
    g = ctx.gradient(lam, u, regularised=False)
    q = sum(quad.gradient_compliance(lam, u[:, :, i], t[:, :, i]) for i in range(2))
    assert _rel(g, q) <= 1e-10, (g, q)
    g = ctx.gradient(lam, u, regularised=True)
    q = sum(quad.gradient_reg(lam, u[:, :, i], t[:, :, i]) for i in range(2))
    assert _rel(g, q) <= 1e-10, (g, q)
    x = lam * np.array([[0.5, 1.5], [1.0, 0.7]])
    am = oracle.patch_upsample(x, n, n)
    up = oracle.pdps(f, am, maxiter=2500)
    g = ctx.gradient(x, up, regularised=False)
    q = sum(quad.gradient_compliance(am, up[:, :, i], t[:, :, i], grid_shape=x.shape) for i in range(2))
    assert _rel(g, q) <= 1e-10, (g, q)
    g = ctx.gradient(x, up, regularised=True)
    q = sum(quad.gradient_reg(am, up[:, :, i], t[:, :, i], grid_shape=x.shape) for i in range(2))
    assert _rel(g, q) <= 1e-10, (g, q)


def test_fp32_context_gradient(bp, ctx32, oracle, datasets):
    """fp32 mode (north_star: ≤ 1e-5 for the image and the gradient).  An fp32 context solves the lower level in fp32
    (bit-identical to the fp32 oracle: tests/test_gpu_pdps.py) and forms the gradient in fp64 FROM THAT u and the fp32-stored
    truth; the bar is therefore taken against the oracle evaluated on the same fp32 u — the gradient of a different u is a
    different number (δu = 1e-10 already moves it by 1e-4).  Scalar and patch λ, both branches."""
    t, f = (a[:64, :64, :2].copy(order="F") for a in datasets["faces_train_128_10"])
    t32 = t.astype(np.float32).astype(np.float64)
    ctx32.set_dataset((t, f))
    eo = bp.eval_opts(bp.pdps_opts(maxiter=1500))
    x = np.array([[0.03, 0.08], [0.05, 0.06]])
    am = oracle.patch_upsample(x, 64, 64)
    for lam, a, grid in ((0.06, 0.06, None), (x, am, x.shape)):
        for Delta, variant in ((0.1, "nonreg"), (1e-7, "reg")):
            u, cost, g = ctx32.learn_eval(lam, Delta, eo)
            assert np.array_equal(u, u.astype(np.float32).astype(np.float64))          # the fp32 solve's image
            ref = sum(oracle.gradient_dual(variant, a, u[:, :, i], t32[:, :, i], grid_shape=grid) for i in range(2))
            assert _rel(g, ref) <= 1e-5, (variant, g, ref)                             # north_star's fp32 bar
            assert _rel(g, ref) <= 1e-7, (variant, g, ref)                             # what it achieves with room to spare
            assert abs(cost - oracle.cost(u, t32)) <= 1e-6 * cost
