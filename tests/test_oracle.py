"""Pins for the CPU oracle (the reference ships no golden vectors, so the
build creates its own pins).  None of these tests needs a GPU."""
import os

import numpy as np
import pytest

from conftest import ROOT, rel_l2


def test_operator_adjointness(oracle):
    # (1) <∇u, q> = <u, ∇ᵀq> to rounding, C loops and numpy slices, ragged shapes
    rng = np.random.default_rng(1)
    import ctypes
    for (M, N) in [(7, 5), (1, 9), (9, 1), (16, 16), (33, 20)]:
        u = np.asfortranarray(rng.standard_normal((M, N)))
        q1 = np.asfortranarray(rng.standard_normal((M, N)))
        q2 = np.asfortranarray(rng.standard_normal((M, N)))
        g1 = np.zeros_like(u); g2 = np.zeros_like(u); v = np.zeros_like(u)
        P = ctypes.POINTER(ctypes.c_double)
        L = oracle.lib()
        L.oracle_fwd_grad_f64(u.ctypes.data_as(P), M, N, g1.ctypes.data_as(P), g2.ctypes.data_as(P))
        L.oracle_fwd_grad_T_f64(q1.ctypes.data_as(P), q2.ctypes.data_as(P), M, N, v.ctypes.data_as(P))
        lhs = np.sum(g1 * q1) + np.sum(g2 * q2)
        rhs = np.sum(u * v)
        assert abs(lhs - rhs) <= 1e-13 * (abs(lhs) + 1)
        n1, n2 = oracle.grad_np(u)
        assert np.array_equal(n1, g1) and np.array_equal(n2, g2)
        assert np.allclose(oracle.grad_T_np(q1, q2), v, rtol=0, atol=1e-15)
        # matrix(op,n) (S5) acts like the stencil on the column-major vec
        if M == N:
            G = oracle.grad_matrix(M)
            Gu = G @ u.flatten(order="F")
            assert np.array_equal(Gu[:M * M].reshape((M, M), order="F"), g1)
            assert np.array_equal(Gu[M * M:].reshape((M, M), order="F"), g2)
    # last row / column of ∇u are zero (Neumann, S4)
    assert np.all(g1[-1, :] == 0) and np.all(g2[:, -1] == 0)


def test_step_size_recursion(oracle):
    s = oracle.step_sizes(50)
    tau, sigma, omega = s[:, 0], s[:, 1], s[:, 2]
    # /root/reference/src/TVLearningFunctionVec.jl:36-37 with R_K = √8
    assert tau[0] == 5.0 / np.sqrt(8.0) and sigma[0] == (0.99 / 5) / np.sqrt(8.0)
    assert np.allclose(omega, 1.0 / np.sqrt(1.0 + 2.0 * tau), rtol=1e-15)
    # acceleration keeps τσ constant and < 1/R_K²
    assert np.allclose(tau * sigma, tau[0] * sigma[0], rtol=1e-13)
    assert np.all(np.diff(tau) < 0) and np.all(np.diff(sigma) > 0)
    s2 = oracle.step_sizes(5, accel=False)
    assert np.all(s2[:, 2] == 1.0) and np.all(s2[:, 0] == s2[0, 0])


@pytest.mark.parametrize("case", ["scalar", "map", "rho", "init_f", "noaccel"])
def test_c_port_matches_numpy_restatement(oracle, datasets, case):
    # the C loops and the independent slice-based numpy restatement are bit-identical
    f = datasets["cameraman_128_5"][1][:48, :40, 0].copy(order="F")
    kw = dict(maxiter=120)
    alpha = 0.07
    if case == "map":
        alpha = oracle.patch_upsample(np.array([[0.02, 0.1], [0.2, 0.05], [0.01, 0.3]]), 48, 40)
    if case == "rho":
        kw["rho"] = 0.3
    if case == "init_f":
        kw["init_mode"] = 1
    if case == "noaccel":
        kw["accel"] = False
    uc = oracle.pdps(f, alpha, **kw)[:, :, 0]
    un = oracle.pdps_numpy(f, alpha, **kw)
    assert np.array_equal(uc, un)


def test_pdps_limits_and_energy(oracle, datasets):
    f = datasets["cameraman_128_5"][1][:64, :64, 0].copy(order="F")
    # λ → 0: u → f (2)
    # (x⁰ = 0 and the accelerated τ_k ~ 1/k make this an O(1/k) approach)
    u0 = oracle.pdps(f, 1e-9, maxiter=5000)[:, :, 0]
    assert np.abs(u0 - f).max() < 1e-4
    u1 = oracle.pdps(f, 1e-9, maxiter=50, init_mode=1)[:, :, 0]
    assert np.abs(u1 - f).max() < 1e-8
    # λ huge: u → constant = mean(f) (the TV term dominates)
    ub = oracle.pdps(f, 1e3, maxiter=20000)[:, :, 0]
    assert np.abs(ub - f.mean()).max() < 1e-3
    # primal energy at iteration k approaches the minimum monotonically on a coarse grid
    def energy(u, lam):
        g1, g2 = oracle.grad_np(u)
        return 0.5 * np.sum((u - f) ** 2) + lam * np.sum(np.sqrt(g1 * g1 + g2 * g2))
    e = [energy(oracle.pdps(f, 0.1, maxiter=k)[:, :, 0], 0.1) for k in (200, 800, 3200, 12800)]
    assert e[0] > e[1] > e[2] >= e[3] - 1e-9
    # primal–dual gap → 0: dual energy from the numpy restatement's y
    x, y1, y2 = oracle.pdps_numpy(f, 0.1, maxiter=3000, return_dual=True)
    div = oracle.grad_T_np(y1, y2)
    dual = -0.5 * np.sum(div * div) + np.sum(div * f)  # -(½‖∇ᵀy‖² - <∇ᵀy,f>), |y|≤λ
    assert np.all(y1 * y1 + y2 * y2 <= 0.1 ** 2 * (1 + 1e-12))
    gap = energy(x, 0.1) - dual
    assert 0 <= gap < 1e-3 * energy(x, 0.1)


def test_batched_images_are_independent(oracle, datasets):
    # S9: a stack is O independent 2-D problems sharing the step sizes
    f = datasets["faces_train_128_10"][1][:32, :32, :3].copy(order="F")
    u = oracle.pdps(f, 0.05, maxiter=100)
    for o in range(3):
        assert np.array_equal(u[:, :, o], oracle.pdps(f[:, :, o], 0.05, maxiter=100)[:, :, 0])
    # fp32 oracle stays within 1e-5 of fp64
    u32 = oracle.pdps(f, 0.05, maxiter=100, dtype=np.float32)
    assert rel_l2(u32, u) < 1e-5


def test_patchop_adjoint(oracle):
    rng = np.random.default_rng(3)
    for (M, N, m, n) in [(128, 128, 2, 2), (10, 7, 3, 2), (5, 5, 5, 5), (9, 4, 1, 1)]:
        x = rng.standard_normal((m, n)); g = rng.standard_normal((M, N))
        up = oracle.patch_upsample(x, M, N)
        assert up.shape == (M, N)
        assert abs(np.sum(up * g) - np.sum(x * oracle.patch_adjoint(g, m, n))) < 1e-10
    # 2×2 on 128×128: four 64×64 blocks (/root/reference/src/BPLDenoising.jl:356)
    up = oracle.patch_upsample(np.array([[1., 2.], [3., 4.]]), 128, 128)
    assert up[0, 0] == 1 and up[127, 0] == 3 and up[0, 127] == 2 and up[64, 64] == 4 and up[63, 63] == 1


def test_gradient_reg_matches_finite_differences(oracle, datasets):
    # (3): the regularised adjoint gradient is the derivative of the cost
    t, f = (a[:48, :48, 0].copy(order="F") for a in datasets["cameraman_128_5"])
    lam, h = 0.08, 1e-4
    its = 20000  # converged solve, so the fixed-iteration map is differentiable in λ
    u = oracle.pdps(f, lam, maxiter=its)[:, :, 0]
    g = oracle.gradient_reg_scalar(lam, u, t)
    cp = oracle.cost(oracle.pdps(f, lam + h, maxiter=its)[:, :, 0], t)
    cm = oracle.cost(oracle.pdps(f, lam - h, maxiter=its)[:, :, 0], t)
    fd = (cp - cm) / (2 * h)
    assert abs(g - fd) < 2e-3 * abs(fd)
    # and the non-regularised variant agrees with it away from the kink
    g2 = oracle.gradient_scalar(lam, u, t)
    assert abs(g2 - fd) < 0.1 * abs(fd)


@pytest.mark.parametrize("variant", ["reg", "nonreg"])
def test_dual_formulation_matches_literal_system(oracle, datasets, variant):
    # (4): the compliance-form banded Cholesky equals the literal sparse solve
    t, f = (a[:40, :40, 0].copy(order="F") for a in datasets["cameraman_128_5"])
    u = oracle.pdps(f, 0.1, maxiter=3000)[:, :, 0]
    if variant == "reg":
        lit = oracle.gradient_reg_scalar(0.1, u, t, refine=3)
    else:
        lit = oracle.gradient_scalar(0.1, u, t, refine=4)
    dual = oracle.gradient_dual(variant, 0.1, u, t)
    assert abs(dual - lit) <= (1e-9 if variant == "reg" else 1e-6) * abs(lit)
    # patch variant
    x = np.array([[0.05, 0.1], [0.2, 0.08]])
    am = oracle.patch_upsample(x, 40, 40)
    up = oracle.pdps(f, am, maxiter=1500)[:, :, 0]
    if variant == "reg":
        litp = oracle.gradient_reg_patch(am, (2, 2), up, t, refine=3)
    else:
        litp = oracle.gradient_patch(am, (2, 2), up, t, refine=4)
    dualp = oracle.gradient_dual(variant, am, up, t, grid_shape=(2, 2))
    assert rel_l2(dualp, litp) <= (1e-9 if variant == "reg" else 1e-6)


def test_learning_function_protocol(oracle, datasets):
    # tv_op_learning_function: Δ > Δt → gradient, else gradient_reg; sums over images
    t, f = (a[:24, :24, :2].copy(order="F") for a in datasets["faces_train_128_10"])
    u, c, g = oracle.tv_op_learning_function(0.05, (t, f), 0.1, maxiter=300)
    assert u.shape == (24, 24, 2) and c == oracle.cost(u, t)
    g_sum = sum(oracle.gradient_scalar(0.05, u[:, :, i], t[:, :, i]) for i in range(2))
    assert g == g_sum
    _, _, gr = oracle.tv_op_learning_function(0.05, (t, f), 1e-7, u=u)
    gr_sum = sum(oracle.gradient_reg_scalar(0.05, u[:, :, i], t[:, :, i]) for i in range(2))
    assert gr == gr_sum
    x = np.full((2, 2), 0.05)
    _, _, gp = oracle.tv_op_learning_function(x, (t, f), 0.1, maxiter=300)
    assert gp.shape == (2, 2)
    # constant patch grid ≡ scalar λ: same u, and the patch gradient sums to the scalar one
    assert abs(gp.sum() - g) < 1e-4 * abs(g)


def test_oracle_pins(oracle, datasets):
    """Oracle-derived known answers (tools/make_oracle_pins.py): guard against the
    oracle drifting.  They are NOT reference outputs (parity unpinned)."""
    pins = np.load(os.path.join(ROOT, "tests", "golden", "oracle_pins.npz"))
    t, f = datasets["cameraman_128_5"]
    u = oracle.pdps(f, 0.1, maxiter=int(pins["maxiter"]))
    assert np.array_equal(u[::8, ::8, 0], pins["u_sub"])
    assert abs(oracle.cost(u, t) - float(pins["cost"])) <= 1e-12 * float(pins["cost"])
    g = oracle.gradient_reg_scalar(0.1, u[:, :, 0], t[:, :, 0])
    assert abs(g - float(pins["grad_reg"])) <= 1e-9 * abs(float(pins["grad_reg"]))
    g = oracle.gradient_scalar(0.1, u[:, :, 0], t[:, :, 0])
    assert abs(g - float(pins["grad"])) <= 1e-4 * abs(float(pins["grad"]))  # LU-order noise, SURVEY §7.3-2


def test_fused_cpu_variant_is_bit_identical(oracle):
    """bench.py's stronger CPU baseline (one sweep per iteration instead of the reference's separate passes) performs the
    same IEEE operations per pixel: identical bits, fp64 and fp32, scalar λ and λ-map, ragged shapes."""
    rng = np.random.default_rng(5)
    for shape in ((40, 33, 3), (1, 17, 1), (19, 1, 2), (64, 64, 2)):
        f = np.asfortranarray(np.round(rng.random(shape) * 255) / 255)
        am = 0.03 + 0.1 * rng.random(shape[:2])
        for dt in (np.float64, np.float32):
            assert np.array_equal(oracle.pdps(f, 0.08, maxiter=70, dtype=dt), oracle.pdps(f, 0.08, maxiter=70, dtype=dt, fused=True))
            assert np.array_equal(oracle.pdps(f, am, maxiter=50, dtype=dt), oracle.pdps(f, am, maxiter=50, dtype=dt, fused=True))


def test_band_cholesky_c_loop_equals_the_numpy_loop(oracle):
    """oracle.gradient_dual factorises with liboracle's C loop (oracle_chol_band_guard); the numpy loop it replaced is
    kept as its cross-check: identical bits, pivot floor included."""
    rng = np.random.default_rng(3)
    Nd, bw = 1500, 70
    B = np.zeros((Nd, Nd))
    for r in range(1, bw + 1):
        if 5 <= r < 20:
            continue                                     # structural zeros inside the band
        v = rng.uniform(-1, 1, Nd - r)
        B += np.diag(v, -r)
    B[100, :] = 0.0
    D = rng.uniform(0.5, 1.5, Nd)
    D[100] = 1e-20                                       # a vanished pivot (an uncoupled unknown): the floor has to act
    A = B @ B.T + np.diag(D)                             # SPD, half-bandwidth 2·bw
    w = 2 * bw
    ab = np.zeros((w + 1, Nd))
    for r in range(w + 1):
        ab[r, :Nd - r] = np.diag(A, -r)
    a1, a2 = ab.copy(), ab.copy()
    g1 = oracle._chol_band_guard(a1, 1e-3)
    g2 = oracle._chol_band_guard_py(a2, 1e-3)
    assert g1 == g2 >= 1
    assert np.all(np.isfinite(a1)) and np.array_equal(a1, a2)
