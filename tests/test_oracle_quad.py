"""What IS the reference's gradient at the 1e-10 level?  Pins of the binary128 arbiter (oracle/quad_adjoint.c, oracle/quad.py).

The reference assembles the adjoint system of `gradient` (/root/reference/src/TVLearningFunctionVec.jl:98-135) with entries from
1 to 1/eps() and solves it with a sparse LU in double.  Solved in binary128 instead:
  * the literal system with exactly formed entries and the multiplier-space (compliance) form the CUDA path factorises give the
    SAME functional (to the last bit of the double they are rounded to) — the reformulation is exact;
  * the literal system with entries rounded to double in the reference's operation order differs from that by 1e-10 … 1e-7:
    the noise of forming the projector `Den − prodKuKu` by cancellation and scaling it by α/|∇u|.  This, not the solver, is
    the gap that round 1 could only bound by 1e-6; a double-precision sparse LU with extended-precision refinement
    (oracle.gradient_scalar) does not even reach the exact solution of its own matrix (another 1e-8 … 1e-7);
  * the double-precision compliance solve (oracle.gradient_dual, the CPU twin of the CUDA path) reproduces the binary128 value
    to 1e-14.
gradient_reg (:137-161) is well posed: double assembly costs ≤ 1e-13, SuperLU + refinement is within 1e-9 of binary128."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import quad


def _rel(a, b):
    a, b = np.atleast_1d(a).astype(float), np.atleast_1d(b).astype(float)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _crop(datasets, name, n, lam, its=2500):
    t, f = datasets[name]
    t = np.asfortranarray(t[40:40 + n, 50:50 + n, 0])
    f = np.asfortranarray(f[40:40 + n, 50:50 + n, :1])
    u = orc.pdps(f, lam, maxiter=its)[:, :, 0]
    return t, u


@pytest.mark.parametrize("name,lam", [("cameraman_128_5", 0.1), ("faces_train_128_10", 0.05), ("circle_128_10", 0.02)])
def test_scalar_gradient_formulations_in_binary128(oracle, datasets, name, lam):
    t, u = _crop(datasets, name, 28, lam)
    lit_exact = quad.gradient_literal(lam, u, t, assemble_quad=True)
    lit_double = quad.gradient_literal(lam, u, t, assemble_quad=False)
    comp = quad.gradient_compliance(lam, u, t)
    assert _rel(comp, lit_exact) <= 1e-14                      # the compliance form IS the reference's system
    noise = _rel(lit_double, lit_exact)
    assert noise <= 1e-5                                       # … whose double-rounded entries move it by this much (1e-10 … 1e-7)
    assert _rel(orc.gradient_dual("nonreg", lam, u, t), comp) <= 1e-12
    # the double-precision literal solve with refinement sits within solver noise + assembly noise of both
    assert _rel(orc.gradient_scalar(lam, u, t, refine=4), lit_double) <= 1e-6
    # regularised branch: well posed
    reg_exact = quad.gradient_reg(lam, u, t, assemble_quad=True)
    reg_double = quad.gradient_reg(lam, u, t, assemble_quad=False)
    assert _rel(reg_double, reg_exact) <= 1e-12
    assert _rel(orc.gradient_reg_scalar(lam, u, t, refine=4), reg_exact) <= 1e-9
    assert _rel(orc.gradient_dual("reg", lam, u, t), reg_exact) <= 1e-12


def test_patch_gradient_formulations_in_binary128(oracle, datasets):
    n = 24
    x = np.array([[0.02, 0.05, 0.03], [0.04, 0.01, 0.06]])
    am = orc.patch_upsample(x, n, n)
    t, f = datasets["circle_128_10"]
    t = np.asfortranarray(t[30:30 + n, 40:40 + n, 0])
    u = orc.pdps(np.asfortranarray(f[30:30 + n, 40:40 + n, :1]), am, maxiter=2500)[:, :, 0]
    lit_exact = quad.gradient_literal(am, u, t, grid_shape=x.shape, assemble_quad=True)
    comp = quad.gradient_compliance(am, u, t, grid_shape=x.shape)
    assert _rel(comp, lit_exact) <= 1e-14
    assert _rel(orc.gradient_dual("nonreg", am, u, t, grid_shape=x.shape), comp) <= 1e-12
    assert _rel(orc.gradient_patch(am, x.shape, u, t, refine=4), quad.gradient_literal(am, u, t, grid_shape=x.shape)) <= 1e-6
    reg_exact = quad.gradient_reg(am, u, t, grid_shape=x.shape, assemble_quad=True)
    assert _rel(orc.gradient_reg_patch(am, x.shape, u, t, refine=4), reg_exact) <= 1e-9
    assert _rel(orc.gradient_dual("reg", am, u, t, grid_shape=x.shape), reg_exact) <= 1e-12


def test_p_itself_matches(oracle, datasets):
    """not only the functional: the adjoint state p of both forms agrees in binary128"""
    t, u = _crop(datasets, "cameraman_128_5", 20, 0.1, its=1500)
    _, p_lit = quad.gradient_literal(0.1, u, t, assemble_quad=True, return_p=True)
    _, p_comp = quad.gradient_compliance(0.1, u, t, return_p=True)
    assert _rel(p_comp, p_lit) <= 1e-14
