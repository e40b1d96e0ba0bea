"""Pins of the sum-of-regularisers oracle (oracle/sumregs.py) that need no GPU: operator adjointness,
energy decrease, the literal adjoint systems of SumRegsLearningFunction.jl against their compliance-form
restatement, finite differences."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import sumregs as sr


def test_operators_and_adjoints():
    rng = np.random.default_rng(0)
    for kind in sr.KINDS:
        G = sr.op_matrix(kind, 7, 5)
        u = rng.standard_normal((7, 5))
        d1, d2 = sr._grad(kind, u)
        assert np.allclose(np.concatenate([d1.flatten(order="F"), d2.flatten(order="F")]), G @ u.flatten(order="F"))
        q1, q2 = rng.standard_normal((7, 5)), rng.standard_normal((7, 5))
        p1, p2 = sr._grad(kind, np.arange(35.0).reshape(7, 5) ** 1.3)
        q1, q2 = q1 * (p1 != 0), q2 * (p2 != 0)          # duals live where the operator has a row
        assert abs(np.sum(d1 * q1 + d2 * q2) - np.sum(u * sr._grad_T(kind, q1, q2))) < 1e-12
        assert np.allclose(G.T @ np.concatenate([q1.flatten(order="F"), q2.flatten(order="F")]),
                           sr._grad_T(kind, q1, q2).flatten(order="F"))
        assert np.all(np.abs(G @ np.ones(35)) < 1e-15)   # constants are in every kernel
    # operator norm bound used for the step sizes (S12)
    K = np.vstack([sr.op_matrix(k, 9, 9).toarray() for k in sr.KINDS])
    assert np.linalg.norm(K, 2) <= sr.OPNORM3


def test_pdps_decreases_the_energy_and_matches_tv_when_two_weights_vanish(datasets):
    t, f = datasets["cameraman_128_5"]
    f = np.asfortranarray(f[:40, :40, :])
    al = [0.03, 0.02, 0.04]

    def energy(v):
        e = 0.5 * np.sum((v - f[:, :, 0]) ** 2)
        for k, kind in enumerate(sr.KINDS):
            d1, d2 = sr._grad(kind, v)
            e += al[k] * np.sum(np.sqrt(d1 * d1 + d2 * d2))
        return e

    e = [energy(sr.sumregs_pdps(f, al, maxiter=m)[:, :, 0]) for m in (50, 300, 1500)]
    assert e[0] > e[1] > e[2] and e[2] < 0.5 * energy(f[:, :, 0])
    a = sr.sumregs_pdps(f, [0.07, 0.0, 0.0], maxiter=200)
    b = orc.pdps(f, 0.07, maxiter=200, opnorm=sr.OPNORM3)
    assert np.abs(a - b).max() < 1e-13
    a32 = sr.sumregs_pdps(f, al, maxiter=100, dtype=np.float32)
    assert a32.dtype == np.float32 and np.abs(a32 - sr.sumregs_pdps(f, al, maxiter=100)).max() < 1e-4


@pytest.mark.parametrize("variant", ["reg", "nonreg"])
def test_literal_systems_match_the_compliance_form(datasets, variant):
    t, f = datasets["faces_train_128_10"]
    t, f = np.asfortranarray(t[40:64, 40:64, :1]), np.asfortranarray(f[40:64, 40:64, :1])
    x = np.array([0.03, 0.02, 0.04])
    u = sr.sumregs_pdps(f, list(x), maxiter=500)
    if variant == "reg":
        lit = sr.sumregs_gradient_reg(x, u[:, :, 0], t[:, :, 0], refine=3)
    else:
        lit = sr.sumregs_gradient(x, u[:, :, 0], t[:, :, 0], refine=3)
    du = sr.sumregs_gradient_dual(variant, x, u[:, :, 0], t[:, :, 0])
    assert np.all(np.abs(lit - du) <= 1e-9 * np.abs(lit)), (lit, du)
    # patch non-regularised variant (:330-407) has a compliance form too; the patch regularised one
    # (:195-262) is row-scaled by a different map per operator and has none
    if variant == "nonreg":
        xp = np.stack([np.array([[0.03, 0.05], [0.02, 0.04]]) * s for s in (1.0, 0.7, 1.3)], axis=2)
        maps = [orc.patch_upsample(xp[:, :, k], 24, 24) for k in range(3)]
        lit = sr.sumregs_gradient(maps, u[:, :, 0], t[:, :, 0], grid_shape=(2, 2), refine=3, eps_act=sr.EPS)
        du = sr.sumregs_gradient_dual("nonreg", maps, u[:, :, 0], t[:, :, 0], grid_shape=(2, 2))
        assert np.all(np.abs(lit - du) <= 1e-8 * np.abs(lit).max())
    else:
        with pytest.raises(NotImplementedError):
            sr.sumregs_gradient_dual("reg", [np.ones((24, 24))] * 3, u[:, :, 0], t[:, :, 0], grid_shape=(2, 2))


def test_learning_function_shapes_and_branches(datasets):
    t, f = datasets["cameraman_128_5"]
    t, f = np.asfortranarray(t[:20, :20, :]), np.asfortranarray(f[:20, :20, :])
    x = np.array([0.02, 0.02, 0.02])
    u, c, g = sr.sumregs_learning_function(x, (t, f), 0.01, maxiter=200)       # Δ > Δt = 1e-3: sumregs_gradient
    u2, c2, g2 = sr.sumregs_learning_function(x, (t, f), 1e-4, maxiter=200)    # Δ ≤ Δt: sumregs_gradient_reg
    assert g.shape == (3,) and g2.shape == (3,) and c == c2 == orc.cost(u, t)
    assert np.all(np.sign(g) == np.sign(g2)) and not np.allclose(g, g2)
    xp = 0.02 * np.ones((2, 2, 3))
    up, cp, gp = sr.sumregs_learning_function(xp, (t, f), 0.1, maxiter=200)
    assert gp.shape == (2, 2, 3) and np.allclose(up, u) and np.allclose(gp.sum(axis=(0, 1)), g, rtol=1e-5)


def test_patch_regularised_system_reduces_to_the_scalar_one_for_uniform_maps(datasets):
    """Pin for the row-scaled patch system (:195-262), which has no compliance form to be checked against: with three
    UNIFORM maps its rows are scaled by constants, it is the scalar system (:112-167) at the same γ, and the patch
    gradients must add up to the scalar gradient (PatchOp adjoint = block sums, S7)."""
    t, f = datasets["faces_train_128_10"]
    t, f = np.asfortranarray(t[40:60, 40:60, :1]), np.asfortranarray(f[40:60, 40:60, :1])
    x = np.array([0.03, 0.02, 0.04])
    u = sr.sumregs_pdps(f, list(x), maxiter=400)
    maps = [np.full((20, 20), v) for v in x]
    for gamma in (1e3, 1e8):
        gp = sr.sumregs_gradient_reg(maps, u[:, :, 0], t[:, :, 0], grid_shape=(2, 2), gamma=gamma, refine=3)
        gs = sr.sumregs_gradient_reg(x, u[:, :, 0], t[:, :, 0], gamma=gamma, refine=3)
        assert gp.shape == (2, 2, 3)
        assert np.all(np.abs(gp.sum(axis=(0, 1)) - gs) <= 1e-9 * np.abs(gs).max()), (gamma, gp.sum(axis=(0, 1)), gs)
