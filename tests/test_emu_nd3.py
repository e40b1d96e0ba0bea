"""CPU check of the scalar `sumregs_gradient_reg` on the nested-dissection solver (bpldenoising_b200/csrc/
nd_sumregs.cuh + nd_solver.cuh at coupling radius 2) where no GPU exists: the device code is compiled with g++
against tests/emu/emu_cuda.h and run for one image in the launch order of gradient_nd.cuh's run_gradient3_nd_reg,
then compared with the oracle's literal system (/root/reference/src/SumRegsLearningFunction.jl:112-167, γ = 1e3).
The GPU parity test proper is tests/test_gpu_sumregs.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import sumregs as sr

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
CSRC = os.path.join(HERE, "..", "bpldenoising_b200", "csrc")


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(EMU, "_build", "libemu_nd3.so")
    srcs = [os.path.join(EMU, "emu_nd3.cpp"), os.path.join(EMU, "emu_cuda.h")] + \
           [os.path.join(CSRC, f) for f in ("nd_symbolic.h", "nd_solver.cuh", "nd_sumregs.cuh", "lu_band.cuh",
                                            "sumregs_stencils.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", "-DBPLTV_EMU", "-ffp-contract=off",
                        "-o", out, srcs[0]], check=True)
    lib = C.CDLL(out)
    lib.emu_nd3_gradient_reg.restype = C.c_int
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _case(n, seed):
    rng = np.random.default_rng(seed)
    t = np.round(rng.random((n, n)) * 255) / 255
    u = np.asfortranarray(t + 0.05 * rng.standard_normal((n, n)))
    u[:3, :3] = u[0, 0]                      # an exactly flat block: |∇_k u| = 0, the γ·I tensors
    u[n - 4:, n - 5:] = u[n - 1, n - 1]
    return np.asfortranarray(t), u


def _run(lib, u, t, x, gamma=1e3, refine=1, leaf=4, small=1, want_ast=False):
    n = u.shape[0]
    out, stats, p = np.zeros(3), np.zeros(6), np.zeros(n * n)
    ast = np.zeros(13 * n * n) if want_ast else None
    rc = lib.emu_nd3_gradient_reg(n, _ptr(u.flatten(order="F")), _ptr(t.flatten(order="F")),
                                  _ptr(np.asarray(x, dtype=np.float64)), C.c_double(gamma), refine, leaf, int(small),
                                  _ptr(out), _ptr(stats), _ptr(p), _ptr(ast))
    assert rc == 0
    return out, stats, p, (None if ast is None else ast.reshape(n * n, 13))


def _literal_matrix(x, u, gamma):
    n = u.shape[0]
    uf = u.flatten(order="F")
    A = sp.identity(n * n, format="csr")
    for k, kind in enumerate(sr.KINDS):
        G = sr.op_matrix(kind, n)
        BmC, _ = sr._sets_reg(G, uf, gamma)
        A = A + x[k] * (G.T @ BmC @ G)
    return A.toarray()


def test_stencil_form_is_the_literal_matrix(lib):
    """nd3_stencil_kernel's 13 forward offsets per node = the lower triangle of I + Σ α_k G_kᵀ(B_k − C_k)G_k (:160)"""
    n = 11
    t, u = _case(n, 3)
    x = np.array([0.05, 0.04, 0.06])
    _, _, _, ast = _run(lib, u, t, x, want_ast=True)
    A = _literal_matrix(x, u, 1e3)
    S = np.zeros_like(A)
    h = 0
    offs = [(0, 0), (1, 0), (2, 0)] + [(di, dj) for dj in (1, 2) for di in (-2, -1, 0, 1, 2)]
    for h, (di, dj) in enumerate(offs):
        for v in range(n * n):
            i, j = v % n, v // n
            if 0 <= i + di < n and 0 <= j + dj < n:
                w = (i + di) + n * (j + dj)
                S[w, v] = ast[v, h]
                S[v, w] = ast[v, h]
            else:
                assert ast[v, h] == 0.0
    assert np.abs(S - A).max() <= 1e-15 * np.abs(A).max()
    assert np.abs(A - A.T).max() <= 1e-13 * np.abs(A).max()         # symmetric: why the Cholesky applies


@pytest.mark.parametrize("n,leaf", [(8, 4), (13, 4), (24, 6)])
def test_scalar_reg_gradient(lib, n, leaf):
    t, u = _case(n, 40 + n)
    x = np.array([0.05, 0.04, 0.06])
    got, stats, p, _ = _run(lib, u, t, x, leaf=leaf)
    lit = sr.sumregs_gradient_reg(x, u, t, refine=3)
    assert stats[1] == 0 and stats[0] <= 1e-15, stats
    assert np.all(np.abs(got - lit) <= 1e-11 * np.abs(lit).max()), (got, lit)


def test_small_front_kernels_equal_the_generic_ones(lib):
    t, u = _case(20, 9)
    x = np.array([0.1, 0.02, 0.07])
    g0, s0, p0, _ = _run(lib, u, t, x, small=0)
    g1, s1, p1, _ = _run(lib, u, t, x, small=1)
    assert np.all(np.abs(g1 - g0) <= 1e-12 * np.abs(g0).max())
    assert np.linalg.norm(p1 - p0) <= 1e-12 * np.linalg.norm(p0)


def test_one_operator_switched_off(lib):
    """α_k = 0 for two operators: the TV gradient_reg system with γ = 1e3 (forward differences only)"""
    t, u = _case(16, 5)
    x = np.array([0.08, 0.0, 0.0])
    got, stats, _, _ = _run(lib, u, t, x)
    lit = sr.sumregs_gradient_reg(x, u, t, refine=3)
    assert np.all(np.abs(got - lit) <= 1e-11 * np.abs(lit).max()), (got, lit)


# ---------------------------------------------------------------------------
# sumregs_gradient (non-regularised) in multiplier space: 3-6 unknowns per pixel on the same tree
# ---------------------------------------------------------------------------
def _run_mult(lib, u, t, x=None, maps=None, grid=(1, 1), refine=1, leaf=4, csize=1, smem_limit=1 << 30):
    n = u.shape[0]
    ng = grid[0] * grid[1]
    out, stats, p = np.zeros(3 * ng), np.zeros(6), np.zeros(n * n)
    am = None if maps is None else np.concatenate([np.asarray(m).flatten(order="F") for m in maps])
    a3 = None if x is None else np.asarray(x, dtype=np.float64)
    lib.emu_nd3_gradient_mult.restype = C.c_int
    rc = lib.emu_nd3_gradient_mult(n, _ptr(u.flatten(order="F")), _ptr(t.flatten(order="F")), _ptr(a3), _ptr(am),
                                   grid[0], grid[1], C.c_double(1e-12), C.c_double(sr.EPS), refine, leaf, csize, C.c_longlong(smem_limit), _ptr(out),
                                   _ptr(stats), _ptr(p))
    assert rc == 0
    return out.reshape(3, grid[1], grid[0]).transpose(2, 1, 0), stats, p      # [operator][patch] → (pi, pj, operator)


@pytest.mark.parametrize("n,leaf", [(8, 4), (13, 4), (20, 5)])
def test_scalar_nonreg_gradient_multiplier_form(lib, n, leaf):
    """≤ 1e-10 against the compliance-form CPU checker (the same formulation), ≤ 1e-6 against the refined literal
    saddle-point system (:264-327) — the bars of tests/test_gpu_sumregs.py"""
    t, u = _case(n, 70 + n)
    x = np.array([0.05, 0.04, 0.06])
    got, stats, _ = _run_mult(lib, u, t, x=x, leaf=leaf)
    assert stats[2] == 0 and stats[3] > 3 * n * n, stats          # flat pixels carry two modes per operator
    du = sr.sumregs_gradient_dual("nonreg", x, u, t)
    assert np.all(np.abs(got[0, 0] - du) <= 1e-10 * np.abs(du).max()), (got, du, stats)
    lit = sr.sumregs_gradient(x, u, t, refine=3)
    assert np.all(np.abs(got[0, 0] - lit) <= 1e-6 * np.abs(lit).max()), (got, lit)


def test_patch_nonreg_gradient_multiplier_form(lib):
    from oracle import oracle as orc
    n = 16
    t, u = _case(n, 12)
    xp = np.stack([np.array([[0.03, 0.05], [0.02, 0.04]]) * s for s in (1.0, 0.7, 1.3)], axis=2)
    maps = [np.asfortranarray(orc.patch_upsample(xp[:, :, k], n, n)) for k in range(3)]
    got, stats, _ = _run_mult(lib, u, t, maps=maps, grid=(2, 2))
    dp = sr.sumregs_gradient_dual("nonreg", maps, u, t, grid_shape=(2, 2))
    assert stats[2] == 0
    assert got.shape == dp.shape and np.all(np.abs(got - dp) <= 1e-10 * np.abs(dp).max()), (got, dp)


def test_cluster_shared_front_factorisation_is_invisible(lib):
    """nd_factor_cluster_kernel: the front dealt over the CTAs of a cluster (assembly, write-back and trailing tiles shared,
    diagonal block and panel redundant) gives the bits of the single-CTA kernel, for 2 and 3 CTAs per front"""
    t, u = _case(10, 31)
    x = np.array([0.05, 0.04, 0.06])
    g1, s1, p1 = _run_mult(lib, u, t, x=x)
    for cs in (2, 3):
        g, s, p = _run_mult(lib, u, t, x=x, csize=cs)
        assert np.array_equal(g, g1) and np.array_equal(p, p1) and s[0] == s1[0] and s[1] == s1[1], cs


def test_eight_column_block_steps_for_fronts_beyond_the_panel(lib):
    """fronts whose 16-column panel would not fit in shared memory take the 8-column kernels (nd_factor8_*): here the limit
    is set so low that the upper levels do — the gradient agrees with the 16-column factorisation to rounding, alone and
    shared by a cluster"""
    t, u = _case(13, 17)
    x = np.array([0.05, 0.04, 0.06])
    g16, s16, p16 = _run_mult(lib, u, t, x=x)
    assert s16[5] == 0
    g8, s8, p8 = _run_mult(lib, u, t, x=x, smem_limit=20000)
    assert s8[5] > 0 and s8[2] == 0
    assert np.all(np.abs(g8 - g16) <= 1e-11 * np.abs(g16).max()) and np.linalg.norm(p8 - p16) <= 1e-11 * np.linalg.norm(p16)
    g8c, s8c, p8c = _run_mult(lib, u, t, x=x, smem_limit=20000, csize=2)
    assert np.array_equal(g8c, g8) and np.array_equal(p8c, p8)
