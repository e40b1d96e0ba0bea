"""CPU oracle for the bpltv hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this module.  The product package
``bpldenoising_b200`` never does.

PARITY UNPINNED: the reference's solver arithmetic lives in
un-vendored, un-pinned Julia packages and the reference ships no golden
vectors; Julia is not installed.  This module restates

* the lower-level solve (``pdps``: C restatement in ``bpltv_oracle.c``;
  ``pdps_numpy``: an independent slice-based numpy restatement used to
  cross-check the C port),
* the learning function and both gradient variants *literally*, by assembling
  the same sparse matrices the reference assembles and calling a sparse direct
  solver (scipy SuperLU instead of Julia's UMFPACK/CHOLMOD), optionally with
  extended-precision iterative refinement,

each function citing the reference file:line it follows.  Assumptions S1–S9
are listed in docs/SEMANTICS.md.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# /root/reference/src/TVLearningFunctionVec.jl:33-43
DEFAULT_PARAMS = dict(rho=0.0, tau0=5.0, sigma0=0.99 / 5, accel=True, maxiter=5000)
OPNORM = float(np.sqrt(8.0))  # S2: R_K = opnorm_estimate(FwdGradientOp) = √8


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so (gcc).  Building the checker is not using it."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, n) for n in ("bpltv_oracle.c", "pdps_body.inc", "Makefile")]
    if force or not os.path.exists(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        for suf, ct in (("f64", ctypes.c_double), ("f32", ctypes.c_float)):
            P = ctypes.POINTER(ct)
            f = getattr(L, f"oracle_pdps_{suf}")
            f.restype = ctypes.c_int
            f.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, ctypes.c_int,
                          ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, P]
            ff = getattr(L, f"oracle_pdps_fused_{suf}")
            ff.restype = ctypes.c_int
            ff.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, ctypes.c_int,
                           ctypes.c_double, ctypes.c_double, ctypes.c_double,
                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, P]
            g = getattr(L, f"oracle_fwd_grad_{suf}")
            g.restype = None
            g.argtypes = [P, ctypes.c_int, ctypes.c_int, P, P]
            gt = getattr(L, f"oracle_fwd_grad_T_{suf}")
            gt.restype = None
            gt.argtypes = [P, P, ctypes.c_int, ctypes.c_int, P]
            c = getattr(L, f"oracle_cost_{suf}")
            c.restype = ctypes.c_double
            c.argtypes = [P, P, ctypes.c_size_t]
        L.oracle_step_sizes.restype = None
        L.oracle_step_sizes.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_double)]
        L.oracle_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _fortran3(a, dtype):
    a = np.asarray(a, dtype=dtype)
    if a.ndim == 2:
        a = a[:, :, None]
    return np.asfortranarray(a)


def _ptr(a):
    ct = ctypes.c_double if a.dtype == np.float64 else ctypes.c_float
    return a.ctypes.data_as(ctypes.POINTER(ct))


def step_sizes(maxiter, tau0=5.0, sigma0=0.99 / 5, opnorm=OPNORM, accel=True):
    out = np.zeros((max(maxiter, 1), 3))
    lib().oracle_step_sizes(tau0, sigma0, opnorm, int(accel), maxiter,
                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out[:maxiter]


# --------------------------------------------------------------------------
# PatchOp (S7): block-constant up-sampling of an m×n grid to M×N, adjoint =
# block sums.  Call sites /root/reference/src/TVLearningFunctionVec.jl:58-60,
# 166,171,214,253.  Pixel i (0-based) belongs to patch floor(i*m/M).
# --------------------------------------------------------------------------
def patch_index(M, m):
    return (np.arange(M) * m) // M


def patch_upsample(x, M, N):
    x = np.asarray(x, dtype=np.float64)
    m, n = x.shape
    return np.asfortranarray(x[np.ix_(patch_index(M, m), patch_index(N, n))])


def patch_adjoint(g, m, n):
    M, N = g.shape
    out = np.zeros((m, n))
    np.add.at(out, (patch_index(M, m)[:, None], patch_index(N, n)[None, :]), g)
    return out


# --------------------------------------------------------------------------
# Lower-level solve
# --------------------------------------------------------------------------
def pdps(f, alpha, *, maxiter=5000, tau0=5.0, sigma0=0.99 / 5, rho=0.0, accel=True,
         opnorm=OPNORM, init_mode=0, dtype=np.float64, nthreads=0, fused=False):
    """denoise(data, x, op): /root/reference/src/TVLearningFunctionVec.jl:45-70.

    ``alpha``: scalar, or M×N map (already up-sampled).  Returns M×N×O (Fortran).
    ``fused``: the single-sweep variant of the same recursion (bit-identical; the stronger CPU baseline of bench.py).
    """
    f3 = _fortran3(f, dtype)
    M, N, O = f3.shape
    al = np.asarray(alpha, dtype=dtype)
    is_map = al.ndim == 2
    if is_map:
        assert al.shape == (M, N)
        al = np.asfortranarray(al)
    else:
        al = al.reshape(1)
    u = np.zeros_like(f3, order="F")
    suf = "f64" if dtype == np.float64 else "f32"
    if fused:
        assert rho == 0.0
        rc = getattr(lib(), "oracle_pdps_fused_" + suf)(_ptr(f3), M, N, O, _ptr(al), int(is_map), tau0, sigma0, opnorm,
                                                        int(accel), maxiter, init_mode, nthreads, _ptr(u))
    else:
        rc = getattr(lib(), "oracle_pdps_" + suf)(_ptr(f3), M, N, O, _ptr(al), int(is_map), rho, tau0, sigma0, opnorm,
                                                  int(accel), maxiter, init_mode, nthreads, _ptr(u))
    if rc != 0:
        raise RuntimeError(f"oracle_pdps failed: {rc}")
    return u


def grad_np(u):
    """∇ as slices (S4) — independent of the C loops."""
    g1 = np.zeros_like(u)
    g2 = np.zeros_like(u)
    g1[:-1, :] = u[1:, :] - u[:-1, :]
    g2[:, :-1] = u[:, 1:] - u[:, :-1]
    return g1, g2


def grad_T_np(y1, y2):
    v = np.zeros_like(y1)
    M, N = y1.shape
    if M > 1:
        v[1:-1, :] = y1[:-2, :] - y1[1:-1, :]
        v[0, :] = -y1[0, :]
        v[-1, :] = y1[-2, :]
    w = np.zeros_like(y1)
    if N > 1:
        w[:, 1:-1] = y2[:, :-2] - y2[:, 1:-1]
        w[:, 0] = -y2[:, 0]
        w[:, -1] = y2[:, -2]
    return v + w


def pdps_numpy(f, alpha, *, maxiter=5000, tau0=5.0, sigma0=0.99 / 5, rho=0.0, accel=True,
               opnorm=OPNORM, init_mode=0, return_dual=False):
    """Independent numpy restatement of the recursion for one
    2-D image in fp64; used only to cross-check the C port."""
    b = np.asarray(f, dtype=np.float64)
    x = b.copy() if init_mode else np.zeros_like(b)
    y1 = np.zeros_like(b)
    y2 = np.zeros_like(b)
    al = np.asarray(alpha, dtype=np.float64)
    sigma = sigma0 / opnorm
    tau = tau0 / opnorm
    gamma = 1.0
    for _ in range(maxiter):
        omega = 1.0 / np.sqrt(1.0 + 2.0 * gamma * tau) if accel else 1.0
        dx = grad_T_np(y1, y2)
        xb = x
        x = (x - tau * (dx - b)) / (1.0 + tau)
        xb = (1.0 + omega) * x - omega * xb
        d1, d2 = grad_np(xb)
        y1 = y1 + sigma * d1
        y2 = y2 + sigma * d2
        if rho != 0.0:
            den = 1.0 + sigma * rho / al
            y1 = y1 / den
            y2 = y2 / den
        n2 = y1 * y1 + y2 * y2
        over = n2 > al * al
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.where(over, al / np.sqrt(n2), 1.0)
        y1 = np.where(over, y1 * s, y1)
        y2 = np.where(over, y2 * s, y2)
        if accel:
            tau, sigma = tau * omega, sigma / omega
    if return_dual:
        return x, y1, y2
    return x


def cost(u, ubar):
    """0.5*norm₂²(u-ū): /root/reference/src/TVLearningFunctionVec.jl:20."""
    d = np.asarray(u, dtype=np.float64) - np.asarray(ubar, dtype=np.float64)
    return 0.5 * float(np.sum(d * d))


# --------------------------------------------------------------------------
# Literal sparse-matrix gradient path
# --------------------------------------------------------------------------
def grad_matrix(M, N=None):
    """matrix(op, n) (S5): sparse 2MN×MN forward-difference matrix on the
    column-major vec; rows 0:MN = component 1 (along i), MN:2MN = component 2."""
    N = M if N is None else N

    def D(n):
        d = sp.diags([-np.ones(n), np.ones(n - 1)], [0, 1], shape=(n, n), format="lil")
        d[n - 1, :] = 0
        return d.tocsr()

    G1 = sp.kron(sp.identity(N), D(M), format="csr")
    G2 = sp.kron(D(N), sp.identity(M), format="csr")
    return sp.vstack([G1, G2], format="csr")


def xi(Gu):
    """Pixel 2-norm duplicated to length 2n² (S6)."""
    n2 = Gu.size // 2
    nrm = np.sqrt(Gu[:n2] ** 2 + Gu[n2:] ** 2)
    return np.concatenate([nrm, nrm])


def prodesc(a, b):
    """[diag(a1 b1) diag(a1 b2); diag(a2 b1) diag(a2 b2)] (S6)."""
    n2 = a.size // 2
    a1, a2, b1, b2 = a[:n2], a[n2:], b[:n2], b[n2:]
    return sp.bmat([[sp.diags(a1 * b1), sp.diags(a1 * b2)],
                    [sp.diags(a2 * b1), sp.diags(a2 * b2)]], format="csr")


def scalarprod(a, b):
    n2 = a.size // 2
    return a[:n2] * b[:n2] + a[n2:] * b[n2:]


def _solve(A, b, refine=0):
    """Sparse direct solve (SuperLU), optionally followed by `refine` steps of
    iterative refinement with residuals in x87 extended precision."""
    A = A.tocsc()
    lu = spla.splu(A)
    x = lu.solve(b)
    if refine:
        Al = A.astype(np.longdouble).tocsr()
        bl = b.astype(np.longdouble)
        xl = x.astype(np.longdouble)
        for _ in range(refine):
            r = bl - Al @ xl
            xl = xl + lu.solve(np.asarray(r, dtype=np.float64)).astype(np.longdouble)
        x = np.asarray(xl, dtype=np.float64)
    return x


EPS = float(np.finfo(np.float64).eps)


def gradient_scalar(alpha, u, ubar, refine=0, return_p=False):
    """gradient(α::Real, op, u::2D, ū::2D): TVLearningFunctionVec.jl:98-135."""
    n = u.shape[0]
    assert u.shape == (n, n), "reference assumes square images (:102)"
    uv = np.asarray(u, dtype=np.float64).flatten(order="F")
    ub = np.asarray(ubar, dtype=np.float64).flatten(order="F")
    G = grad_matrix(n)
    Gu = G @ uv
    nGu = xi(Gu)
    act = (nGu < 1e-12).astype(np.float64)
    inact = 1.0 - act
    Act, Inact = sp.diags(act), sp.diags(inact)
    den = Inact @ nGu + act
    Den = sp.diags(1.0 / den)
    prodKuKu = prodesc(Gu / den ** 3, Gu)
    I = sp.identity(n * n)
    Adj = sp.bmat([[I, -G.T],
                   [Act @ G + Inact @ (alpha * (Den - prodKuKu)) @ G, Inact + EPS * Act]],
                  format="csc")
    Track = np.concatenate([uv - ub, np.zeros(2 * n * n)])
    mult = _solve(Adj, Track, refine)
    p = mult[: n * n]
    g = float(np.sum(scalarprod(G @ p, Inact @ (Den @ Gu))))
    if return_p:
        return -g, p
    return -g


def gradient_reg_scalar(alpha, u, ubar, gamma=1e8, refine=0, return_p=False):
    """gradient_reg(α::Real, op, u::2D, ū::2D): TVLearningFunctionVec.jl:137-161."""
    n = u.shape[0]
    assert u.shape == (n, n)
    uv = np.asarray(u, dtype=np.float64).flatten(order="F")
    ub = np.asarray(ubar, dtype=np.float64).flatten(order="F")
    G = grad_matrix(n)
    Gu = G @ uv
    nGu = xi(Gu)
    act1 = nGu - 1.0 / gamma
    act = (np.maximum(0.0, act1) != 0).astype(np.float64)
    inact = 1.0 - act
    Act, Inact = sp.diags(act), sp.diags(inact)
    den = Act @ nGu + inact
    Den = sp.diags(1.0 / den)
    prodGuGu = prodesc(Gu / den ** 3, Gu)
    I = sp.identity(n * n)
    B = gamma * Inact
    C = Act @ (prodGuGu - Den)
    A = I + alpha * (G.T @ (B - C) @ G)
    p = _solve(A, ub - uv, refine)
    g = float(np.sum(scalarprod(G @ p, Act @ (Den @ Gu) + gamma * (Inact @ Gu))))
    if return_p:
        return g, p
    return g


def gradient_patch(alpha_map, grid_shape, u, ubar, refine=0):
    """gradient(α::AbstractArray, op, pOp, u::2D, ū::2D): :219-254; `alpha_map`
    is p(α), the M×N up-sampled map (:171)."""
    n = u.shape[0]
    assert u.shape == (n, n)
    uv = np.asarray(u, dtype=np.float64).flatten(order="F")
    ub = np.asarray(ubar, dtype=np.float64).flatten(order="F")
    av = np.asarray(alpha_map, dtype=np.float64).flatten(order="F")
    G = grad_matrix(n)
    Gu = G @ uv
    nGu = xi(Gu)
    act = (nGu < 1e-12).astype(np.float64)
    inact = 1.0 - act
    Act, Inact = sp.diags(act), sp.diags(inact)
    den = Inact @ nGu + act
    Den = sp.diags(1.0 / den)
    prodKuKu = prodesc(Gu / den ** 3, Gu)
    I = sp.identity(n * n)
    A2 = sp.diags(np.concatenate([av, av]))
    Adj = sp.bmat([[I, -G.T],
                   [Act @ G + Inact @ A2 @ (Den - prodKuKu) @ G,
                    Inact + np.sqrt(EPS) * Act]], format="csc")
    Track = np.concatenate([uv - ub, np.zeros(2 * n * n)])
    mult = _solve(Adj, Track, refine)
    p = mult[: n * n]
    g = -scalarprod(G @ p, Inact @ (Den @ Gu))
    g = g.reshape((n, n), order="F")
    return patch_adjoint(g, *grid_shape)


def gradient_reg_patch(alpha_map, grid_shape, u, ubar, gamma=1e8, refine=0):
    """gradient_reg(α::AbstractArray, op, pOp, u, ū): :192-215 (row-scaled,
    non-symmetric system `I + α[:] .* G'*(B-C)*G`, :212)."""
    m_, n_ = u.shape
    assert m_ == n_
    n = m_
    uv = np.asarray(u, dtype=np.float64).flatten(order="F")
    ub = np.asarray(ubar, dtype=np.float64).flatten(order="F")
    av = np.asarray(alpha_map, dtype=np.float64).flatten(order="F")
    G = grad_matrix(n)
    Gu = G @ uv
    nGu = xi(Gu)
    act1 = nGu - 1.0 / gamma
    act = (np.maximum(0.0, act1) != 0).astype(np.float64)
    inact = 1.0 - act
    Act, Inact = sp.diags(act), sp.diags(inact)
    den = Act @ nGu + inact
    Den = sp.diags(1.0 / den)
    prodGuGu = prodesc(Gu / den ** 3, Gu)
    I = sp.identity(n * n)
    B = gamma * Inact
    C = Act @ (prodGuGu - Den)
    A = I + sp.diags(av) @ (G.T @ (B - C) @ G)
    p = _solve(A, ub - uv, refine)
    g = p * (G.T @ (Act @ (Den @ Gu) + gamma * (Inact @ Gu)))
    g = g.reshape((n, n), order="F")
    return patch_adjoint(g, *grid_shape)


def tv_op_learning_function(x, data, Delta, Delta_t=1e-6, refine=0, u=None, **pdps_kw):
    """tv_op_learning_function(x,data,Δ;Δt=1e-6): TVLearningFunctionVec.jl:14-27.

    ``u`` may be supplied (e.g. the GPU's denoised stack) so the gradient
    solvers are compared on bit-identical u.
    """
    ubar = _fortran3(data[0], np.float64)
    f = _fortran3(data[1], np.float64)
    M, N, O = f.shape
    xa = np.asarray(x, dtype=np.float64)
    scalar = xa.ndim == 0
    if scalar:
        alpha = float(xa)
    else:
        alpha = patch_upsample(xa, M, N)
    if u is None:
        u = pdps(f, alpha, **pdps_kw)
    c = cost(u, ubar)
    if scalar:
        g = 0.0
        for i in range(O):  # serial sum, i ascending (:76-81, :89-94)
            if Delta > Delta_t:
                g += gradient_scalar(alpha, u[:, :, i], ubar[:, :, i], refine)
            else:
                g += gradient_reg_scalar(alpha, u[:, :, i], ubar[:, :, i], refine=refine)
    else:
        g = np.zeros(xa.shape)
        for i in range(O):  # :168-173, :183-188
            if Delta > Delta_t:
                g += gradient_patch(alpha, xa.shape, u[:, :, i], ubar[:, :, i], refine)
            else:
                g += gradient_reg_patch(alpha, xa.shape, u[:, :, i], ubar[:, :, i],
                                        refine=refine)
    return u, c, g


# --------------------------------------------------------------------------
# Dual (compliance-form) restatement of the four adjoint systems.
#
# Every variant (:98-135, :137-161, :192-215, :219-254) is, after eliminating
# the multipliers, (C + Gᵀ D G) p = r with a diagonal C > 0 and a block-diagonal
# D of per-pixel 2×2 tensors that are either isotropic s·I ("iso": flat pixels)
# or rank one s·t tᵀ with t ⟂ ∇u ("aniso").  Writing D = Bᵀ-modes with
# compliances E = 1/s gives the equivalent SPD system in multiplier space
#     (diag(E) + B C⁻¹ Bᵀ) ζ = B C⁻¹ r,      p = C⁻¹ (r - Bᵀ ζ),
# whose entries are all O(1) — no 1/eps or γ penalties — so a banded Cholesky in
# fp64 is accurate where the penalty form is not.  This is the
# formulation the CUDA path factorises; the function below is its CPU checker.
# --------------------------------------------------------------------------
def dual_setup(variant, alpha, u, ubar, gamma=1e8, act_tol=1e-12, eps_act=None):
    """variant ∈ {'reg','nonreg'}; alpha scalar or M×N map.  Returns a dict of the
    per-pixel / per-node quantities of the dual system."""
    n = u.shape[0]
    assert u.shape == (n, n)
    N = n * n
    uv = np.asarray(u, dtype=np.float64).flatten(order="F")
    tv = np.asarray(ubar, dtype=np.float64).flatten(order="F")
    patch = np.ndim(alpha) == 2
    av = np.asarray(alpha, dtype=np.float64).flatten(order="F") if patch else np.full(N, float(alpha))
    G = grad_matrix(n)
    G1, G2 = G[:N], G[N:]
    g1, g2 = G1 @ uv, G2 @ uv
    nrm = np.sqrt(g1 * g1 + g2 * g2)
    if variant == "reg":
        iso = ~(np.maximum(0.0, nrm - 1.0 / gamma) != 0)
        an = ~iso
        safe = np.where(an, nrm, 1.0)
        w1 = np.where(an, g1 / safe, gamma * g1)
        w2 = np.where(an, g2 / safe, gamma * g2)
        if patch:   # (diag(1/a) + GᵀDG) p = (ū-u)/a
            cinv, rc = av.copy(), tv - uv
            E = np.where(iso, 1.0 / gamma, nrm)
            kind = "node"
        else:
            cinv, rc = np.ones(N), tv - uv
            E = np.where(iso, 1.0 / (av * gamma), nrm / av)
            kind = "pixel"
        sign = 1.0
    else:
        if eps_act is None:
            eps_act = np.sqrt(EPS) if patch else EPS
        iso = nrm < act_tol
        an = ~iso
        safe = np.where(an, nrm, 1.0)
        w1 = np.where(an, g1 / safe, 0.0)
        w2 = np.where(an, g2 / safe, 0.0)
        cinv, rc = np.ones(N), uv - tv
        E = np.where(iso, eps_act, nrm / av)
        kind = "pixel"
        sign = -1.0
    ea = np.where(iso, 1.0, -g2 / safe)
    eb = np.where(iso, 0.0, g1 / safe)
    return dict(n=n, N=N, G1=G1, G2=G2, iso=iso, ea=ea, eb=eb, E=E, w1=w1, w2=w2,
                cinv=cinv, rc=rc, sign=sign, kind=kind)


def _chol_band_guard(ab, guard):
    """In-place lower banded Cholesky (ab[r, j] = A[j+r, j]) with a pivot floor: the C loop of liboracle.so
    (oracle_chol_band_guard), bit-identical to the numpy loop below and 10-50 × faster at 128² / 256²."""
    assert ab.flags.c_contiguous and ab.dtype == np.float64
    L = lib()
    L.oracle_chol_band_guard.restype = ctypes.c_longlong
    L.oracle_chol_band_guard.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_longlong, ctypes.c_double]
    g = L.oracle_chol_band_guard(ab.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ab.shape[0] - 1, ab.shape[1], float(guard))
    if g < 0:
        raise MemoryError("oracle_chol_band_guard")
    return int(g)


def _chol_band_guard_py(ab, guard):
    """The same factorisation as a numpy loop (kept as the cross-check of the C loop: tests/test_oracle.py)."""
    bw = ab.shape[0] - 1
    Nd = ab.shape[1]
    guarded = 0
    for j in range(Nd):
        d = ab[0, j]
        if not (d > guard):
            d = guard
            guarded += 1
        d = np.sqrt(d)
        ab[0, j] = d
        m = min(bw, Nd - 1 - j)
        if m == 0:
            continue
        l = ab[1:m + 1, j] / d
        ab[1:m + 1, j] = l
        for b in np.nonzero(l)[0] + 1:
            ab[0:m - b + 1, j + b] -= l[b - 1:m] * l[b - 1]
    return guarded


def gradient_dual(variant, alpha, u, ubar, grid_shape=None, refine=1, guard_rel=1e-13, **kw):
    """Gradient through the dual banded-Cholesky formulation (CPU checker of the
    CUDA solver).  Returns a float (scalar α) or an m×n array (patch α)."""
    import scipy.linalg as sla

    s = dual_setup(variant, alpha, u, ubar, **kw)
    N, n, iso = s["N"], s["n"], s["iso"]
    nm = np.where(iso, 2, 1)
    off = np.concatenate([[0], np.cumsum(nm)])
    B1 = sp.diags(s["ea"]) @ s["G1"] + sp.diags(s["eb"]) @ s["G2"]
    B2 = s["G2"][iso]            # second mode of an iso pixel: e = (0, 1)
    Bfull = sp.vstack([B1, B2]).tocsr()
    idx = np.concatenate([off[:-1], off[:-1][iso] + 1])
    perm = np.argsort(idx)
    B = Bfull[perm]
    Evec = np.concatenate([s["E"], s["E"][iso]])[perm]
    A = (sp.diags(Evec) + B @ sp.diags(s["cinv"]) @ B.T).tocoo()
    bw = int(np.max(np.abs(A.row - A.col)))
    ab = np.zeros((bw + 1, A.shape[0]))
    m = A.row >= A.col
    ab[A.row[m] - A.col[m], A.col[m]] = A.data[m]
    _chol_band_guard(ab, guard_rel * float(np.max(s["cinv"])))
    b = B @ s["rc"]
    zeta = sla.cho_solve_banded((ab, True), b)
    for _ in range(refine):
        p = s["rc"] - s["cinv"] * (B.T @ zeta)
        res = B @ p - Evec * zeta
        zeta = zeta + sla.cho_solve_banded((ab, True), res)
    p = s["rc"] - s["cinv"] * (B.T @ zeta)
    if s["kind"] == "pixel":
        fpix = s["sign"] * ((s["G1"] @ p) * s["w1"] + (s["G2"] @ p) * s["w2"])
    else:
        fpix = s["sign"] * p * (s["G1"].T @ s["w1"] + s["G2"].T @ s["w2"])
    if grid_shape is None:
        return float(np.sum(fpix))
    return patch_adjoint(fpix.reshape((n, n), order="F"), *grid_shape)
