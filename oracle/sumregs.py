"""CPU oracle for the sum-of-regularisers path — TEST INFRASTRUCTURE ONLY (see oracle.py).

Restates /root/reference/src/SumRegsLearningFunction.jl: `sumregs_denoise` (:38-85, which calls
the un-vendored `sumregs_denoise_pdps`), and the four adjoint systems literally
(`sumregs_gradient_reg` scalar :112-167, patch :195-262; `sumregs_gradient` scalar :264-327, patch
:330-407), each assembled as the reference assembles it and handed to a sparse direct solver, plus
the compliance-form ("dual") restatement the CUDA path factorises.

PARITY UNPINNED, like the TV path: `BwdGradientOp`, `CenteredGradientOp` and
`sumregs_denoise_pdps` live in un-vendored packages.  Assumptions (docs/SEMANTICS.md S10-S13):
  S10  ∇ᵇ: backward differences, zero in the FIRST row / column; adjoint = exact transpose.
  S11  ∇ᶜ: centred differences ½(u[i+1]-u[i-1]) on interior rows / columns, zero on the boundary;
       adjoint = exact transpose.
  S12  sumregs_denoise_pdps = the S1 recursion on K = (∇ᶠ; ∇ᵇ; ∇ᶜ) with one dual field and one
       projection radius α_k per operator; Δx = (∇ᶠᵀy¹ + ∇ᵇᵀy²) + ∇ᶜᵀy³; R_K = √(8+8+2).
  S13  matrix(op, n) of ∇ᵇ / ∇ᶜ stacks the two components like S5.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import oracle as _o

OPNORM3 = float(np.sqrt(18.0))   # S12
EPS = _o.EPS
KINDS = ("fwd", "bwd", "ctr")


# --------------------------------------------------------------------------
# operators as sparse matrices on the column-major vec (S5, S10, S11, S13)
# --------------------------------------------------------------------------
def _d1(kind, n):
    d = sp.lil_matrix((n, n))
    for i in range(n):
        if kind == "fwd" and i + 1 < n:
            d[i, i], d[i, i + 1] = -1.0, 1.0
        elif kind == "bwd" and i >= 1:
            d[i, i], d[i, i - 1] = 1.0, -1.0
        elif kind == "ctr" and 1 <= i <= n - 2:
            d[i, i + 1], d[i, i - 1] = 0.5, -0.5
    return d.tocsr()


def op_matrix(kind, M, N=None):
    N = M if N is None else N
    G1 = sp.kron(sp.identity(N), _d1(kind, M), format="csr")
    G2 = sp.kron(_d1(kind, N), sp.identity(M), format="csr")
    return sp.vstack([G1, G2], format="csr")


# --------------------------------------------------------------------------
# lower-level solve (S12), one IEEE operation per operator, in this order
# --------------------------------------------------------------------------
def _shift(a, di, dj):
    """b[i,j] = a[i+di, j+dj], zero outside the image."""
    M, N = a.shape
    b = np.zeros_like(a)
    si = slice(max(0, -di), min(M, M - di)); sj = slice(max(0, -dj), min(N, N - dj))
    ti = slice(max(0, di), min(M, M + di)); tj = slice(max(0, dj), min(N, N + dj))
    b[si, sj] = a[ti, tj]
    return b


def _grad(kind, xb):
    M, N = xb.shape
    half = xb.dtype.type(0.5)
    d1 = np.zeros_like(xb); d2 = np.zeros_like(xb)
    if kind == "fwd":
        d1[:-1, :] = xb[1:, :] - xb[:-1, :]
        d2[:, :-1] = xb[:, 1:] - xb[:, :-1]
    elif kind == "bwd":
        d1[1:, :] = xb[1:, :] - xb[:-1, :]
        d2[:, 1:] = xb[:, 1:] - xb[:, :-1]
    else:
        d1[1:-1, :] = half * (xb[2:, :] - xb[:-2, :])
        d2[:, 1:-1] = half * (xb[:, 2:] - xb[:, :-2])
    return d1, d2


def _grad_T(kind, y1, y2):
    half = y1.dtype.type(0.5)
    if kind == "fwd":     # (y1[i-1]-y1[i]) + (y2[j-1]-y2[j]); the duals are zero where ∇ᶠ is
        return (_shift(y1, -1, 0) - y1) + (_shift(y2, 0, -1) - y2)
    if kind == "bwd":     # (y1[i]-y1[i+1]) + (y2[j]-y2[j+1])
        return (y1 - _shift(y1, 1, 0)) + (y2 - _shift(y2, 0, 1))
    return half * (_shift(y1, -1, 0) - _shift(y1, 1, 0)) + half * (_shift(y2, 0, -1) - _shift(y2, 0, 1))


def sumregs_pdps(f, alphas, *, maxiter=5000, tau0=5.0, sigma0=0.99 / 5, accel=True, opnorm=OPNORM3,
                 init_mode=0, dtype=np.float64):
    """sumregs_denoise(data, x, op₁, op₂, op₃[, pOp]) (:38-85).  `alphas`: three scalars, or three
    M×N maps (p(x)[:,:,k], :63-68).  Returns M×N×O (Fortran)."""
    f3 = _o._fortran3(f, dtype)
    M, N, O = f3.shape
    al = [np.asarray(a, dtype=dtype) for a in alphas]
    steps = _o.step_sizes(maxiter, tau0, sigma0, opnorm, accel)
    one = dtype(1)
    out = np.zeros_like(f3, order="F")
    for o in range(O):
        b = np.ascontiguousarray(f3[:, :, o])
        x = b.copy() if init_mode else np.zeros_like(b)
        y = [[np.zeros_like(b), np.zeros_like(b)] for _ in range(3)]
        for it in range(maxiter):
            tau, sigma, omega = (dtype(v) for v in steps[it])
            dx = (_grad_T("fwd", *y[0]) + _grad_T("bwd", *y[1])) + _grad_T("ctr", *y[2])
            xn = (x - tau * (dx - b)) / (one + tau)
            xb = (one + omega) * xn - omega * x
            x = xn
            for k, kind in enumerate(KINDS):
                d1, d2 = _grad(kind, xb)
                v1 = y[k][0] + sigma * d1
                v2 = y[k][1] + sigma * d2
                n2 = v1 * v1 + v2 * v2
                outside = n2 > al[k] * al[k]
                with np.errstate(divide="ignore", invalid="ignore"):   # n2 = 0 lanes are discarded below
                    sc = al[k] / np.sqrt(n2)
                    y[k][0] = np.where(outside, v1 * sc, v1)
                    y[k][1] = np.where(outside, v2 * sc, v2)
        out[:, :, o] = x
    return out


# --------------------------------------------------------------------------
# literal adjoint systems
# --------------------------------------------------------------------------
def _sets_reg(G, uv, gamma):
    Gu = G @ uv
    nGu = _o.xi(Gu)
    act = (np.maximum(0.0, nGu - 1.0 / gamma) != 0).astype(np.float64)
    inact = 1.0 - act
    Act, Inact = sp.diags(act), sp.diags(inact)
    den = Act @ nGu + inact
    Den = sp.diags(1.0 / den)
    prod = _o.prodesc(Gu / den ** 3, Gu)
    B = gamma * Inact
    C = Act @ (prod - Den)
    w = Act @ (Den @ Gu) + gamma * (Inact @ Gu)
    return B - C, w


def _sets_nonreg(G, uv):
    Gu = G @ uv
    nGu = _o.xi(Gu)
    act = (nGu < 1e-12).astype(np.float64)
    inact = 1.0 - act
    Act, Inact = sp.diags(act), sp.diags(inact)
    den = Inact @ nGu + act
    Den = sp.diags(1.0 / den)
    prod = _o.prodesc(Gu / den ** 3, Gu)
    return Act, Inact, Den, prod, Gu


def _vecs(u, ubar):
    n = u.shape[0]
    assert u.shape == (n, n), "the reference assumes square images"
    return n, np.asarray(u, dtype=np.float64).flatten(order="F"), np.asarray(ubar, dtype=np.float64).flatten(order="F")


def sumregs_gradient_reg(x, u, ubar, grid_shape=None, gamma=None, refine=0):
    """sumregs_gradient_reg: scalar `x` = 3-vector (:112-167, γ = 1e3) or three M×N maps with
    `grid_shape` = (m, n) (:195-262, γ = 1e8, row-scaled system)."""
    n, uv, ub = _vecs(u, ubar)
    patch = grid_shape is not None
    gamma = (1e8 if patch else 1e3) if gamma is None else gamma
    Gs = [op_matrix(k, n) for k in KINDS]
    A = sp.identity(n * n, format="csr")
    ws = []
    for k in range(3):
        BmC, w = _sets_reg(Gs[k], uv, gamma)
        T = Gs[k].T @ BmC @ Gs[k]
        if patch:
            A = A + sp.diags(np.asarray(x[k], dtype=np.float64).flatten(order="F")) @ T   # x₁[:] .* G₁'*(B₁-C₁)*G₁ (:246)
        else:
            A = A + float(x[k]) * T
        ws.append(w)
    p = _o._solve(A, ub - uv, refine)
    if not patch:
        return np.array([p @ (Gs[k].T @ ws[k]) for k in range(3)])
    gx = np.zeros(tuple(grid_shape) + (3,))
    for k in range(3):
        g = (p * (Gs[k].T @ ws[k])).reshape((n, n), order="F")
        gx[:, :, k] = _o.patch_adjoint(g, *grid_shape)
    return gx


def sumregs_gradient(x, u, ubar, grid_shape=None, refine=0, eps_act=EPS):
    """sumregs_gradient: scalar (:264-327) or patch (:330-407); 7n² saddle-point system."""
    n, uv, ub = _vecs(u, ubar)
    patch = grid_shape is not None
    N = n * n
    Gs = [op_matrix(k, n) for k in KINDS]
    rows = [[sp.identity(N)] + [-G.T for G in Gs]]
    ws = []
    for k in range(3):
        Act, Inact, Den, prod, Gu = _sets_nonreg(Gs[k], uv)
        if patch:
            av = np.asarray(x[k], dtype=np.float64).flatten(order="F")
            blk = Act @ Gs[k] + Inact @ sp.diags(np.concatenate([av, av])) @ (Den - prod) @ Gs[k]
        else:
            blk = Act @ Gs[k] + Inact @ (float(x[k]) * (Den - prod)) @ Gs[k]
        row = [blk, None, None, None]
        row[1 + k] = Inact + eps_act * Act
        rows.append(row)
        ws.append(Inact @ (Den @ Gu))
    Adj = sp.bmat(rows, format="csc")
    Track = np.concatenate([uv - ub, np.zeros(6 * N)])
    p = _o._solve(Adj, Track, refine)[:N]
    if not patch:
        return -np.array([p @ (Gs[k].T @ ws[k]) for k in range(3)])
    gx = np.zeros(tuple(grid_shape) + (3,))
    for k in range(3):
        g = (-p * (Gs[k].T @ ws[k])).reshape((n, n), order="F")
        gx[:, :, k] = _o.patch_adjoint(g, *grid_shape)
    return gx


def sumregs_learning_function(x, data, Delta, Delta_t=1e-3, refine=0, u=None, **pdps_kw):
    """sumregs_learning_function(x, data, Δ; Δt=1e-3) (:8-36): x a 3-vector or an m×n×3 array."""
    ubar = _o._fortran3(data[0], np.float64)
    f = _o._fortran3(data[1], np.float64)
    M, N, O = f.shape
    xa = np.asarray(x, dtype=np.float64)
    patch = xa.ndim == 3
    if patch:
        maps = [_o.patch_upsample(xa[:, :, k], M, N) for k in range(3)]
        grid = xa.shape[:2]
    else:
        maps, grid = [float(v) for v in xa], None
    if u is None:
        u = sumregs_pdps(f, maps, **pdps_kw)
    c = _o.cost(u, ubar)
    g = np.zeros(xa.shape)
    for i in range(O):   # serial sum, i ascending (:90-97, :102-109)
        if Delta > Delta_t:
            g = g + sumregs_gradient(maps, u[:, :, i], ubar[:, :, i], grid, refine)
        else:
            g = g + sumregs_gradient_reg(maps, u[:, :, i], ubar[:, :, i], grid, refine=refine)
    return u, c, g


# --------------------------------------------------------------------------
# compliance-form ("dual") restatement: the algorithm of the CUDA path on the CPU
# --------------------------------------------------------------------------
def sumregs_gradient_dual(variant, x, u, ubar, grid_shape=None, gamma=None, act_tol=1e-12, eps_act=EPS,
                          refine=1, guard_rel=1e-13):
    """(C + Σ_k G_kᵀ D_k G_k) p = r with per-pixel tensors D_kq = s·I (flat) or s·t tᵀ (t ⟂ ∇_k u),
    solved as (diag(E) + B C⁻¹ Bᵀ) ζ = B C⁻¹ r, p = C⁻¹(r − Bᵀ ζ), E = 1/s, modes numbered
    pixel-major (operator-minor) so that the matrix is banded.  variant ∈ {'reg', 'nonreg'}."""
    import scipy.linalg as sla

    n, uv, ub = _vecs(u, ubar)
    N = n * n
    patch = grid_shape is not None
    if patch and variant == "reg":
        # (I + Σ_k diag(a_k) G_kᵀ(B_k−C_k)G_k) p = r (:246) is row-scaled by a DIFFERENT map per
        # operator: unlike the TV case (one map: divide the rows by it) it cannot be symmetrised,
        # so it has no compliance form with an SPD matrix.  Only the literal solve exists.
        raise NotImplementedError("patch sumregs_gradient_reg has no symmetric compliance form")
    gamma = 1e3 if gamma is None else gamma
    rowsB, Es, key, Ws = [], [], [], []
    cinv = np.ones(N)
    for k, kind in enumerate(KINDS):
        G = op_matrix(kind, n)
        G1, G2 = G[:N], G[N:]
        g1, g2 = G1 @ uv, G2 @ uv
        nrm = np.sqrt(g1 * g1 + g2 * g2)
        av = np.asarray(x[k], dtype=np.float64).flatten(order="F") if patch else np.full(N, float(x[k]))
        if variant == "reg":
            iso = ~(np.maximum(0.0, nrm - 1.0 / gamma) != 0)
            safe = np.where(iso, 1.0, nrm)
            w1 = np.where(iso, gamma * g1, g1 / safe); w2 = np.where(iso, gamma * g2, g2 / safe)
            E = np.where(iso, 1.0 / (av * gamma), nrm / av)
        else:
            iso = nrm < act_tol
            safe = np.where(iso, 1.0, nrm)
            w1 = np.where(iso, 0.0, g1 / safe); w2 = np.where(iso, 0.0, g2 / safe)
            E = np.where(iso, eps_act, nrm / av)
        ea = np.where(iso, 1.0, -g2 / safe); eb = np.where(iso, 0.0, g1 / safe)
        rowsB.append(sp.diags(ea) @ G1 + sp.diags(eb) @ G2); Es.append(E); key.append(np.arange(N) * 6 + 2 * k)
        rowsB.append(G2[iso]); Es.append(E[iso]); key.append(np.arange(N)[iso] * 6 + 2 * k + 1)
        Ws.append((G1, G2, w1, w2))
    perm = np.argsort(np.concatenate(key))
    B = sp.vstack(rowsB).tocsr()[perm]
    Evec = np.concatenate(Es)[perm]
    if variant == "reg":
        rc, sign = ub - uv, 1.0
    else:
        rc, sign = uv - ub, -1.0
    A = (sp.diags(Evec) + B @ sp.diags(cinv) @ B.T).tocoo()
    bw = int(np.max(np.abs(A.row - A.col)))
    ab = np.zeros((bw + 1, A.shape[0]))
    m = A.row >= A.col
    np.add.at(ab, (A.row[m] - A.col[m], A.col[m]), A.data[m])
    _o._chol_band_guard(ab, guard_rel)
    b = B @ (cinv * rc)
    zeta = sla.cho_solve_banded((ab, True), b)
    for _ in range(refine):
        p = cinv * (rc - B.T @ zeta)
        zeta = zeta + sla.cho_solve_banded((ab, True), B @ p - Evec * zeta)
    p = cinv * (rc - B.T @ zeta)
    out = []
    for (G1, G2, w1, w2) in Ws:
        if patch:
            fpix = -p * (G1.T @ w1 + G2.T @ w2)                      # (:395-397)
        else:
            fpix = sign * ((G1 @ p) * w1 + (G2 @ p) * w2)
        out.append(fpix)
    if not patch:
        return np.array([float(np.sum(fp)) for fp in out])
    gx = np.zeros(tuple(grid_shape) + (3,))
    for k in range(3):
        gx[:, :, k] = _o.patch_adjoint(out[k].reshape((n, n), order="F"), *grid_shape)
    return gx
