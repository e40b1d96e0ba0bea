"""TEST INFRASTRUCTURE: binary128 (__float128) solves of the adjoint systems of
/root/reference/src/TVLearningFunctionVec.jl:98-254 (oracle/quad_adjoint.c) — the arbiter of the gradient parity bars.

`gradient_literal(..., assemble_quad=False)` is the reference's matrix (entries rounded to double in the reference's
operation order) solved exactly; `assemble_quad=True` the system its formulas mean; `gradient_compliance` the
multiplier-space form the CUDA path factorises, in binary128.  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-s", "-C", _HERE, "libquad.so"], check=True)
        _lib = C.CDLL(os.path.join(_HERE, "libquad.so"))
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _prep(alpha, u, ubar):
    n = u.shape[0]
    assert u.shape == (n, n)
    uv = np.ascontiguousarray(np.asarray(u, dtype=np.float64).flatten(order="F"))
    tv = np.ascontiguousarray(np.asarray(ubar, dtype=np.float64).flatten(order="F"))
    amap = None
    a = 0.0
    if np.ndim(alpha) == 2:
        amap = np.ascontiguousarray(np.asarray(alpha, dtype=np.float64).flatten(order="F"))
    else:
        a = float(alpha)
    return n, uv, tv, a, amap


def _finish(n, fpix, grid_shape):
    if grid_shape is None:
        return float(np.sum(fpix))
    return orc.patch_adjoint(fpix.reshape((n, n), order="F"), *grid_shape)


def gradient_literal(alpha, u, ubar, grid_shape=None, assemble_quad=False, act_tol=1e-12, eps_act=None, return_p=False):
    """gradient(α, op, u, ū) (:98-135; patch :219-254) by a binary128 band LU with partial pivoting."""
    n, uv, tv, a, amap = _prep(alpha, u, ubar)
    if eps_act is None:
        eps_act = np.sqrt(orc.EPS) if amap is not None else orc.EPS
    p, f = np.zeros(n * n), np.zeros(n * n)
    rc = _load().quad_gradient_literal(n, _p(uv), _p(tv), C.c_double(a), _p(amap), C.c_double(act_tol), C.c_double(eps_act),
                                       int(assemble_quad), _p(p), _p(f))
    if rc != 0:
        raise RuntimeError(f"quad_gradient_literal: {rc}")
    g = _finish(n, f, grid_shape)
    return (g, p) if return_p else g


def gradient_compliance(alpha, u, ubar, grid_shape=None, act_tol=1e-12, eps_act=None, return_p=False):
    """The multiplier-space form (diag(E) + B Bᵀ) ζ = B r, p = r − Bᵀζ in binary128."""
    n, uv, tv, a, amap = _prep(alpha, u, ubar)
    if eps_act is None:
        eps_act = np.sqrt(orc.EPS) if amap is not None else orc.EPS
    p, f = np.zeros(n * n), np.zeros(n * n)
    rc = _load().quad_gradient_compliance(n, _p(uv), _p(tv), C.c_double(a), _p(amap), C.c_double(act_tol),
                                          C.c_double(eps_act), _p(p), _p(f))
    if rc != 0:
        raise RuntimeError(f"quad_gradient_compliance: {rc}")
    g = _finish(n, f, grid_shape)
    return (g, p) if return_p else g


def gradient_reg(alpha, u, ubar, grid_shape=None, gamma=1e8, assemble_quad=False, return_p=False):
    """gradient_reg(α, op, u, ū) (:137-161; patch :192-215, row-scaled) in binary128."""
    n, uv, tv, a, amap = _prep(alpha, u, ubar)
    p, f = np.zeros(n * n), np.zeros(n * n)
    rc = _load().quad_gradient_reg(n, _p(uv), _p(tv), C.c_double(a), _p(amap), C.c_double(gamma), int(assemble_quad),
                                   _p(p), _p(f))
    if rc != 0:
        raise RuntimeError(f"quad_gradient_reg: {rc}")
    g = _finish(n, f, grid_shape)
    return (g, p) if return_p else g
