"""Host-side mirror of the reference's learning-function interface over libbpltv.

Same names, argument meaning and error behaviour as
/root/reference/src/TVLearningFunctionVec.jl (``tv_op_learning_function`` :14-27,
``denoise`` :45-70, ``gradient`` / ``gradient_reg`` :72-96,:163-190) and
/root/reference/src/BPLDenoising.jl (``TVDenoise`` :41-82, ``L2CostFunction``
:84-86), so that a trust-region driver written against the reference calls this
module unchanged.  Arrays are M×N×O float64, column-major like Julia's
(``np.asfortranarray``); every numerical operation happens in CUDA kernels behind
the C ABI.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import EvalOpts, PdpsOpts, Stats, check

_DP = C.POINTER(C.c_double)


def _stack(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2:
        a = a[:, :, None]
    if a.ndim != 3:
        raise ValueError("expected an M×N or M×N×O array")
    return np.asfortranarray(a)


def _lam(x):
    xa = np.asarray(x, dtype=np.float64)
    if xa.ndim == 0:
        return np.asfortranarray(xa.reshape(1, 1)), True
    if xa.ndim != 2:
        raise ValueError("λ must be a real number or a 2-D array (patch grid)")
    return np.asfortranarray(xa), False


def _lam3(x):
    """Sum-of-regularisers parameter: a 3-vector (`x::AbstractVector{Float64}`) or an m×n×3 array."""
    xa = np.asarray(x, dtype=np.float64)
    if xa.shape == (3,):
        return np.ascontiguousarray(xa), 1, 1
    if xa.ndim == 3 and xa.shape[2] == 3:
        return np.asfortranarray(xa), xa.shape[0], xa.shape[1]
    raise ValueError("the sum-of-regularisers parameter is a 3-vector or an m×n×3 array")


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_DP)


def pdps_opts(**kw) -> PdpsOpts:
    """`denoising_default_params` (:33-43) overridden by keyword arguments, like the
    reference's `denoising_default_params ⬿ kwargs`."""
    o = PdpsOpts()
    _lib.load().bpltv_default_pdps_opts(C.byref(o))
    alias = {"τ₀": "tau0", "σ₀": "sigma0", "ρ": "rho"}
    for k, v in kw.items():
        k = alias.get(k, k)
        if k in ("verbose_iter", "save_results", "save_iterations"):
            continue  # iterator / logging knobs of the reference: no effect on iterates (S8)
        if not hasattr(o, k) or k == "reserved":
            raise TypeError(f"unknown PDPS option {k!r}")
        setattr(o, k, type(getattr(o, k))(v))
    return o


def eval_opts(pdps: Optional[PdpsOpts] = None, **kw) -> EvalOpts:
    o = EvalOpts()
    _lib.load().bpltv_default_eval_opts(C.byref(o))
    if pdps is not None:
        o.pdps = pdps
    alias = {"Δt": "delta_t", "γ": "gamma"}
    for k, v in kw.items():
        k = alias.get(k, k)
        if not hasattr(o, k) or k in ("reserved", "reserved0", "pdps"):
            raise TypeError(f"unknown evaluation option {k!r}")
        setattr(o, k, type(getattr(o, k))(v))
    return o


def sumregs_eval_opts(pdps: Optional[PdpsOpts] = None, **kw) -> EvalOpts:
    """Defaults of the sum-of-regularisers path: opnorm √18 (S12), Δt = 1e-3, γ = 1e3
    (/root/reference/src/SumRegsLearningFunction.jl:8, :117)."""
    o = EvalOpts()
    _lib.load().bpltv_default_sumregs_eval_opts(C.byref(o))
    if pdps is not None:
        o.pdps = pdps
    alias = {"Δt": "delta_t", "γ": "gamma"}
    for k, v in kw.items():
        k = alias.get(k, k)
        if not hasattr(o, k) or k in ("reserved", "reserved0", "pdps"):
            raise TypeError(f"unknown evaluation option {k!r}")
        setattr(o, k, type(getattr(o, k))(v))
    return o


def sumregs_pdps_opts(**kw) -> PdpsOpts:
    o = sumregs_eval_opts().pdps
    for k, v in kw.items():
        if not hasattr(o, k) or k == "reserved":
            raise TypeError(f"unknown PDPS option {k!r}")
        setattr(o, k, type(getattr(o, k))(v))
    return o


def _point_at_nccl():
    """libbpltv binds NCCL at run time ($BPLTV_NCCL_LIB, else libnccl.so.2 by soname).  Where no system copy is on the
    loader path, point it at the copy pip puts beside torch (nvidia/nccl/lib) — a path, not an import of torch."""
    if os.environ.get("BPLTV_NCCL_LIB"):
        return
    import importlib.util
    spec = importlib.util.find_spec("nvidia")
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["BPLTV_NCCL_LIB"] = cand
            return


class Context:
    """Owns a libbpltv context (streams, device buffers, resident dataset)."""

    def __init__(self, devices: Optional[Sequence[int]] = None, precision: int = 64):
        self._L = _lib.load()
        devs = list(devices) if devices is not None else [0]
        arr = (C.c_int * len(devs))(*devs)
        h = C.c_void_p()
        check(self._L.bpltv_create(arr, len(devs), precision, C.byref(h)))
        self._h = h
        self.precision = precision
        self.devices = devs
        self.shape = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.bpltv_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- dataset -------------------------------------------------------------
    def set_dataset(self, data):
        truth, noisy = _stack(data[0]), _stack(data[1])
        if truth.shape != noisy.shape:
            raise ValueError("truth and noisy stacks differ in shape")
        M, N, O = noisy.shape
        check(self._L.bpltv_set_dataset(self._h, _ptr(truth), _ptr(noisy), M, N, O))
        self.shape = (M, N, O)

    # ---- lower-level solve -----------------------------------------------------
    def denoise(self, data, x, opts: Optional[PdpsOpts] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        lam, _ = _lam(x)
        o = opts if opts is not None else pdps_opts()
        if data is None:
            if self.shape is None:
                raise _lib.BpltvError(-4, "no resident dataset")
            M, N, O = self.shape
            fptr = None
        else:
            f = _stack(data)
            M, N, O = f.shape
            fptr = _ptr(f)
        if out is not None:
            if out.shape != (M, N, O) or out.dtype != np.float64 or not out.flags.f_contiguous:
                raise ValueError("out must be an M×N×O float64 Fortran-ordered array")
            u = out
        else:
            u = np.zeros((M, N, O), order="F")
        check(self._L.bpltv_denoise(self._h, fptr, M, N, O, _ptr(lam), lam.shape[0], lam.shape[1],
                                    C.byref(o), _ptr(u)))
        return u

    # ---- learning function -------------------------------------------------------
    def learn_eval(self, x, Delta, opts: Optional[EvalOpts] = None, return_u: bool = True):
        if self.shape is None:
            raise _lib.BpltvError(-4, "no resident dataset: call set_dataset first")
        lam, scalar = _lam(x)
        o = opts if opts is not None else eval_opts()
        M, N, O = self.shape
        u = np.zeros((M, N, O), order="F") if return_u else None
        cost = C.c_double()
        grad = np.zeros(lam.shape, order="F")
        check(self._L.bpltv_learn_eval(self._h, _ptr(lam), lam.shape[0], lam.shape[1], float(Delta),
                                       C.byref(o), _ptr(u) if return_u else None, C.byref(cost),
                                       _ptr(grad)))
        g = float(grad[0, 0]) if scalar else grad
        return u, cost.value, g

    def gradient(self, x, u, regularised: bool, opts: Optional[EvalOpts] = None):
        lam, scalar = _lam(x)
        o = opts if opts is not None else eval_opts()
        us = _stack(u)
        if self.shape is None or us.shape != self.shape:
            raise ValueError("u must have the shape of the resident dataset")
        grad = np.zeros(lam.shape, order="F")
        check(self._L.bpltv_gradient(self._h, _ptr(us), _ptr(lam), lam.shape[0], lam.shape[1],
                                     int(bool(regularised)), C.byref(o), _ptr(grad)))
        return float(grad[0, 0]) if scalar else grad

    # ---- sum-of-regularisers interface (SumRegsLearningFunction.jl) ---------------------
    def sumregs_denoise(self, data, x, opts: Optional[PdpsOpts] = None) -> np.ndarray:
        lam, lm, ln = _lam3(x)
        o = opts if opts is not None else sumregs_eval_opts().pdps
        if data is None:
            if self.shape is None:
                raise _lib.BpltvError(-4, "no resident dataset")
            M, N, O = self.shape
            fptr = None
        else:
            f = _stack(data)
            M, N, O = f.shape
            fptr = _ptr(f)
        u = np.zeros((M, N, O), order="F")
        check(self._L.bpltv_sumregs_denoise(self._h, fptr, M, N, O, _ptr(lam), lm, ln, C.byref(o), _ptr(u)))
        return u

    def sumregs_learn_eval(self, x, Delta, opts: Optional[EvalOpts] = None, return_u: bool = True):
        if self.shape is None:
            raise _lib.BpltvError(-4, "no resident dataset: call set_dataset first")
        lam, lm, ln = _lam3(x)
        o = opts if opts is not None else sumregs_eval_opts()
        M, N, O = self.shape
        u = np.zeros((M, N, O), order="F") if return_u else None
        cost = C.c_double()
        grad = np.zeros(np.asarray(x).shape, order="F")
        check(self._L.bpltv_sumregs_learn_eval(self._h, _ptr(lam), lm, ln, float(Delta), C.byref(o),
                                               _ptr(u) if return_u else None, C.byref(cost), _ptr(grad)))
        return u, cost.value, grad

    def sumregs_gradient(self, x, u, regularised: bool, opts: Optional[EvalOpts] = None):
        lam, lm, ln = _lam3(x)
        o = opts if opts is not None else sumregs_eval_opts()
        us = _stack(u)
        if self.shape is None or us.shape != self.shape:
            raise ValueError("u must have the shape of the resident dataset")
        grad = np.zeros(np.asarray(x).shape, order="F")
        check(self._L.bpltv_sumregs_gradient(self._h, _ptr(us), _ptr(lam), lm, ln, int(bool(regularised)),
                                             C.byref(o), _ptr(grad)))
        return grad

    # ---- λ-sweeps (cost curves, validation) ------------------------------------------
    def sweep(self, parameters, opts: Optional[PdpsOpts] = None, return_u: bool = False,
              return_sqerr: bool = False):
        """All parameter sets × the resident images as ONE batch of independent solves.
        `parameters`: sequence of L scalars, or of L equally shaped 2-D grids.  Returns
        costs (L,), and optionally per-image squared errors (O, L) and u (M, N, O, L)."""
        if self.shape is None:
            raise _lib.BpltvError(-4, "no resident dataset: call set_dataset first")
        ps = [np.asarray(p, dtype=np.float64) for p in parameters]
        if not ps:
            raise ValueError("empty parameter range")
        if ps[0].ndim == 0:
            lm = ln = 1
        elif ps[0].ndim == 2:
            lm, ln = ps[0].shape
        else:
            raise ValueError("parameters must be scalars or 2-D grids")
        if any(p.shape != ps[0].shape for p in ps):
            raise ValueError("all parameter sets must have the same shape")
        L = len(ps)
        lams = np.ascontiguousarray(np.stack([np.asfortranarray(p.reshape(lm, ln)).ravel(order="F") for p in ps]))
        o = opts if opts is not None else pdps_opts()
        M, N, O = self.shape
        costs = np.zeros(L)
        sq = np.zeros((O, L), order="F") if return_sqerr else None
        u = np.zeros((M, N, O, L), order="F") if return_u else None
        check(self._L.bpltv_sweep(self._h, _ptr(lams), L, lm, ln, C.byref(o), _ptr(costs),
                                  _ptr(sq) if return_sqerr else None, _ptr(u) if return_u else None))
        out = [costs]
        if return_sqerr:
            out.append(sq)
        if return_u:
            out.append(u)
        return out[0] if len(out) == 1 else tuple(out)

    def stats(self) -> dict:
        s = Stats()
        check(self._L.bpltv_get_stats(self._h, C.byref(s)))
        return s.asdict()

    # ---- one process per GPU: the job's collective inside the library (include/bpltv.h, bpltv_comm_*) -------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """The 128-byte NCCL id rank 0 creates and ships to the other ranks (any transport)."""
        _point_at_nccl()
        buf = (C.c_ubyte * 128)()
        check(_lib.load().bpltv_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        """Join the job's communicator (collective).  From then on learn_eval / learn_eval_device /
        sumregs_learn_eval return the loss and gradient summed over all ranks (one ncclAllReduce per evaluation)."""
        _point_at_nccl()
        if len(unique_id) != 128:
            raise ValueError("the NCCL id has 128 bytes")
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        check(self._L.bpltv_comm_init(self._h, nranks, rank, buf))

    def comm_destroy(self):
        check(self._L.bpltv_comm_destroy(self._h))

    def selftest(self, mode: int, count: int, seed: int = 1, what: int = 0) -> dict:
        """Arithmetic self-test (include/bpltv.h, bpltv_selftest): the strict kernels' projection scale against the IEEE
        operations on `count` generated operand pairs of the context's precision."""
        res = (C.c_ulonglong * 4)()
        check(self._L.bpltv_selftest(self._h, what, mode, count, seed, res))
        return {"took": int(res[0]), "mismatches": int(res[1]), "first_a_bits": int(res[2]) & ((1 << 63) - 1),
                "first_alpha_bits": int(res[3])}

    # ---- device-pointer entry points (torch tensors; plumbing only) --------------
    def denoise_device(self, d_noisy_ptr: int, M: int, N: int, O: int, x, d_out_ptr: int,
                       opts: Optional[PdpsOpts] = None, stream: int = 0):
        lam, _ = _lam(x)
        o = opts if opts is not None else pdps_opts()
        check(self._L.bpltv_denoise_device(self._h, C.c_void_p(d_noisy_ptr), M, N, O, _ptr(lam),
                                           lam.shape[0], lam.shape[1], C.byref(o),
                                           C.c_void_p(d_out_ptr), C.c_void_p(stream)))

    def set_dataset_device(self, d_truth_ptr: int, d_noisy_ptr: int, M: int, N: int, O: int,
                           stream: int = 0):
        check(self._L.bpltv_set_dataset_device(self._h, C.c_void_p(d_truth_ptr),
                                               C.c_void_p(d_noisy_ptr), M, N, O, C.c_void_p(stream)))
        self.shape = (M, N, O)

    def learn_eval_device(self, x, Delta, d_costgrad_ptr: int, opts: Optional[EvalOpts] = None,
                          d_u_ptr: int = 0, stream: int = 0):
        lam, _ = _lam(x)
        o = opts if opts is not None else eval_opts()
        check(self._L.bpltv_learn_eval_device(self._h, _ptr(lam), lam.shape[0], lam.shape[1],
                                              float(Delta), C.byref(o),
                                              C.c_void_p(d_u_ptr) if d_u_ptr else None,
                                              C.c_void_p(d_costgrad_ptr), C.c_void_p(stream)))


# ------------------------------------------------------------------------------
# module-level functions with the reference's names
# ------------------------------------------------------------------------------
_default_ctx: Optional[Context] = None
_resident = None      # (context, truth array, noisy array) of the last upload — strong references, compared with `is`


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def _ensure_resident(ctx: Context, data):
    """The dataset is constant across the ≤21 evaluations of a learn run
    (/root/reference/src/TRBox.jl:210,227): upload it once per (truth, noisy) pair.  The pair is recognised by object
    identity of arrays this module keeps alive (an `id()` of a freed temporary can be reused by the next one), so a new
    array object — a slice, a copy — is always uploaded again.  In-place edits of an array already uploaded are NOT seen:
    call `ctx.set_dataset(data)` after editing, like the Julia binding's `set_dataset!`."""
    global _resident
    if (_resident is None or _resident[0] is not ctx or _resident[1] is not data[0] or _resident[2] is not data[1]
            or ctx.shape is None):
        ctx.set_dataset(data)
        _resident = (ctx, data[0], data[1])


def denoise(data, x, op=None, ctx: Optional[Context] = None, **kwargs) -> np.ndarray:
    """denoise(data, x, op; kwargs...) (:45-70).  `op` is accepted for signature
    parity (the reference only ever passes FwdGradientOp(), :17)."""
    ctx = ctx or default_context()
    return ctx.denoise(data, x, pdps_opts(**kwargs))


def TVDenoise(data, parameter, ctx: Optional[Context] = None, **kwargs) -> np.ndarray:
    """TVDenoise(data, parameter) (/root/reference/src/BPLDenoising.jl:41-82): the same
    solve with maxiter = 10000."""
    kwargs.setdefault("maxiter", 10000)
    return denoise(data, parameter, ctx=ctx, **kwargs)


def L2CostFunction(u, true_) -> float:
    """0.5*norm₂²(u-true_) (/root/reference/src/BPLDenoising.jl:84-86) for host arrays
    the caller already holds (validation tables); the learning function's cost is
    reduced on the device."""
    d = np.asarray(u, dtype=np.float64) - np.asarray(true_, dtype=np.float64)
    return 0.5 * float(np.vdot(d, d))


def tv_op_learning_function(x, data, Δ, Δt: float = 1e-6, ctx: Optional[Context] = None, **kwargs):
    """tv_op_learning_function(x, data, Δ; Δt=1e-6, kwargs...) → (u, cost, grad) (:14-27)."""
    ctx = ctx or default_context()
    _ensure_resident(ctx, data)
    eo = eval_opts(pdps_opts(**kwargs), delta_t=Δt)
    return ctx.learn_eval(x, Δ, eo)


def gradient(α, op, u, ū, ctx: Optional[Context] = None):
    """gradient(α, op, u, ū) (:72-83 scalar, :163-175 patch)."""
    ctx = ctx or default_context()
    _ensure_resident(ctx, (ū, ū))
    return ctx.gradient(α, u, regularised=False)


def gradient_reg(α, op, u, ū, ctx: Optional[Context] = None):
    """gradient_reg(α, op, u, ū) (:85-96 scalar, :177-190 patch)."""
    ctx = ctx or default_context()
    _ensure_resident(ctx, (ū, ū))
    return ctx.gradient(α, u, regularised=True)


# ------------------------------------------------------------------------------
# λ-sweeps and validation (/root/reference/src/BPLDenoising.jl:92-111, :128-158, :176-178, :381-415)
# ------------------------------------------------------------------------------
def generate_cost(data, parameter_range, num_samples: Optional[int] = None, ctx: Optional[Context] = None,
                  **kwargs) -> np.ndarray:
    """generate_cost(dataset, parameter_range, L2CostFunction, TVDenoise; num_samples) (:92-111)
    for a dataset already in memory: costs[i] = 0.5‖TVDenoise(data, parameter_range[i]) - true‖².
    The reference loops over the range; here the whole range is one batched launch."""
    ctx = ctx or default_context()
    truth, noisy = _stack(data[0]), _stack(data[1])
    if num_samples is not None:
        truth, noisy = np.asfortranarray(truth[:, :, :num_samples]), np.asfortranarray(noisy[:, :, :num_samples])
    ctx.set_dataset((truth, noisy))
    global _resident_key
    _resident_key = None
    kwargs.setdefault("maxiter", 10000)   # TVDenoise (:51)
    return ctx.sweep(list(parameter_range), pdps_opts(**kwargs))


def generate_scalar_tv_cost(data, parameter_range, num_samples: int = 1, ctx: Optional[Context] = None, **kwargs):
    """generate_scalar_tv_cost(dataset, parameter_range; num_samples=1) (:128-130)."""
    return generate_cost(data, parameter_range, num_samples=num_samples, ctx=ctx, **kwargs)


def generate_2d_tv_cost(data, parameter_range_1, parameter_range_2, num_samples: int = 1,
                        ctx: Optional[Context] = None, **kwargs) -> np.ndarray:
    """generate_2d_tv_cost (:136-158, :176-178): costs[i, j] for the 2×1 patch parameter
    α = [range_1[i]; range_2[j]] .* ones(2,1)."""
    grids = [np.array([[a], [b]], dtype=np.float64) for a in parameter_range_1 for b in parameter_range_2]
    costs = generate_cost(data, grids, num_samples=num_samples, ctx=ctx, **kwargs)
    return costs.reshape(len(parameter_range_1), len(parameter_range_2))


def validate_tv_parameter(parameter, data, ctx: Optional[Context] = None, **kwargs) -> dict:
    """validate_tv_parameter(parameter; dataset_name) (:381-415) for a dataset in memory:
    u = TVDenoise(noisy, parameter); cost; the per-image quality table (orig/out SSIM and
    PSNR) and its means.  PSNR comes from the device-side per-image squared errors."""
    from . import quality
    ctx = ctx or default_context()
    truth, noisy = _stack(data[0]), _stack(data[1])
    ctx.set_dataset((truth, noisy))
    global _resident_key
    _resident_key = None
    kwargs.setdefault("maxiter", 10000)
    costs, sq, u = ctx.sweep([parameter], pdps_opts(**kwargs), return_u=True, return_sqerr=True)
    u = np.asfortranarray(u[:, :, :, 0])
    M, N, O = truth.shape
    rows = []
    for i in range(O):
        rows.append({
            "img_num": i + 1,
            "orig_ssim": quality.assess_ssim(truth[:, :, i], noisy[:, :, i]),
            "orig_psnr": quality.assess_psnr(truth[:, :, i], noisy[:, :, i]),
            "out_ssim": quality.assess_ssim(truth[:, :, i], u[:, :, i]),
            "out_psnr": quality.psnr_from_sqerr(sq[i, 0], M * N),
        })
    return {"u": u, "cost": float(costs[0]), "table": rows,
            "mean_ssim": float(np.mean([r["out_ssim"] for r in rows])),
            "mean_psnr": float(np.mean([r["out_psnr"] for r in rows]))}


# ------------------------------------------------------------------------------
# sum-of-regularisers interface (/root/reference/src/SumRegsLearningFunction.jl)
# ------------------------------------------------------------------------------
def sumregs_denoise(data, x, op1=None, op2=None, op3=None, pOp=None, ctx: Optional[Context] = None, **kwargs):
    """sumregs_denoise(data, x, op₁, op₂, op₃[, pOp]) (:38-85); the operator arguments are accepted
    for signature parity (the reference always passes Fwd/Bwd/CenteredGradientOp, :9-11)."""
    ctx = ctx or default_context()
    return ctx.sumregs_denoise(data, x, sumregs_pdps_opts(**kwargs))


def sumregs_learning_function(x, data, Δ, Δt: float = 1e-3, ctx: Optional[Context] = None, **kwargs):
    """sumregs_learning_function(x, data, Δ; Δt=1e-3) → (u, cost, grad) (:8-36); x a 3-vector or m×n×3."""
    ctx = ctx or default_context()
    _ensure_resident(ctx, data)
    eo = sumregs_eval_opts(sumregs_pdps_opts(**kwargs), delta_t=Δt)
    return ctx.sumregs_learn_eval(x, Δ, eo)
