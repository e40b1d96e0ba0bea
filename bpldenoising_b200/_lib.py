"""ctypes binding of libbpltv.so (include/bpltv.h).  Fails loudly when the CUDA
library is missing — there is no CPU path in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BPLTV_LIB", os.path.join(_HERE, "libbpltv.so"))

STRICT, FAST = 0, 1
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_MARCH, KERNEL_RESIDENT, KERNEL_TBLOCK = 0, 1, 2, 3, 4


class PdpsOpts(C.Structure):
    _fields_ = [("tau0", C.c_double), ("sigma0", C.c_double), ("rho", C.c_double),
                ("opnorm", C.c_double), ("accel", C.c_int), ("maxiter", C.c_int),
                ("init_mode", C.c_int), ("arith", C.c_int), ("kernel", C.c_int),
                ("tblock", C.c_int), ("reserved", C.c_int * 4)]


class EvalOpts(C.Structure):
    _fields_ = [("pdps", PdpsOpts), ("delta_t", C.c_double), ("gamma", C.c_double),
                ("act_tol", C.c_double), ("eps_act", C.c_double), ("solver_tol", C.c_double),
                ("solver_maxit", C.c_int), ("solver", C.c_int), ("force_branch", C.c_int),
                ("reserved0", C.c_int), ("gamma_patch", C.c_double), ("reserved", C.c_int * 2)]


class Stats(C.Structure):
    _fields_ = [("ms_upload", C.c_double), ("ms_pdps", C.c_double), ("ms_cost", C.c_double),
                ("ms_gradient", C.c_double), ("ms_download", C.c_double), ("ms_total", C.c_double),
                ("pdps_iterations", C.c_longlong), ("pixel_iterations", C.c_longlong),
                ("solver_iterations", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("solver_max_relres", C.c_double), ("pdps_kernel_used", C.c_int),
                ("n_devices", C.c_int), ("tblock_depth", C.c_int), ("reserved", C.c_int * 5)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("reserved")}


# every symbol include/bpltv.h declares
EXPORTS = [
    "bpltv_default_pdps_opts", "bpltv_default_eval_opts", "bpltv_create", "bpltv_destroy",
    "bpltv_set_dataset", "bpltv_denoise", "bpltv_learn_eval", "bpltv_gradient", "bpltv_sweep", "bpltv_default_sumregs_eval_opts",
    "bpltv_sumregs_denoise", "bpltv_sumregs_learn_eval", "bpltv_sumregs_gradient",
    "bpltv_denoise_device", "bpltv_set_dataset_device", "bpltv_learn_eval_device",
    "bpltv_get_stats", "bpltv_last_error", "bpltv_version", "bpltv_reload_env", "bpltv_selftest",
    "bpltv_comm_unique_id", "bpltv_comm_init", "bpltv_comm_destroy",
]

_lib = None


class BpltvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbpltv error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  bpldenoising_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.bpltv_default_pdps_opts.argtypes = [C.POINTER(PdpsOpts)]
    L.bpltv_default_pdps_opts.restype = None
    L.bpltv_default_eval_opts.argtypes = [C.POINTER(EvalOpts)]
    L.bpltv_default_eval_opts.restype = None
    L.bpltv_create.argtypes = [ip, C.c_int, C.c_int, C.POINTER(vp)]
    L.bpltv_destroy.argtypes = [vp]
    L.bpltv_set_dataset.argtypes = [vp, dp, dp, C.c_int, C.c_int, C.c_int]
    L.bpltv_denoise.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int,
                                C.POINTER(PdpsOpts), dp]
    L.bpltv_learn_eval.argtypes = [vp, dp, C.c_int, C.c_int, C.c_double, C.POINTER(EvalOpts),
                                   dp, dp, dp]
    L.bpltv_gradient.argtypes = [vp, dp, dp, C.c_int, C.c_int, C.c_int, C.POINTER(EvalOpts), dp]
    L.bpltv_sweep.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, C.POINTER(PdpsOpts), dp, dp, dp]
    L.bpltv_default_sumregs_eval_opts.argtypes = [C.POINTER(EvalOpts)]
    L.bpltv_default_sumregs_eval_opts.restype = None
    L.bpltv_sumregs_denoise.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int,
                                        C.POINTER(PdpsOpts), dp]
    L.bpltv_sumregs_learn_eval.argtypes = [vp, dp, C.c_int, C.c_int, C.c_double, C.POINTER(EvalOpts), dp, dp, dp]
    L.bpltv_sumregs_gradient.argtypes = [vp, dp, dp, C.c_int, C.c_int, C.c_int, C.POINTER(EvalOpts), dp]
    L.bpltv_denoise_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int,
                                       C.POINTER(PdpsOpts), vp, vp]
    L.bpltv_set_dataset_device.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
    L.bpltv_learn_eval_device.argtypes = [vp, dp, C.c_int, C.c_int, C.c_double,
                                          C.POINTER(EvalOpts), vp, vp, vp]
    L.bpltv_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.bpltv_selftest.argtypes = [vp, C.c_int, C.c_int, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
    L.bpltv_comm_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
    L.bpltv_comm_init.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]
    L.bpltv_comm_destroy.argtypes = [vp]
    L.bpltv_last_error.restype = C.c_char_p
    L.bpltv_version.restype = C.c_int
    L.bpltv_reload_env.restype = None
    L.bpltv_reload_env.argtypes = []
    for name in EXPORTS:
        if name not in ("bpltv_default_pdps_opts", "bpltv_default_eval_opts", "bpltv_default_sumregs_eval_opts",
                        "bpltv_last_error", "bpltv_reload_env"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise BpltvError(rc, load().bpltv_last_error().decode("utf-8", "replace"))


def reload_env():
    """Re-read the BPLTV_* developer switches (the library snapshots them when the first context is created)."""
    load().bpltv_reload_env()
