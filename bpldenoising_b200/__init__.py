"""bpldenoising_b200 — B200-native (sm_100a) TV-denoising solve and λ-gradient behind
the learning-function interface of dvillacis/BPLDenoising.

Only the hot path lives here: the CUDA kernels + C ABI (`csrc/`, `libbpltv.so`,
`include/bpltv.h`) and the host-side mirror of the reference interface.  Importing
the package loads the CUDA library and fails loudly when it has not been built.
"""
from . import _lib
from ._lib import (BpltvError, reload_env, FAST, KERNEL_AUTO, KERNEL_GENERIC, KERNEL_MARCH, KERNEL_RESIDENT,
                   KERNEL_TBLOCK, STRICT)

_lib.load()  # no silent CPU path: ImportError if libbpltv.so is absent

from .learning import (Context, L2CostFunction, TVDenoise, default_context, denoise, eval_opts,  # noqa: E402
                       generate_2d_tv_cost, generate_cost, generate_scalar_tv_cost, gradient, gradient_reg,
                       pdps_opts, sumregs_denoise, sumregs_eval_opts, sumregs_learning_function, sumregs_pdps_opts,
                       tv_op_learning_function,
                       validate_tv_parameter)
from .datasets import load_dataset, synthetic_dataset, testdataset  # noqa: E402
from . import quality, results  # noqa: E402,F401
from .parallel import shard_range  # noqa: E402
from . import trbox  # noqa: E402,F401

__all__ = [
    "BpltvError", "reload_env", "Context", "L2CostFunction", "TVDenoise", "default_context", "denoise",
    "eval_opts", "gradient", "gradient_reg", "pdps_opts", "tv_op_learning_function",
    "synthetic_dataset", "shard_range", "generate_cost", "generate_scalar_tv_cost", "generate_2d_tv_cost",
    "validate_tv_parameter", "sumregs_denoise", "sumregs_learning_function", "sumregs_eval_opts", "sumregs_pdps_opts", "load_dataset", "testdataset", "quality", "STRICT", "FAST", "KERNEL_AUTO", "KERNEL_GENERIC",
    "KERNEL_MARCH", "KERNEL_RESIDENT", "KERNEL_TBLOCK",
]
