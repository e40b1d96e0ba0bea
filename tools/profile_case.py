"""Small fixed cases for ncu (one GPU, short): `python tools/profile_case.py pdps|grad|resident`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "pdps"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
arith = bp.FAST if os.environ.get("BPLTV_ARITH", "strict") == "fast" else bp.STRICT
with bp.Context([0], int(os.environ.get("BPLTV_PREC", "64"))) as ctx:
    if case == "pdps":      # BASELINE config 4 shape, a few iterations of the streaming kernel
        t, f = bp.synthetic_dataset(512, 512, 64, seed=20240601)
        u = ctx.denoise(f, 0.1, bp.pdps_opts(maxiter=iters, kernel=bp.KERNEL_MARCH, arith=arith))
        print("pdps ok", float(u.mean()), ctx.stats()["ms_pdps"])
    elif case == "tblock":  # same shape through the temporally blocked kernel (depth from BPLTV_TBLOCK_T)
        t, f = bp.synthetic_dataset(512, 512, 64, seed=20240601)
        u = ctx.denoise(f, 0.1, bp.pdps_opts(maxiter=iters, kernel=bp.KERNEL_TBLOCK, arith=arith,
                                             tblock=int(os.environ.get("BPLTV_TBLOCK_T", "4"))))
        print("tblock ok", float(u.mean()), ctx.stats()["ms_pdps"])
    elif case == "resident":
        t, f = bp.synthetic_dataset(128, 128, 10, seed=7)
        u = ctx.denoise(f, 0.1, bp.pdps_opts(maxiter=iters, kernel=bp.KERNEL_RESIDENT, arith=arith))
        print("resident ok", float(u.mean()), ctx.stats()["ms_pdps"])
    else:                   # one learn_eval on 2 images 128×128 with few PDPS iterations
        data = bp.synthetic_dataset(128, 128, 2, seed=7)
        ctx.set_dataset(data)
        u, c, g = ctx.learn_eval(0.1, 0.1, bp.eval_opts(bp.pdps_opts(maxiter=iters)))
        print("grad ok", c, g, ctx.stats()["ms_gradient"])
