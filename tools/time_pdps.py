"""Device-resident timing of the streaming PDPS kernels at BASELINE config 4 shape:
`python tools/time_pdps.py [iters]` prints ms/iteration, Gpixel-iter/s and the fraction of the
single-pass HBM roofline for kernel A and kernel C at T = 2, 3, 4 (strict and fast arithmetic)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 240
prec = int(os.environ.get("BPLTV_PREC", "64"))
M, N, O = (int(v) for v in os.environ.get("BPLTV_SHAPE", "512,512,64").split(","))
peak = 6542.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
t, f = bp.synthetic_dataset(M, N, O, seed=20240601)
word = prec // 8
with bp.Context([0], prec) as ctx:
    ctx.set_dataset((t, f))
    for arith, an in ((bp.STRICT, "strict"), (bp.FAST, "fast")):
        for kern, depth in ((bp.KERNEL_MARCH, 1), (bp.KERNEL_TBLOCK, 2), (bp.KERNEL_TBLOCK, 3), (bp.KERNEL_TBLOCK, 4)):
            best = 1e30
            for rep in range(3):
                ctx.denoise(None, 0.1, bp.pdps_opts(maxiter=iters, kernel=kern, tblock=depth, arith=arith))
                best = min(best, ctx.stats()["ms_pdps"])
            gp = M * N * O * iters / best / 1e6
            print("prec %d %-6s T=%d  %.4f ms/iter  %.1f Gpixel-iter/s  frac of single-pass HBM roofline %.3f" %
                  (prec, an, depth, best / iters, gp, gp * 7 * word / peak), flush=True)
