"""BASELINE config 5's regularised branch on ONE GPU's share (128 images 256×256, 5000 inner iterations):
gradient_reg through the multiplier-space Cholesky vs the node-space band LU (ms, device events)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
O = int(sys.argv[1]) if len(sys.argv) > 1 else 128
data = bp.synthetic_dataset(256, 256, O, seed=20240602)
with bp.Context([0], 64) as c:
    c.set_dataset(data)
    eo = bp.eval_opts(bp.pdps_opts(maxiter=5000))
    u, cost, g = c.learn_eval(0.1, 1e-7, eo)
    print("PDPS %.1f ms" % c.stats()["ms_pdps"], flush=True)
    res = {}
    for mode in ("0", "1"):
        os.environ["BPLTV_GRAD_REG_LU"] = mode
        ms = []
        for k in range(2):
            gg = c.gradient(0.1, u, regularised=True)
            ms.append(c.stats()["ms_gradient"])
        res[mode] = (min(ms), float(gg))
    print("%d x 256x256, 5000 its: gradient_reg Cholesky %.1f ms, band LU %.1f ms, rel diff %.1e" %
          (O, res["0"][0], res["1"][0], abs(res["0"][1] - res["1"][1]) / abs(res["0"][1])), flush=True)
