"""gradient_reg of the TV learning function (ms, device events): multiplier-space banded Cholesky vs the
node-space band LU (BPLTV_GRAD_REG_LU), on synthetic stacks shaped like BASELINE configs 1, 2, 3 and 5."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402

cases = (("1x128 scalar", 128, 1, 0.1, 5000), ("10x128 scalar", 128, 10, 0.1, 5000),
         ("1x128 patch", 128, 1, np.array([[0.05, 0.1], [0.08, 0.02]]), 5000),
         ("10x128 patch", 128, 10, np.array([[0.05, 0.1], [0.08, 0.02]]), 5000), ("148x128 scalar", 128, 148, 0.1, 1000),
         ("32x256 scalar", 256, 32, 0.1, 1000), ("128x256 scalar", 256, 128, 0.1, 500))
for name, n, O, x, its in cases:
    data = bp.synthetic_dataset(n, n, O, seed=7)
    with bp.Context([0], 64) as c:
        c.set_dataset(data)
        eo = bp.eval_opts(bp.pdps_opts(maxiter=its))
        res = {}
        for mode in ("0", "1"):
            os.environ["BPLTV_GRAD_REG_LU"] = mode
            ms = []
            for k in range(3):
                u, cost, g = c.learn_eval(x, 1e-7, eo)
                ms.append(c.stats()["ms_gradient"])
            res[mode] = (min(ms), np.asarray(g, dtype=np.float64).copy())
        os.environ.pop("BPLTV_GRAD_REG_LU")
        d = np.abs(res["0"][1] - res["1"][1]).max() / np.abs(res["0"][1]).max()
        print("%-16s gradient_reg: Cholesky %.1f ms, band LU %.1f ms, rel diff %.1e" % (name, res["0"][0], res["1"][0], d), flush=True)
