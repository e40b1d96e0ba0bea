"""λ-gradient of the TV learning function (ms, device events): nested-dissection solver (default) vs the banded
factorisations of round 1 (eval_opts.solver = 1), both branches, on stacks shaped like BASELINE configs 1, 2, 3 and 5."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402

P = np.array([[0.05, 0.1], [0.08, 0.02]])
cases = (("1x128 scalar", 128, 1, 0.1, 5000), ("10x128 scalar", 128, 10, 0.1, 5000), ("1x128 patch", 128, 1, P, 5000),
         ("148x128 scalar", 128, 148, 0.1, 1000), ("32x256 scalar", 256, 32, 0.1, 1000), ("128x256 scalar", 256, 128, 0.1, 5000))
only = sys.argv[1:] and sys.argv[1]
for name, n, O, x, its in cases:
    if only and only not in name:
        continue
    data = bp.synthetic_dataset(n, n, O, seed=7)
    with bp.Context([0], 64) as c:
        c.set_dataset(data)
        for Delta, br in ((0.1, "gradient    "), (1e-7, "gradient_reg")):
            res = {}
            for solver in (2, 1):
                if solver == 1 and n == 256 and O > 32:
                    continue
                eo = bp.eval_opts(bp.pdps_opts(maxiter=its), solver=solver)
                ms = []
                for k in range(3):
                    u, cost, g = c.learn_eval(x, Delta, eo, return_u=False)
                    st = c.stats()
                    ms.append(st["ms_gradient"])
                res[solver] = (min(ms), np.asarray(g, dtype=np.float64).copy(), st["solver_max_relres"], st["ms_pdps"], st["kernel_launches"])
            d = np.abs(res[2][1] - res[1][1]).max() / np.abs(res[1][1]).max() if 1 in res else float("nan")
            print("%-16s %s: ND %.2f ms (relres %.1e, %d launches), band %s ms, rel diff %.1e, pdps %.1f ms" %
                  (name, br, res[2][0], res[2][2], res[2][4], ("%.2f" % res[1][0]) if 1 in res else "-", d, res[2][3]), flush=True)
