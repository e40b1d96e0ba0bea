"""What one multifrontal solve (forward + backward sweep, residual, update) costs inside a gradient: the same call with
1, 2 and 3 refinement steps (eval_opts.solver_maxit); device events, best of 3."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
from bpldenoising_b200 import learning as L  # noqa: E402
z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
t = np.asfortranarray(z["cameraman_128_5/true"][:, :, :1] / 255.0)
f = np.asfortranarray(z["cameraman_128_5/data"][:, :, :1] / 255.0)
with bp.Context([0], 64) as c:
    c.set_dataset((t, f))
    xs = np.array([0.001, 0.001, 0.001])
    us = c.sumregs_denoise(None, xs, bp.sumregs_pdps_opts(maxiter=2000))
    u = c.denoise(f, 0.1, bp.pdps_opts(maxiter=5000))
    for name, call in (("sumregs_gradient", lambda o: c.sumregs_gradient(xs, us, False, L.sumregs_eval_opts(solver_maxit=o))),
                       ("sumregs_gradient_reg", lambda o: c.sumregs_gradient(xs, us, True, L.sumregs_eval_opts(solver_maxit=o))),
                       ("gradient", lambda o: c.gradient(0.1, u, False, L.eval_opts(solver_maxit=o))),
                       ("gradient_reg", lambda o: c.gradient(0.1, u, True, L.eval_opts(solver_maxit=o)))):
        ms = []
        for o in (1, 2, 3):
            best = 1e30
            for _ in range(3):
                call(o)
                best = min(best, c.stats()["ms_gradient"])
            ms.append(best)
        print("%-22s refine 1/2/3: %.2f %.2f %.2f ms -> one solve+residual %.2f ms, rest (factorisation etc.) %.2f ms" %
              (name, ms[0], ms[1], ms[2], ms[1] - ms[0], ms[0] - 2 * (ms[1] - ms[0])), flush=True)
