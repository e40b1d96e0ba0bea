"""scalar_bilevel_sumregs_learn / patch_bilevel_sumregs_learn on the reference's dataset (cameraman_128_5, one sample,
BPLDenoising.jl:306-314, :423-481) with the nested-dissection gradients (default) and with the banded solvers of round 1
(BPLTV_GRAD_SOLVER=1): wall time of the learn run, per-evaluation gradient time and launches (≈ 70+: nested dissection;
< 10: a band solver — the multiplier form falls back when a front exceeds shared memory)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
from bpldenoising_b200 import trbox, learning  # noqa: E402
z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
t = np.asfortranarray(z["cameraman_128_5/true"][:, :, :1] / 255.0)
f = np.asfortranarray(z["cameraman_128_5/data"][:, :, :1] / 255.0)
for solver in ("0", "1"):
    os.environ["BPLTV_GRAD_SOLVER"] = solver
    bp.reload_env()
    with bp.Context([0], 64) as c:
        for name, fn in (("scalar", trbox.scalar_bilevel_sumregs_learn), ("patch", trbox.patch_bilevel_sumregs_learn)):
            log = []
            orig = learning.sumregs_learning_function

            def wrapped(x, ds, D, ctx=None, **kw):
                r = orig(x, ds, D, ctx=ctx, **kw)
                st = ctx.stats()
                log.append((st["ms_pdps"], st["ms_gradient"], st["kernel_launches"]))
                return r
            learning.sumregs_learning_function = wrapped
            try:
                for rep in range(2):
                    log.clear()
                    t0 = time.perf_counter()
                    res = fn((t, f), ctx=c)
                    dt = time.perf_counter() - t0
            finally:
                learning.sumregs_learning_function = orig
            g = np.array([l[1] for l in log]); p = np.array([l[0] for l in log]); k = np.array([l[2] for l in log])
            print("solver %s %-6s: %.3f s, %d evaluations, cost %.6f, x mean %.5f; per evaluation: solve %.1f ms, gradient %.1f ms "
                  "(min %.1f max %.1f), launches min %d max %d" %
                  (solver, name, dt, res.evaluations, res.log[-1].function_value, float(np.mean(res.x)), p.mean(), g.mean(),
                   g.min(), g.max(), k.min(), k.max()), flush=True)
os.environ.pop("BPLTV_GRAD_SOLVER")
