"""Per-evaluation wall time of the config-3 learn run (patch λ), repeated: where does the run-to-run variance come from?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
from bpldenoising_b200 import trbox  # noqa: E402
z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
name = sys.argv[1] if len(sys.argv) > 1 else "circle_128_10"
t = np.asfortranarray(z[name + "/true"].astype(np.float64) / z[name + "/true_div"])
f = np.asfortranarray(z[name + "/data"].astype(np.float64) / z[name + "/data_div"])
x0 = 1e-4 * np.ones((2, 2))
with bp.Context([0], 64) as c:
    c.set_dataset((t, f))
    for rep in range(4):
        trace = []
        def lf(xx, d_, D):
            t0 = time.perf_counter()
            out = bp.tv_op_learning_function(xx, d_, D, ctx=c)
            st = c.stats()
            trace.append(((time.perf_counter() - t0) * 1e3, st["ms_pdps"], st["ms_gradient"], st["ms_total"], D))
            return out
        t0 = time.perf_counter()
        res = trbox.bilevel_learn((t, f), lf, x0, dict(Delta0=1e-4))
        print("run %d: %.3f s, %d evaluations" % (rep, time.perf_counter() - t0, res.evaluations))
        print("   wall ms:", " ".join("%.1f" % a[0] for a in trace))
        print("   pdps+grad:", " ".join("%.1f" % (a[1] + a[2]) for a in trace))
        print("   Delta:", " ".join("%.0e" % a[4] for a in trace))
