"""One gradient call per branch on an O×n×n stack (profiling target: ncu --kernel-name regex:nd_ ...)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
n = int(sys.argv[1]); O = int(sys.argv[2]); its = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
t, f = bp.synthetic_dataset(n, n, O, seed=7)
with bp.Context([0], 64) as c:
    u = c.denoise(f, 0.1, bp.pdps_opts(maxiter=its))
    c.set_dataset((t, f))
    for rep in range(2):
        for reg in (False, True):
            g = c.gradient(0.1, u, reg)
            print(n, O, "reg" if reg else "nonreg", g, c.stats()["ms_gradient"], flush=True)
