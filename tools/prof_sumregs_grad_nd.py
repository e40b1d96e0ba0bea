"""One sumregs_gradient call (non-regularised, nested dissection in multiplier space) on the reference's 128x128 data
(profiling target: ncu --kernel-name regex:nd ...).  argv: dataset images"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "cameraman_128_5"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
t = np.asfortranarray(z[name + "/true"][:, :, :k] / 255.0)
f = np.asfortranarray(z[name + "/data"][:, :, :k] / 255.0)
with bp.Context([0], 64) as c:
    c.set_dataset((t, f))
    xs = np.array([0.001, 0.001, 0.001])
    us = c.sumregs_denoise(None, xs, bp.sumregs_pdps_opts(maxiter=2000))
    for rep in range(2):
        g = c.sumregs_gradient(xs, us, regularised=False)
        st = c.stats()
        print(name, k, g, st["ms_gradient"], st["solver_max_relres"], flush=True)
