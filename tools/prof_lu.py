"""One patch sumregs_gradient_reg evaluation (band LU) on cameraman_128_5, for ncu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
t = np.asfortranarray(z["cameraman_128_5/true"][:, :, :1] / 255.0)
f = np.asfortranarray(z["cameraman_128_5/data"][:, :, :1] / 255.0)
with bp.Context([0], 64) as c:
    c.set_dataset((t, f))
    xp = 0.001 * np.ones((2, 2, 3)); xp[1, 0, :] *= 1.5
    up = c.sumregs_denoise(None, xp, bp.sumregs_pdps_opts(maxiter=200))
    g = c.sumregs_gradient(xp, up, regularised=True)
    print(g.ravel()[:3], c.stats()["ms_gradient"])
