import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bpldenoising_b200 as bp
x = np.array([0.03, 0.02, 0.04])
for k in (1, 4):
    t, f = bp.synthetic_dataset(256, 256, k, seed=3)
    with bp.Context([0], 64) as c:
        c.set_dataset((t, f))
        u = np.asfortranarray(c.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=500)))
        best = 1e9
        for _ in range(3):
            g = c.sumregs_gradient(x, u, regularised=False); best = min(best, c.stats()["ms_gradient"])
        st = c.stats()
        print("256x256 x%d sumregs_gradient (nested dissection, 8-column steps on the top levels): %.1f ms, relres %.1e, %d launches" % (k, best, st["solver_max_relres"], st["kernel_launches"]), flush=True)
