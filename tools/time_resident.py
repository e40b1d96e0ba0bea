"""Kernel B vs its temporally blocked variant (BPLTV_RESIDENT_TB) on the reference's 128×128 configurations."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402

for O, its in ((1, 5000), (10, 5000), (10, 10000), (4, 5000)):
    t, f = bp.synthetic_dataset(128, 128, O, seed=3)
    res = {}
    for prec in (64, 32):
        for tb in ("0", "1"):
            os.environ["BPLTV_RESIDENT_TB"] = tb
            bp.reload_env()
            with bp.Context([0], prec) as c:
                ms = []
                for k in range(3):
                    u = c.denoise(f, 0.1, bp.pdps_opts(maxiter=its))
                    ms.append(c.stats()["ms_pdps"])
                res[(prec, tb)] = (min(ms), u, c.stats()["pdps_kernel_used"])
        same = np.array_equal(res[(prec, "0")][1], res[(prec, "1")][1])
        print("%2d x 128x128, %5d its, fp%d: kernel B %.2f ms, blocked %.2f ms (kernel id %d), identical %s" %
              (O, its, prec, res[(prec, "0")][0], res[(prec, "1")][0], res[(prec, "1")][2], same), flush=True)
os.environ.pop("BPLTV_RESIDENT_TB")
