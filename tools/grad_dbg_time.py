"""Timing decomposition of the TV factorisation with the kernel's debug switches (results are wrong
under them): BPLTV_GRAD_DBG bit 1 = no trailing tiles, 2 = no pivot chain, 4 = no panel."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import bpldenoising_b200 as bp
z = np.load("tests/golden/datasets.npz")
t = np.asfortranarray(z["cameraman_128_5/true"].astype(float)/255); f = np.asfortranarray(z["cameraman_128_5/data"].astype(float)/255)
with bp.Context([0], 64) as c:
    c.set_dataset((t, f))
    u = c.denoise(None, 0.1, bp.pdps_opts(maxiter=5000))
    for dbg in ("0", "1", "2", "4", "3", "7"):
        os.environ["BPLTV_GRAD_DBG"] = dbg
        ms = []
        for _ in range(3):
            try:
                c.gradient(0.1, u, False)
            except bp.BpltvError:
                pass
            ms.append(c.stats()["ms_gradient"])
        print("dbg", dbg, "gradient ms", ["%.2f" % m for m in ms], flush=True)
