import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import bpldenoising_b200 as bp
z = np.load("tests/golden/datasets.npz")
t = np.asfortranarray(z["cameraman_128_5/true"].astype(float)/255); f = np.asfortranarray(z["cameraman_128_5/data"].astype(float)/255)
with bp.Context([0], 64) as c:
    c.set_dataset((t, f))
    u = c.denoise(None, 0.1, bp.pdps_opts(maxiter=300))
    print(c.gradient(0.1, u, False))
    us = c.sumregs_denoise(None, np.array([0.001]*3), bp.sumregs_pdps_opts(maxiter=50))
    print(c.sumregs_gradient(np.array([0.001]*3), us, False))
