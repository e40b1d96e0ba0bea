"""Cycle attribution inside grad_factor_kernel (developer build with -DBPLTV_FACTOR_TIMING, see
tools/_timing/): prints, for lane 0 of warps 0 / 1 / 4 / 15, the average cycles per block step spent in
the panel phase, waiting at the mid-step barrier, in the look-ahead (warp 0) or tile (other warps) work,
and waiting at the end-of-step barrier."""
import os, sys
import numpy as np
os.environ["BPLTV_LIB"] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_timing", "libbpltv_timing.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bpldenoising_b200 as bp  # noqa: E402
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "datasets.npz"))
t = np.asfortranarray(z["cameraman_128_5/true"].astype(float) / 255); f = np.asfortranarray(z["cameraman_128_5/data"].astype(float) / 255)
with bp.Context([0], 64) as c:
    c.set_dataset((t, f))
    u = c.denoise(None, 0.1)
    for reg in (False, True):
        print("== gradient_reg" if reg else "== gradient", flush=True)
        g = c.gradient(0.1, u, reg)
        print("   value", g, "ms", c.stats()["ms_gradient"], flush=True)
