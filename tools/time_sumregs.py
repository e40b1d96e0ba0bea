"""Streaming throughput of the sum-of-regularisers solve at BASELINE config 4's shape (64 × 512×512):
Gpixel-iter/s and the fraction of its HBM roofline (23 words = 184 B per pixel-iteration in fp64)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
peak = 6542.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
t, f = bp.synthetic_dataset(512, 512, 64, seed=20240601)
x = np.array([0.03, 0.02, 0.04])
for prec in (64, 32):
    with bp.Context([0], prec) as c:
        c.set_dataset((t, f))
        for arith, an in ((bp.STRICT, "strict"), (bp.FAST, "fast")):
            best = 1e30
            for rep in range(3):
                c.sumregs_denoise(None, x, bp.sumregs_pdps_opts(maxiter=iters, arith=arith))
                best = min(best, c.stats()["ms_pdps"])
            gp = 512 * 512 * 64 * iters / best / 1e6
            print("prec %d %-6s %.4f ms/iter  %.1f Gpixel-iter/s  %.3f of the %d B/pixel-iter HBM roofline" %
                  (prec, an, best / iters, gp, gp * 23 * prec / 8 / peak, 23 * prec // 8), flush=True)
