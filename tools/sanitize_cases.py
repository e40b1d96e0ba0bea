"""Small invocations of the kernels and host paths changed in round 2, for compute-sanitizer
(`compute-sanitizer --tool memcheck|synccheck|racecheck python tools/sanitize_cases.py`)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402

t, f = bp.synthetic_dataset(128, 128, 3, seed=7)
for prec in (64, 32):
    with bp.Context([0], prec) as c:
        u = c.denoise(f, 0.1, bp.pdps_opts(maxiter=7, kernel=bp.KERNEL_RESIDENT))              # async halo exchange, 16- and 8-CTA clusters
        u1 = c.denoise(f[:, :, :1].copy(order="F"), 0.1, bp.pdps_opts(maxiter=7, kernel=bp.KERNEL_RESIDENT))
        us = c.sumregs_denoise(f[:, :, :1].copy(order="F"), np.array([0.01, 0.02, 0.03]), bp.sumregs_pdps_opts(maxiter=6, kernel=bp.KERNEL_RESIDENT))
        ut = c.denoise(f, 0.1, bp.pdps_opts(maxiter=9, kernel=bp.KERNEL_TBLOCK, tblock=4))      # BallScale in the temporally blocked kernel
        print("prec", prec, "resident", float(u.mean()), float(u1.mean()), "sumregs", float(us.mean()), "tblock", float(ut.mean()), flush=True)
with bp.Context([0], 64) as c:
    t64, f64 = bp.synthetic_dataset(64, 64, 2, seed=3)
    c.set_dataset((t64, f64))
    eo = bp.eval_opts(bp.pdps_opts(maxiter=60))
    for Delta in (0.1, 1e-7):                                                                   # nested-dissection gradient, both branches
        _, cost, g = c.learn_eval(0.08, Delta, eo)
        print("learn_eval", Delta, cost, g, flush=True)
    big_t, big_f = bp.synthetic_dataset(512, 512, 8, seed=1)                                     # 16 MiB: staged pageable copies
    ub = c.denoise(big_f, 0.1, bp.pdps_opts(maxiter=4))
    print("staged", float(ub.mean()), c.stats()["ms_upload"], c.stats()["ms_download"], flush=True)
    print("selftest", c.selftest(2, 1 << 20), flush=True)
