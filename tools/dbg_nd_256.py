"""Locate images whose nested-dissection adjoint solve misses the tolerance (developer tool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
O = int(sys.argv[1]) if len(sys.argv) > 1 else 128
its = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
t, f = bp.synthetic_dataset(256, 256, O, seed=7)
worst = []
with bp.Context([0], 64) as c:
    u = c.denoise(f, 0.1, bp.pdps_opts(maxiter=its))
    eo = bp.eval_opts(solver_tol=0.0)
    for i in range(O):
        c.set_dataset((t[:, :, i:i + 1].copy(order="F"), f[:, :, i:i + 1].copy(order="F")))
        g = c.gradient(0.1, u[:, :, i:i + 1].copy(order="F"), False, eo)
        rr = c.stats()["solver_max_relres"]
        gb = c.gradient(0.1, u[:, :, i:i + 1].copy(order="F"), False, bp.eval_opts(solver=1))
        worst.append((rr, i, g, gb))
        if rr > 1e-10 or abs(g - gb) > 1e-9 * abs(gb):
            print("image", i, "relres %.3e" % rr, "nd", g, "band", gb, flush=True)
    # all together
    c.set_dataset((t, f))
    try:
        g = c.gradient(0.1, u, False, eo)
        print("all together: relres %.3e g %r  sum of singles %r" % (c.stats()["solver_max_relres"], g, sum(w[2] for w in worst)))
    except Exception as e:
        print("all together failed:", e)
worst.sort(reverse=True)
print("worst singles:", [(w[1], "%.2e" % w[0]) for w in worst[:5]])
i = worst[0][1]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dbg_nd_256.npz"), u=u[:, :, i], t=t[:, :, i], f=f[:, :, i], i=i)
