"""What the λ-gradient of an fp32 context achieves against the oracle on the same fp32 image (relative errors)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
from oracle import oracle, quad  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
t, f = (np.asfortranarray(z["faces_train_128_10/" + k][:64, :64, :2] / 255.0) for k in ("true", "data"))
t32 = t.astype(np.float32).astype(np.float64)
x = np.array([[0.03, 0.08], [0.05, 0.06]])
am = oracle.patch_upsample(x, 64, 64)
rel = lambda a, b: float(np.linalg.norm(np.atleast_1d(a) - np.atleast_1d(b)) / np.linalg.norm(np.atleast_1d(b)))
with bp.Context([0], 32) as c:
    c.set_dataset((t, f))
    eo = bp.eval_opts(bp.pdps_opts(maxiter=1500))
    for lam, a, grid in ((0.06, 0.06, None), (x, am, x.shape)):
        for Delta, variant in ((0.1, "nonreg"), (1e-7, "reg")):
            u, cost, g = c.learn_eval(lam, Delta, eo)
            ref = sum(oracle.gradient_dual(variant, a, u[:, :, i], t32[:, :, i], grid_shape=grid) for i in range(2))
            if variant == "nonreg":
                q = sum(quad.gradient_compliance(a, u[:, :, i], t32[:, :, i], grid_shape=grid) for i in range(2))
            else:
                q = sum(quad.gradient_reg(a, u[:, :, i], t32[:, :, i], grid_shape=grid) for i in range(2))
            print("patch" if grid else "scalar", variant, "gpu vs dual %.2e  gpu vs binary128 %.2e  dual vs binary128 %.2e  relres %.1e" %
                  (rel(g, ref), rel(g, q), rel(ref, q), c.stats()["solver_max_relres"]), flush=True)
