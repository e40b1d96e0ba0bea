"""Band-LU gradients against the cluster size of the factorisation (BPLTV_LU_CLUSTER), ms, device events."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
for name, k in (("cameraman_128_5", 1), ("faces_train_128_10", 10)):
    t = np.asfortranarray(z[name + "/true"][:, :, :k] / 255.0); f = np.asfortranarray(z[name + "/data"][:, :, :k] / 255.0)
    with bp.Context([0], 64) as c:
        c.set_dataset((t, f))
        xs = np.array([0.001, 0.001, 0.001])
        xp = 0.001 * np.ones((2, 2, 3)); xp[1, 0, :] *= 1.5
        us = c.sumregs_denoise(None, xs, bp.sumregs_pdps_opts(maxiter=2000))
        up = c.sumregs_denoise(None, xp, bp.sumregs_pdps_opts(maxiter=2000))
        ref = {}
        for cs in ("1", "2", "4", "8", "16", ""):
            if cs: os.environ["BPLTV_LU_CLUSTER"] = cs
            else: os.environ.pop("BPLTV_LU_CLUSTER", None)
            line = []
            for label, x, u in (("scalar reg", xs, us), ("patch reg", xp, up)):
                best = 1e30
                for rep in range(3):
                    g = c.sumregs_gradient(x, u, regularised=True)
                    best = min(best, c.stats()["ms_gradient"])
                ref.setdefault(label, g)
                line.append("%s %.1f ms%s" % (label, best, "" if np.array_equal(g, ref[label]) else " DIFFERENT"))
            print("%s x%d cluster %s: %s" % (name, k, cs or "auto", ", ".join(line)), flush=True)
os.environ["BPLTV_GRAD_REG_LU"] = "1"
for n, O, its in ((128, 1, 3000), (256, 8, 1000), (256, 32, 1000)):
    data = bp.synthetic_dataset(n, n, O, seed=7)
    with bp.Context([0], 64) as c:
        c.set_dataset(data)
        eo = bp.eval_opts(bp.pdps_opts(maxiter=its))
        for cs in ("1", ""):
            if cs: os.environ["BPLTV_LU_CLUSTER"] = cs
            else: os.environ.pop("BPLTV_LU_CLUSTER", None)
            best = 1e30
            for rep in range(3):
                c.learn_eval(0.1, 1e-7, eo); best = min(best, c.stats()["ms_gradient"])
            print("TV gradient_reg via LU %dx%d x%d cluster %s: %.1f ms" % (n, n, O, cs or "auto", best), flush=True)
