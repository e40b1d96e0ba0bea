"""Wall time of the sum-of-regularisers gradients on the reference's datasets (ms, device events):
scalar sumregs_gradient_reg through the multiplier-space Cholesky vs the node-space band LU
(BPLTV_SUMREGS_REG_LU), the patch variant (band LU only), and the non-regularised branch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))


def data(name, k):
    return (np.asfortranarray(z[name + "/true"][:, :, :k] / 255.0), np.asfortranarray(z[name + "/data"][:, :, :k] / 255.0))


def timed(c, x, reg, u, reps=3):
    best, g = 1e30, None
    for _ in range(reps):
        g = c.sumregs_gradient(x, u, regularised=reg)
        best = min(best, c.stats()["ms_gradient"])
    return best, g


for name, k in (("cameraman_128_5", 1), ("faces_train_128_10", 10)):
    t, f = data(name, k)
    with bp.Context([0], 64) as c:
        c.set_dataset((t, f))
        xs = np.array([0.001, 0.001, 0.001])
        xp = 0.001 * np.ones((2, 2, 3)); xp[1, 0, :] *= 1.5
        us = c.sumregs_denoise(None, xs, bp.sumregs_pdps_opts(maxiter=2000))
        up = c.sumregs_denoise(None, xp, bp.sumregs_pdps_opts(maxiter=2000))
        os.environ["BPLTV_GRAD_SOLVER"] = "1"            # the banded solvers of round 1
        os.environ["BPLTV_SUMREGS_REG_LU"] = "0"
        bp.reload_env()
        ms_c, g_c = timed(c, xs, True, us)
        os.environ["BPLTV_SUMREGS_REG_LU"] = "1"
        bp.reload_env()
        ms_l, g_l = timed(c, xs, True, us)
        os.environ.pop("BPLTV_GRAD_SOLVER")
        os.environ.pop("BPLTV_SUMREGS_REG_LU")
        bp.reload_env()
        print("%s x%d scalar reg: Cholesky %.1f ms, band LU %.1f ms, rel diff %.2e" %
              (name, k, ms_c, ms_l, np.abs(g_c - g_l).max() / np.abs(g_c).max()), flush=True)
        ms_nd, g_nd = timed(c, xs, True, us)
        print("%s x%d scalar reg: nested dissection (W=2) %.2f ms, rel diff to band LU %.2e, relres %.1e, %d launches" %
              (name, k, ms_nd, np.abs(g_nd - g_l).max() / np.abs(g_l).max(), c.stats()["solver_max_relres"],
               c.stats()["kernel_launches"]), flush=True)
        ms_p, g_p = timed(c, xp, True, up)
        print("%s x%d patch reg (band LU): %.1f ms" % (name, k, ms_p), flush=True)
        ms_n, g_n = timed(c, xs, False, us)
        st_n = c.stats()
        ms_pn, g_pn = timed(c, xp, False, up)
        os.environ["BPLTV_GRAD_SOLVER"] = "1"
        bp.reload_env()
        ms_nb, g_nb = timed(c, xs, False, us, reps=2)
        ms_pnb, g_pnb = timed(c, xp, False, up, reps=2)
        os.environ.pop("BPLTV_GRAD_SOLVER")
        bp.reload_env()
        print("%s x%d non-reg: nested dissection scalar %.1f ms (relres %.1e, %d launches, %.0f MB per image), patch %.1f ms; "
              "band Cholesky %.1f / %.1f ms; rel diff %.1e / %.1e" %
              (name, k, ms_n, st_n["solver_max_relres"], st_n["kernel_launches"], st_n.get("grad_bytes_per_image", 0) / 1e6,
               ms_pn, ms_nb, ms_pnb, np.abs(g_n - g_nb).max() / np.abs(g_nb).max(),
               np.abs(g_pn - g_pnb).max() / np.abs(g_pnb).max()), flush=True)
t, f = bp.synthetic_dataset(256, 256, 2, seed=20240602)
with bp.Context([0], 64) as c:
    c.set_dataset((t, f))
    xp = 0.01 * np.ones((2, 2, 3)); xp[1, 0, :] *= 1.5
    up = c.sumregs_denoise(None, xp, bp.sumregs_pdps_opts(maxiter=500))
    ms_p, g_p = timed(c, xp, True, up, reps=2)
    print("synthetic 256x256 x2 patch reg (band LU): %.1f ms, finite %s" % (ms_p, bool(np.all(np.isfinite(g_p)))), flush=True)
# many images: scalar sumregs_gradient_reg on the nested-dissection solver vs the band LU
for n, k in ((128, 148), (256, 32)):
    t, f = bp.synthetic_dataset(n, n, k, seed=20240602)
    with bp.Context([0], 64) as c:
        c.set_dataset((t, f))
        xs = np.array([0.01, 0.01, 0.01])
        us = c.sumregs_denoise(None, xs, bp.sumregs_pdps_opts(maxiter=300))
        ms_nd, g_nd = timed(c, xs, True, us, reps=3)
        rr = c.stats()["solver_max_relres"]
        os.environ["BPLTV_GRAD_SOLVER"] = "1"
        bp.reload_env()
        try:
            ms_l, g_l = timed(c, xs, True, us, reps=2)
        except bp.BpltvError as e:
            ms_l, g_l = float("nan"), g_nd
            print("band LU refused:", e)
        os.environ.pop("BPLTV_GRAD_SOLVER")
        bp.reload_env()
        print("synthetic %dx%d x%d scalar reg: nested dissection %.1f ms (relres %.1e), band LU %.1f ms, rel diff %.2e" %
              (n, n, k, ms_nd, rr, ms_l, np.abs(g_nd - g_l).max() / np.abs(g_l).max()), flush=True)
