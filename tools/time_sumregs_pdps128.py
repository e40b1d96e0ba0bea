"""sumregs_denoise at the reference's dataset size (128×128, 5000 iterations; ms, device events): the
cluster-resident kernel vs the streaming pair, 1 and 10 images."""
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
        x = np.array([0.001, 0.001, 0.001])
        out = {}
        for label, kern, env in (("stream", bp.KERNEL_GENERIC, None), ("resident cs<=8", bp.KERNEL_RESIDENT, "8"), ("resident cs<=16", bp.KERNEL_RESIDENT, "16")):
            if env: os.environ["BPLTV_RESIDENT_CS"] = env
            best = 1e30
            for rep in range(3):
                u = c.sumregs_denoise(None, x, bp.sumregs_pdps_opts(maxiter=5000, kernel=kern))
                best = min(best, c.stats()["ms_pdps"])
            out[label] = (best, u)
            os.environ.pop("BPLTV_RESIDENT_CS", None)
        same = all(np.array_equal(out["stream"][1], v[1]) for v in out.values())
        print("%s x%d: %s; identical %s" % (name, k, ", ".join("%s %.1f ms" % (l, v[0]) for l, v in out.items()), same), flush=True)
