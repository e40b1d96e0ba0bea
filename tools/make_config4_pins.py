"""Write tests/golden/config4_pins.json: per-image SHA-256 of the ORACLE's denoised stack at
BASELINE config 4's full size (64 × 512×512, λ = 0.1, 1000 iterations; fp64 and fp32) plus the
loss, so that the GPU box can check bit-exactness at full size without running the oracle for
minutes.  Oracle-derived pins (they guard the CUDA path and the oracle against drift); they are not
outputs of the Julia reference (parity unpinned).  ~2 minutes on 8 cores."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from bpldenoising_b200.datasets import synthetic_dataset  # noqa: E402  (pure numpy; no CUDA needed)

M, N, O, ITERS, LAM, SEED = 512, 512, 64, 1000, 0.1, 20240601
truth, noisy = synthetic_dataset(M, N, O, seed=SEED)
out = {"M": M, "N": N, "O": O, "iterations": ITERS, "lambda": LAM, "seed": SEED,
       "noisy_sha256": hashlib.sha256(noisy.tobytes(order="F")).hexdigest()}
for name, dt in (("f64", np.float64), ("f32", np.float32)):
    u = orc.pdps(noisy, LAM, maxiter=ITERS, dtype=dt, nthreads=os.cpu_count())
    out[name] = {"cost": float(orc.cost(u.astype(np.float64), truth)),
                 "u_sha256": [hashlib.sha256(np.ascontiguousarray(u[:, :, o].T).tobytes()).hexdigest() for o in range(O)]}
    print(name, out[name]["cost"], out[name]["u_sha256"][0][:16], flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "config4_pins.json"), "w"), indent=0)
