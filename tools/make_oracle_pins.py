"""Write tests/golden/oracle_pins.npz: oracle-derived known answers on
datasets/cameraman_128_5 (λ=0.1, 5000 iterations — BASELINE config 1).  These pin
the ORACLE against drift; they are not outputs of the Julia reference (which cannot
run here: parity unpinned)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
f = np.asfortranarray(z["cameraman_128_5/data"] / 255.0)
t = np.asfortranarray(z["cameraman_128_5/true"] / 255.0)
u = orc.pdps(f, 0.1, maxiter=5000)
out = dict(maxiter=5000, u_sub=u[::8, ::8, 0], cost=orc.cost(u, t),
           grad_reg=orc.gradient_reg_scalar(0.1, u[:, :, 0], t[:, :, 0]),
           grad=orc.gradient_scalar(0.1, u[:, :, 0], t[:, :, 0]),
           grad_refined=orc.gradient_scalar(0.1, u[:, :, 0], t[:, :, 0], refine=4))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_pins.npz"), **out)
print({k: (v if np.ndim(v) == 0 else np.shape(v)) for k, v in out.items()})
