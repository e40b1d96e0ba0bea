"""60-digit solution (mpmath) of the multiplier-space adjoint system of `gradient`
(/root/reference/src/TVLearningFunctionVec.jl:98-135) on the 16×16 crop of tests/golden/nd_hard_crop.npz, and what a
floating-point Cholesky with the bounded-multiplier pivot rule of bpldenoising_b200/csrc/nd_solver.cuh achieves on it
(p to ~1e-8, the functional to ~1e-6, refinement converging slowly: the near-null space of barely sloped regions).
Writes the value stored as g16_mp60 in the fixture.  ~1 minute."""
import os, sys
import numpy as np, scipy.sparse as sp, scipy.linalg as sla
import mpmath as mp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "nd_hard_crop.npz"))
u, t = np.asfortranarray(z["u"]), np.asfortranarray(z["t"])
s = orc.dual_setup("nonreg", 0.1, u, t)
iso = s["iso"]
off = np.concatenate([[0], np.cumsum(np.where(iso, 2, 1))])
B1 = sp.diags(s["ea"]) @ s["G1"] + sp.diags(s["eb"]) @ s["G2"]
Bfull = sp.vstack([B1, s["G2"][iso]]).tocsr()
perm = np.argsort(np.concatenate([off[:-1], off[:-1][iso] + 1]))
B = Bfull[perm].toarray()
E = np.concatenate([s["E"], s["E"][iso]])[perm]
A = np.diag(E) + B @ B.T
b = B @ s["rc"]


def functional(p):
    return -float(np.sum((s["G1"] @ p) * s["w1"] + (s["G2"] @ p) * s["w2"]))


mp.mp.dps = 60
Bm = mp.matrix(B.tolist())
rm = mp.matrix([mp.mpf(float(x)) for x in s["rc"]])
zm = mp.lu_solve(mp.diag([mp.mpf(float(e)) for e in E]) + Bm * Bm.T, Bm * rm)
pref = np.array([float(x) for x in (rm - Bm.T * zm)])
gref = functional(pref)
print("60-digit functional: %.18g   (fixture: %.18g)" % (gref, float(z["g16_mp60"])))
print("eigenvalues of A below 1e-13: %d of %d" % ((np.linalg.eigvalsh(A) < 1e-13).sum(), A.shape[0]))

# right-looking Cholesky with the pivot rule d ← max(d, 1e-13, a²/16), a = largest entry of the column
Lw = A.copy()
L = np.zeros_like(A)
raised = 0
for j in range(A.shape[0]):
    d = Lw[j, j]
    amax = np.abs(Lw[j + 1:, j]).max() if j + 1 < A.shape[0] else 0.0
    fl = max(1e-13, amax * amax / 16)
    if not d >= fl:
        d = fl
        raised += 1
    L[j, j] = np.sqrt(d)
    L[j + 1:, j] = Lw[j + 1:, j] / L[j, j]
    Lw[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
zeta = sla.solve_triangular(L.T, sla.solve_triangular(L, b, lower=True), lower=False)
for it in range(4):
    p = s["rc"] - B.T @ zeta
    res = B @ p - E * zeta
    print("fp64, %d pivots raised, refinement %d: functional error %.1e, p error %.1e, residual %.1e" %
          (raised, it, abs(functional(p) - gref) / abs(gref), np.linalg.norm(p - pref) / np.linalg.norm(pref),
           np.linalg.norm(res) / np.linalg.norm(b)))
    zeta = zeta + sla.solve_triangular(L.T, sla.solve_triangular(L, res, lower=True), lower=False)
