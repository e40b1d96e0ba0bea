"""Optional: golden vectors from the real Julia reference (tools/dump_reference_vectors.jl).
Skipped unless tests/golden/julia/ exists — Julia is not installed in the build container,
so the repository ships none (parity unpinned, docs/SEMANTICS.md)."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT, rel_l2

JDIR = os.path.join(ROOT, "tests", "golden", "julia")
pytestmark = pytest.mark.skipif(not os.path.isdir(JDIR), reason="no Julia golden vectors supplied")

CASES = {"scalar_nonreg": (0.1, 0.1), "scalar_reg": (0.1, 1e-7),
         "patch_nonreg": (1e-4 * np.ones((2, 2)), 1e-4), "patch_reg": (1e-4 * np.ones((2, 2)), 1e-7)}


def _load(path_u):
    meta = np.fromfile(path_u.replace(".u.f64", ".meta.f64"), dtype="<f8")
    M, N, O = (int(v) for v in meta[:3])
    u = np.fromfile(path_u, dtype="<f8").reshape((M, N, O), order="F")
    ng = int(meta[4])
    return u, float(meta[3]), meta[5:5 + ng]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(JDIR, "*.u.f64"))))
def test_oracle_against_julia(oracle, datasets, path):
    name, tag = os.path.basename(path)[:-6].rsplit("_", 2)[0], "_".join(os.path.basename(path)[:-6].rsplit("_", 2)[1:])
    x, Delta = CASES[tag]
    u, cost, grad = _load(path)
    ou, ocost, ograd = oracle.tv_op_learning_function(x, datasets[name], Delta, refine=3)
    assert rel_l2(ou, u) <= 1e-10
    assert abs(ocost - cost) <= 1e-10 * abs(cost)
    assert rel_l2(np.ravel(ograd), grad) <= (1e-9 if tag.endswith("_reg") else 1e-4)
