"""Host-side mirror of the reference interface: argument handling that needs no GPU."""
import numpy as np
import pytest


def test_lambda_and_stack_normalisation(bp):
    from bpldenoising_b200 import learning as L
    lam, scalar = L._lam(0.1)
    assert scalar and lam.shape == (1, 1)
    lam, scalar = L._lam(np.ones((2, 3)))
    assert not scalar and lam.flags.f_contiguous
    with pytest.raises(ValueError):
        L._lam(np.ones(3))
    s = L._stack(np.zeros((4, 5)))
    assert s.shape == (4, 5, 1) and s.flags.f_contiguous
    with pytest.raises(ValueError):
        L._stack(np.zeros(4))
    assert L.L2CostFunction(np.ones((2, 2)), np.zeros((2, 2))) == 2.0


def test_synthetic_dataset_is_frozen(bp):
    t, f = bp.synthetic_dataset(32, 24, 3, seed=20240601)
    t2, f2 = bp.synthetic_dataset(32, 24, 3, seed=20240601)
    assert np.array_equal(t, t2) and np.array_equal(f, f2)
    assert t.shape == (32, 24, 3) and f.flags.f_contiguous
    assert np.allclose(f * 255, np.round(f * 255)) and f.min() >= 0 and f.max() <= 1
    assert 0.05 < np.std(f - t) < 0.15
