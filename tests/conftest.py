import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cuda.cuInit(0) != 0:
            return False
        return cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def datasets():
    """The reference's datasets (tests/golden/datasets.npz, packed by
    tools/make_dataset_fixtures.py) as Float64 stacks k/255, like load_dataset
    (/root/reference/src/Datasets.jl:54-65)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
    out = {}
    for key in z.files:
        if key.endswith("/true"):
            name = key[:-5]
            t = z[name + "/true"].astype(np.float64) / z[name + "/true_div"].astype(np.float64)
            d = z[name + "/data"].astype(np.float64) / z[name + "/data_div"].astype(np.float64)
            out[name] = (np.asfortranarray(t), np.asfortranarray(d))
    return out


@pytest.fixture(scope="session")
def bp():
    import bpldenoising_b200 as b
    return b


@pytest.fixture(autouse=True)
def _fresh_env_switches():
    """The library snapshots the BPLTV_* switches; tests that flip them call bp.reload_env() themselves, and the snapshot
    is refreshed again after every test so that a restored environment is what the next test sees."""
    yield
    if "bpldenoising_b200" in sys.modules:
        sys.modules["bpldenoising_b200"].reload_env()


@pytest.fixture(scope="module")
def ctx(bp):
    c = bp.Context([0], 64)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx32(bp):
    c = bp.Context([0], 32)
    yield c
    c.close()


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (nb if nb > 0 else 1.0))
