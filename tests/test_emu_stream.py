"""CPU check of the one-thread-per-pixel kernels on the thread emulation of tests/emu/: kernel G
(pdps_generic_kernel), the streaming pair of the sum-of-regularisers solve and the deterministic two-stage cost
reduction, launched as bpltv_api.cu launches them.  Solves BIT-IDENTICAL to the oracle (fp64, fp32, λ-maps, ragged
shapes incl. single rows / columns); cost within 1e-14 (its summation order differs from numpy's)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import sumregs as sr

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
CSRC = os.path.join(HERE, "..", "bpldenoising_b200", "csrc")


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(EMU, "_build", "libemu_stream.so")
    srcs = [os.path.join(EMU, "emu_stream.cpp"), os.path.join(EMU, "emu_cuda.h"), os.path.join(CSRC, "pdps_generic.cuh"),
            os.path.join(CSRC, "pdps_sumregs.cuh"), os.path.join(CSRC, "common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-DBPLTV_EMU",
                        "-o", out, srcs[0]], check=True)
    L = C.CDLL(out)
    L.emu_pdps_generic.restype = C.c_int
    L.emu_sumregs_stream.restype = C.c_int
    L.emu_cost.restype = C.c_double
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("shape", [(9, 7, 2), (1, 6, 1), (5, 1, 2), (40, 3, 1)])
def test_generic_kernel_and_sumregs_streaming_pair_are_bit_identical(lib, shape):
    M, N, O = shape
    rng = np.random.default_rng(M * 7 + N)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    its = 8
    amap = np.asfortranarray(rng.uniform(0.01, 0.1, (M, N)))
    for prec, dt in ((64, np.float64), (32, np.float32)):
        for alpha in (0.08, amap):
            u = np.zeros(shape, order="F")
            am = None if np.ndim(alpha) == 0 else alpha.flatten(order="F")
            rc = lib.emu_pdps_generic(prec, M, N, O, its, 1, _ptr(f), C.c_double(0.08), _ptr(am), _ptr(u))
            assert rc == 0 and np.array_equal(u.astype(dt), orc.pdps(f, alpha, maxiter=its, dtype=dt)), (prec, np.ndim(alpha))
        x = np.array([0.03, 0.012, 0.05])
        maps = [np.asfortranarray(rng.uniform(0.005, 0.08, (M, N))) for _ in range(3)]
        for al3, mp in ((x, None), (None, maps)):
            u = np.zeros(shape, order="F")
            am = None if mp is None else np.concatenate([m.flatten(order="F") for m in mp])
            rc = lib.emu_sumregs_stream(prec, M, N, O, its, 1, _ptr(f), _ptr(al3), _ptr(am), _ptr(u))
            ref = sr.sumregs_pdps(f, list(x) if mp is None else mp, maxiter=its, dtype=dt)
            assert rc == 0 and np.array_equal(u.astype(dt), ref), (prec, mp is None)


def test_cost_reduction_on_the_thread_emulation(lib):
    rng = np.random.default_rng(3)
    for n in (1, 255, 5000, 70001):
        u, ub = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
        got = lib.emu_cost(_ptr(u), _ptr(ub), C.c_longlong(n))
        ref = 0.5 * float(np.sum((u - ub) ** 2))
        assert abs(got - ref) <= 1e-14 * ref + 1e-300, n


def test_patch_upsample_and_sweep_errors_on_the_thread_emulation(lib):
    """PatchOp up-sampling (S7: block-constant, also for grids that do not divide the image) and the per-image squared
    errors of a λ-sweep (virtual image v = set·O + image reads truth image v % O)."""
    rng = np.random.default_rng(11)
    for (M, N, lm, ln) in ((12, 10, 2, 2), (13, 7, 3, 2), (5, 5, 5, 5), (9, 4, 1, 3)):
        x = np.asfortranarray(rng.uniform(0.01, 0.1, (lm, ln)))
        out = np.zeros((M, N), order="F")
        lib.emu_patch_upsample(_ptr(x), lm, ln, _ptr(out), M, N)
        assert np.array_equal(out, orc.patch_upsample(x, M, N)), (M, N, lm, ln)
    plane, O, L = 77, 3, 4
    ub = np.asfortranarray(rng.uniform(0, 1, (plane, O)))
    u = np.asfortranarray(rng.uniform(0, 1, (plane, O * L)))
    out = np.zeros(O * L)
    lib.emu_sqerr_images(_ptr(u), _ptr(ub), plane, O, O * L, _ptr(out))
    ref = np.array([np.sum((u[:, v] - ub[:, v % O]) ** 2) for v in range(O * L)])
    assert np.allclose(out, ref, rtol=1e-14, atol=0)
