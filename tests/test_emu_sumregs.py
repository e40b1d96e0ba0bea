"""CPU check of the cluster-resident sum-of-regularisers solve (sumregs_resident_kernel,
bpldenoising_b200/csrc/pdps_sumregs.cuh) on the thread emulation of tests/emu/: CTAs of a cluster run
concurrently on OS threads, `map_shared_rank` resolves into the other CTA's emulated shared memory, the
strict-arithmetic intrinsics are single IEEE operations (g++ -ffp-contract=off).  The result must be
BIT-IDENTICAL to the oracle (`oracle/sumregs.py`, /root/reference/src/SumRegsLearningFunction.jl:38-85) for every
cluster size — the halo pushes, the ragged last rank and the barrier placement are what this exercises; the GPU
parity test proper is tests/test_gpu_sumregs.py."""
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
    out = os.path.join(EMU, "_build", "libemu_sumregs.so")
    srcs = [os.path.join(EMU, "emu_sumregs.cpp"), os.path.join(EMU, "emu_cuda.h"),
            os.path.join(CSRC, "pdps_sumregs.cuh"), os.path.join(CSRC, "common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-DBPLTV_EMU",
                        "-o", out, srcs[0]], check=True)
    L = C.CDLL(out)
    L.emu_sumregs_resident.restype = C.c_int
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _run(L, f, x, maps, cs, prec=64, strict=1, maxiter=25, init_mode=0, threads=64):
    M, N, O = f.shape
    u = np.zeros((M, N, O), order="F")
    am = None if maps is None else np.concatenate([np.asarray(m, dtype=np.float64).flatten(order="F") for m in maps])
    a3 = None if x is None else np.asarray(x, dtype=np.float64)
    rc = L.emu_sumregs_resident(prec, M, N, O, cs, threads, maxiter, strict, init_mode, _ptr(np.asfortranarray(f)),
                                _ptr(a3), _ptr(am), _ptr(u))
    assert rc == 0, rc
    return u


@pytest.mark.parametrize("shape", [(10, 11, 2), (7, 9, 1), (16, 6, 1)])
def test_resident_sumregs_solve_is_bit_identical_for_every_cluster_size(lib, shape):
    M, N, O = shape
    rng = np.random.default_rng(M * 31 + N)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    x = np.array([0.03, 0.012, 0.05])
    ref = sr.sumregs_pdps(f, list(x), maxiter=25)
    xp = rng.uniform(0.005, 0.08, (2, 3, 3))
    maps = [orc.patch_upsample(xp[:, :, k], M, N) for k in range(3)]
    refp = sr.sumregs_pdps(f, maps, maxiter=25)
    ref32 = sr.sumregs_pdps(f, list(x), maxiter=25, dtype=np.float32)
    for cs in (1, 2, 3, 4):
        if (cs - 1) * -(-N // cs) >= N:
            continue                               # a rank would own no column
        assert np.array_equal(_run(lib, f, x, None, cs), ref), cs
        assert np.array_equal(_run(lib, f, None, maps, cs), refp), cs
        assert np.array_equal(_run(lib, f, x, None, cs, prec=32).astype(np.float32), ref32), cs
    # x⁰ = f (S3) and the fast arithmetic (FMA contraction, rsqrt): within the stated tolerance
    refi = sr.sumregs_pdps(f, list(x), maxiter=25, init_mode=1)
    uf = _run(lib, f, x, None, 2, strict=0, init_mode=1)
    assert np.linalg.norm(uf - refi) <= 1e-10 * np.linalg.norm(refi)
