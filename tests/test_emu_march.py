"""CPU check of kernel A (pdps_march_kernel, bpldenoising_b200/csrc/pdps_march.cuh: the single-pass HBM-streaming
PDPS iteration — column ranges per CTA, the previous column carried in registers, row neighbours by warp shuffle
and a shared-memory slot per warp boundary, ping-pong state buffers) on the thread emulation of tests/emu/.
BIT-IDENTICAL to the oracle for every grid size (range cut), vector width, with ranges that cross image
boundaries, a λ-map and in fp32; the GPU parity tests proper are in tests/test_gpu_pdps.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
CSRC = os.path.join(HERE, "..", "bpldenoising_b200", "csrc")


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(EMU, "_build", "libemu_march.so")
    srcs = [os.path.join(EMU, "emu_march.cpp"), os.path.join(EMU, "emu_cuda.h"),
            os.path.join(CSRC, "pdps_march.cuh"), os.path.join(CSRC, "common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-DBPLTV_EMU",
                        "-o", out, srcs[0]], check=True)
    L = C.CDLL(out)
    L.emu_pdps_march.restype = C.c_int
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _run(L, f, alpha, grid, vec, threads, prec=64, strict=1, maxiter=12):
    M, N, O = f.shape
    u = np.zeros((M, N, O), order="F")
    amap = None if np.ndim(alpha) == 0 else np.asarray(alpha, dtype=np.float64).flatten(order="F")
    rc = L.emu_pdps_march(prec, vec, M, N, O, grid, threads, maxiter, strict, _ptr(np.asfortranarray(f)),
                          C.c_double(float(alpha) if amap is None else 0.0), _ptr(amap), _ptr(u))
    assert rc == 0, rc
    return u


@pytest.mark.parametrize("shape,vec,threads", [((16, 9, 3), 2, 32), ((80, 5, 2), 2, 64), ((24, 7, 2), 1, 32), ((8, 11, 2), 4, 32)])
def test_kernel_a_is_bit_identical_for_every_range_cut(lib, shape, vec, threads):
    M, N, O = shape
    rng = np.random.default_rng(M * 13 + N)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    its = 12
    prec = 32 if vec == 4 else 64                       # four rows per thread is the fp32 layout
    dt = np.float32 if prec == 32 else np.float64
    ref = orc.pdps(f, 0.08, maxiter=its, dtype=dt)
    amap = orc.patch_upsample(np.array([[0.05, 0.1], [0.08, 0.02]]), M, N)
    refm = orc.pdps(f, amap, maxiter=its, dtype=dt)
    for grid in (1, 2, 5, N * O):                       # one range … one column per CTA; 2 and 5 cut inside images
        got = _run(lib, f, 0.08, grid, vec, threads, prec=prec, maxiter=its)
        assert np.array_equal(got.astype(dt), ref), grid
        gotm = _run(lib, f, amap, grid, vec, threads, prec=prec, maxiter=its)
        assert np.array_equal(gotm.astype(dt), refm), grid
    if prec == 64:
        uf = _run(lib, f, 0.08, 3, vec, threads, strict=0, maxiter=its)
        assert np.linalg.norm(uf - ref) <= 1e-10 * np.linalg.norm(ref)
