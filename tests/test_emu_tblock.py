"""CPU check of kernel C (pdps_tblock_kernel, bpldenoising_b200/csrc/pdps_tblock.cuh — the headline kernel of BASELINE
config 4: T iterations of the reference's PDPS recursion (external op_denoise_pdps; docs/SEMANTICS.md S1-S9) per pass over
the stack, software-pipelined along the column march, stage 0 fed through the TMA ring) on the thread emulation of
tests/emu/.  BIT-IDENTICAL to the oracle for every depth, every range cut (ranges shorter than the halo, cuts inside images,
one column per CTA), the ring kernels (16-byte rows per thread) and the direct-load kernels, λ-map, fp32; the GPU parity
tests proper are in tests/test_gpu_pdps.py."""
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
    out = os.path.join(EMU, "_build", "libemu_tblock.so")
    srcs = [os.path.join(EMU, "emu_tblock.cpp"), os.path.join(EMU, "emu_cuda.h"),
            os.path.join(CSRC, "pdps_tblock.cuh"), os.path.join(CSRC, "common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-DBPLTV_EMU",
                        "-o", out, srcs[0]], check=True)
    L = C.CDLL(out)
    L.emu_pdps_tblock.restype = C.c_int
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _run(L, f, alpha, T, grid, vec, prec=64, strict=1, maxiter=12):
    M, N, O = f.shape
    u = np.zeros((M, N, O), order="F")
    amap = None if np.ndim(alpha) == 0 else np.asarray(alpha, dtype=np.float64).flatten(order="F")
    rc = L.emu_pdps_tblock(prec, vec, T, M, N, O, grid, maxiter, strict, _ptr(np.asfortranarray(f)),
                           C.c_double(float(alpha) if amap is None else 0.0), _ptr(amap), _ptr(u))
    assert rc == 0, rc
    return u


@pytest.mark.parametrize("T", [2, 3, 4])
@pytest.mark.parametrize("shape,vec,prec", [((16, 11, 3), 2, 64), ((24, 7, 2), 1, 64), ((40, 9, 2), 2, 64), ((16, 10, 2), 4, 32), ((12, 9, 2), 2, 32)])
def test_kernel_c_is_bit_identical_for_every_depth_and_range_cut(lib, T, shape, vec, prec):
    M, N, O = shape
    rng = np.random.default_rng(M * 7 + N + T)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    its = 3 * T
    dt = np.float32 if prec == 32 else np.float64
    ref = orc.pdps(f, 0.08, maxiter=its, dtype=dt).astype(np.float64)
    amap = orc.patch_upsample(np.array([[0.05, 0.1], [0.08, 0.02]]), M, N)
    refm = orc.pdps(f, amap, maxiter=its, dtype=dt).astype(np.float64)
    for grid in (1, 2, 3, N * O // 2, N * O):              # one range … one column per CTA; cuts inside images
        assert np.array_equal(_run(lib, f, 0.08, T, grid, vec, prec=prec, maxiter=its), ref), grid
    assert np.array_equal(_run(lib, f, amap, T, 3, vec, prec=prec, maxiter=its), refm)


def test_kernel_c_more_than_one_warp_per_column_and_fast_arithmetic(lib):
    """Columns taller than a warp's rows: the warp-boundary slots (x̄ of the first row → the warp above, the finished y1 of
    the last row → the warp below) carry the row neighbours; fast arithmetic stays within the stated tolerance."""
    rng = np.random.default_rng(5)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, (96, 6, 2)) * 255) / 255)       # 48 threads per column at vec 2: two warps
    ref = orc.pdps(f, 0.1, maxiter=8)
    for T in (2, 4):
        assert np.array_equal(_run(lib, f, 0.1, T, 4, 2, maxiter=8), ref), T
    uf = _run(lib, f, 0.1, 4, 2, 2, strict=0, maxiter=8)
    assert np.linalg.norm(uf - ref) <= 1e-10 * np.linalg.norm(ref)
