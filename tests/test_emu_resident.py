"""CPU check of kernel B (pdps_resident_kernel, bpldenoising_b200/csrc/pdps_resident.cuh: the TV solve of BASELINE
configs 1-3 — one launch, one image per thread-block cluster, duals and x̄ in shared memory, boundary columns
pushed through distributed shared memory, row neighbours by warp shuffle) on the thread emulation of tests/emu/.
BIT-IDENTICAL to the oracle (`oracle.pdps`, /root/reference/src/TVLearningFunctionVec.jl:45-70 + docs/SEMANTICS.md)
for every cluster size, with a ragged last rank, a λ-map and in fp32; the GPU parity tests proper are in
tests/test_gpu_pdps.py."""
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
    out = os.path.join(EMU, "_build", "libemu_resident.so")
    srcs = [os.path.join(EMU, "emu_resident.cpp"), os.path.join(EMU, "emu_cuda.h"),
            os.path.join(CSRC, "pdps_resident.cuh"), os.path.join(CSRC, "common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-DBPLTV_EMU",
                        "-o", out, srcs[0]], check=True)
    L = C.CDLL(out)
    L.emu_pdps_resident.restype = C.c_int
    L.emu_pdps_resident_tb.restype = C.c_int
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _run(L, f, alpha, cs, prec=64, strict=1, maxiter=30, threads=64):
    M, N, O = f.shape
    u = np.zeros((M, N, O), order="F")
    amap = None if np.ndim(alpha) == 0 else np.asarray(alpha, dtype=np.float64).flatten(order="F")
    rc = L.emu_pdps_resident(prec, M, N, O, cs, threads, maxiter, strict, _ptr(np.asfortranarray(f)),
                             C.c_double(float(alpha) if amap is None else 0.0), _ptr(amap), _ptr(u))
    assert rc == 0, rc
    return u


@pytest.mark.parametrize("async_halo", [1, 0])
@pytest.mark.parametrize("shape", [(16, 13, 2), (8, 21, 1), (32, 6, 1)])
def test_kernel_b_is_bit_identical_for_every_cluster_size(lib, shape, async_halo):
    """Both halo schemes of kernel B: async_halo = 1 — boundary columns sent with (emulated) st.async, counted on an
    mbarrier of the receiver, CTA barriers between the phases (the default on the GPU); 0 — plain DSMEM stores and two
    cluster barriers per iteration."""
    lib.emu_resident_set_async(async_halo)
    M, N, O = shape
    rng = np.random.default_rng(M * 17 + N)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    ref = orc.pdps(f, 0.08, maxiter=30)
    amap = orc.patch_upsample(np.array([[0.05, 0.1], [0.08, 0.02]]), M, N)
    refm = orc.pdps(f, amap, maxiter=30)
    ref32 = orc.pdps(f, 0.08, maxiter=30, dtype=np.float32)
    ran = 0
    for cs in (1, 2, 3, 4):
        nc = -(-N // cs)
        if (cs - 1) * nc >= N or -(-nc // (64 // (M // 2))) > 4:
            continue                               # a rank without a column / more than four column slots per thread
        assert np.array_equal(_run(lib, f, 0.08, cs), ref), cs
        assert np.array_equal(_run(lib, f, amap, cs), refm), cs
        assert np.array_equal(_run(lib, f, 0.08, cs, prec=32).astype(np.float32), ref32), cs
        ran += 1
    assert ran >= 2
    uf = _run(lib, f, 0.08, 2, strict=0)           # fast arithmetic: within the stated tolerance
    assert np.linalg.norm(uf - ref) <= 1e-10 * np.linalg.norm(ref)


def _run_tb(L, f, alpha, cs, kc, prec=64, strict=1, maxiter=30):
    M, N, O = f.shape
    u = np.zeros((M, N, O), order="F")
    amap = None if np.ndim(alpha) == 0 else np.asarray(alpha, dtype=np.float64).flatten(order="F")
    rc = L.emu_pdps_resident_tb(prec, M, N, O, cs, kc, maxiter, strict, _ptr(np.asfortranarray(f)),
                                C.c_double(float(alpha) if amap is None else 0.0), _ptr(amap), _ptr(u))
    assert rc == 0, rc
    return u


@pytest.mark.parametrize("shape", [(16, 13, 2), (8, 21, 1), (32, 9, 1), (64, 12, 1)])
def test_temporally_blocked_kernel_b_is_bit_identical(lib, shape):
    """pdps_resident_tb_kernel: two iterations per halo exchange, two redundant columns per side.  Bit-identical to the
    oracle for every cluster size (ragged last rank, a rank with fewer columns than the halo is wide), every slot count,
    odd and even iteration counts (the odd tail runs half a super-step), a λ-map and fp32."""
    M, N, O = shape
    rng = np.random.default_rng(M * 31 + N)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    amap = orc.patch_upsample(np.array([[0.05, 0.1], [0.08, 0.02]]), M, N)
    ran = 0
    for maxiter in (1, 2, 7, 30):
        ref = orc.pdps(f, 0.08, maxiter=maxiter)
        refm = orc.pdps(f, amap, maxiter=maxiter)
        ref32 = orc.pdps(f, 0.08, maxiter=maxiter, dtype=np.float32)
        for cs in (1, 2, 3, 4):
            nc = -(-N // cs)
            if nc < 2 or (cs - 1) * nc >= N:
                continue
            for kc in (1, 2, 4):
                if (M // 2) * -(-(nc + 4) // kc) > 768:
                    continue
                assert np.array_equal(_run_tb(lib, f, 0.08, cs, kc, maxiter=maxiter), ref), (maxiter, cs, kc)
                if maxiter in (7, 30) and kc == 2:
                    assert np.array_equal(_run_tb(lib, f, amap, cs, kc, maxiter=maxiter), refm), (maxiter, cs, kc)
                    assert np.array_equal(_run_tb(lib, f, 0.08, cs, kc, prec=32, maxiter=maxiter).astype(np.float32), ref32), (maxiter, cs)
                ran += 1
    assert ran >= 8
    uf = _run_tb(lib, f, 0.08, 2, 2, strict=0)
    ref = orc.pdps(f, 0.08, maxiter=30)
    assert np.linalg.norm(uf - ref) <= 1e-10 * np.linalg.norm(ref)
