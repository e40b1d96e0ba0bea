"""CPU check of the band-LU gradient kernels (bpldenoising_b200/csrc/lu_band.cuh) where no GPU exists:
the device code is compiled with g++ against tests/emu/emu_cuda.h (one OS thread per CUDA thread, CTA /
warp barriers, shuffles) and run for one image, then compared with the oracle's literal row-scaled system
(`sumregs_gradient_reg`, patch variant, /root/reference/src/SumRegsLearningFunction.jl:195-262).  The
emulation checks index arithmetic and barrier placement, not performance; the GPU parity test proper is
tests/test_gpu_sumregs.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc
from oracle import sumregs as sr

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")


def _build(threads):
    out = os.path.join(EMU, "_build", f"libemu_lu_{threads}.so")
    srcs = [os.path.join(EMU, "emu_lu.cpp"), os.path.join(EMU, "emu_cuda.h"),
            os.path.join(HERE, "..", "bpldenoising_b200", "csrc", "lu_band.cuh"),
            os.path.join(HERE, "..", "bpldenoising_b200", "csrc", "sumregs_stencils.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", f"-DLU_THREADS={threads}", "-DBPLTV_EMU",
                        "-o", out, srcs[0]], check=True)
    lib = C.CDLL(out)
    lib.emu_lu_gradient.restype = C.c_int
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _case(n, seed):
    rng = np.random.default_rng(seed)
    t = np.round(rng.random((n, n)) * 255) / 255
    u = np.asfortranarray(t + 0.05 * rng.standard_normal((n, n)))
    u[:3, :3] = u[0, 0]                      # an exactly flat block: |∇_k u| = 0, the γ·I tensors
    return np.asfortranarray(t), u


def _run(lib, n, u, t, maps, alpha3, gamma, grid, vec_in_smem, want_band=False, nops=3, csize=1, use_pin=1,
         expect_pivot_failure=False):
    N = n * n
    bw = ((n if nops == 1 else 2 * n) + 1) & ~1          # lu_band_halfwidth (gradient_lu.cuh): even
    LD = 2 * (bw + 16) + 1
    out = np.zeros(nops * grid[0] * grid[1])
    band = np.zeros((N * LD + 3) & ~3) if want_band else None
    rr, pf, ld = C.c_double(), C.c_int(), C.c_int()
    am = None if maps is None else np.concatenate([m.flatten(order="F") for m in maps])
    a3 = None if alpha3 is None else np.asarray(alpha3, dtype=np.float64)
    rc = lib.emu_lu_gradient(use_pin, csize, nops, n, _ptr(u.flatten(order="F")), _ptr(t.flatten(order="F")), _ptr(am), _ptr(a3),
                             C.c_double(gamma), grid[0], grid[1], 3, int(vec_in_smem), _ptr(out), C.byref(rr),
                             C.byref(pf), _ptr(band), C.byref(ld))
    assert rc == 0 and ld.value == LD and (pf.value == 0 or expect_pivot_failure)
    got = out.reshape(nops, grid[1], grid[0]).transpose(2, 1, 0)      # [operator][patch] → (pi, pj, operator)
    return got, rr.value, (None if band is None else band[:N * LD].reshape(N, LD)), bw + 16


@pytest.mark.parametrize("n,threads,vec_in_smem", [(8, 256, 1), (12, 256, 0), (16, 512, 1)])
def test_band_lu_patch_reg_gradient_on_the_thread_emulation(n, threads, vec_in_smem):
    lib = _build(threads)
    t, u = _case(n, 100 + n)
    xp = np.stack([np.array([[0.03, 0.05], [0.02, 0.04]]) * s for s in (1.0, 0.7, 1.3)], axis=2)
    maps = [np.asfortranarray(orc.patch_upsample(xp[:, :, k], n, n)) for k in range(3)]
    gamma = 1e8                                                   # :200
    got, relres, band, bwx = _run(lib, n, u, t, maps, None, gamma, (2, 2), vec_in_smem, want_band=True)
    # the assembled band = the oracle's literal matrix (:246), entry by entry
    N = n * n
    uf = u.flatten(order="F")
    A = sp.identity(N, format="csr")
    for k, kind in enumerate(sr.KINDS):
        G = sr.op_matrix(kind, n)
        BmC, _ = sr._sets_reg(G, uf, gamma)
        A = A + sp.diags(maps[k].flatten(order="F")) @ (G.T @ BmC @ G)
    A = A.toarray()
    Ab = np.zeros((N, N))
    for i in range(N):
        lo, hi = max(0, i - bwx), min(N, i + bwx + 1)
        Ab[i, lo:hi] = band[i, lo - i + bwx:hi - i + bwx]
    assert np.abs(Ab - A).max() <= 1e-15 * np.abs(A).max()
    # the gradient = the refined literal solve; tolerance 1e-9 (the system's entries span 1 … αγ ≈ 1e7, the
    # same bar as the other regularised variants)
    lit = sr.sumregs_gradient_reg(maps, u, t, grid_shape=(2, 2), gamma=gamma, refine=3)
    assert np.all(np.abs(got - lit) <= 1e-9 * np.abs(lit).max()), (got, lit)
    assert relres <= 1e-8


def test_band_lu_scalar_parameter_on_the_thread_emulation():
    """The same kernels with a scalar 3-vector (alpha_maps == NULL): scalar sumregs_gradient_reg (:112-167, γ = 1e3)."""
    lib = _build(256)
    n = 10
    t, u = _case(n, 7)
    x = np.array([0.05, 0.04, 0.06])
    got, relres, _, _ = _run(lib, n, u, t, None, x, 1e3, (1, 1), 1)
    lit = sr.sumregs_gradient_reg(x, u, t, refine=3)
    assert np.all(np.abs(got.ravel() - lit) <= 1e-12 * np.abs(lit).max()), (got, lit)
    assert relres <= 1e-13


def test_band_lu_tv_gradient_reg_on_the_thread_emulation():
    """nops = 1: gradient_reg of the TV learning function, scalar (/root/reference/src/TVLearningFunctionVec.jl:137-161)
    and patch (:192-215, row-scaled), forward differences only, half-bandwidth n, γ = 1e8."""
    lib = _build(256)
    n = 12
    t, u = _case(n, 11)
    got, relres, _, _ = _run(lib, n, u, t, None, np.array([0.1, 0.0, 0.0]), 1e8, (1, 1), 1, nops=1)
    lit = orc.gradient_reg_scalar(0.1, u, t, refine=3)
    assert abs(got.ravel()[0] - lit) <= 1e-9 * abs(lit), (got, lit)
    x = np.array([[0.05, 0.1], [0.08, 0.02]])
    amap = np.asfortranarray(orc.patch_upsample(x, n, n))
    gotp, relres, _, _ = _run(lib, n, u, t, [amap], None, 1e8, (2, 2), 0, nops=1)
    litp = orc.gradient_reg_patch(amap, (2, 2), u, t, refine=3)
    assert np.all(np.abs(gotp[:, :, 0] - litp) <= 1e-9 * np.abs(litp).max()), (gotp, litp)


def test_band_lu_cluster_factorisation_is_invisible_on_the_thread_emulation():
    """lu_factor_kernel<CL = true>: the trailing tiles dealt over a cluster of CTAs, two cluster barriers per block
    step, one writer — the same bits as the single-CTA kernel (n = 20: 2×2 warp tiles of 32×32 per step)."""
    lib = _build(64)
    n = 20
    t, u = _case(n, 3)
    xp = np.stack([np.array([[0.03, 0.05], [0.02, 0.04]]) * s for s in (1.0, 0.7, 1.3)], axis=2)
    maps = [np.asfortranarray(orc.patch_upsample(xp[:, :, k], n, n)) for k in range(3)]
    one, rr1, _, _ = _run(lib, n, u, t, maps, None, 1e8, (2, 2), 1)
    for cs in (2, 3):
        many, rr, _, _ = _run(lib, n, u, t, maps, None, 1e8, (2, 2), 1, csize=cs)
        assert np.array_equal(one, many) and rr == rr1, cs
    # the cp.async staging of the panel inputs (needs 2·bw = 80 ≤ threads: on with 128 threads, off with 64 above)
    # and its absence give the same bits; a different CTA size only reorders the block sums of the functional
    lib128 = _build(128)
    a, _, _, _ = _run(lib128, n, u, t, maps, None, 1e8, (2, 2), 1, use_pin=1)
    b, _, _, _ = _run(lib128, n, u, t, maps, None, 1e8, (2, 2), 1, use_pin=0)
    assert np.array_equal(a, b) and np.allclose(a, one, rtol=1e-12, atol=0)


@pytest.mark.parametrize("threads,n,nops,cs,pin", [(128, 9, 3, 2, 1), (128, 13, 1, 3, 1), (64, 11, 3, 1, 0), (256, 17, 1, 2, 1)])
def test_band_lu_odd_sizes_clusters_and_staging_on_the_thread_emulation(threads, n, nops, cs, pin):
    """Odd image sizes (the forward-only band has half-width n: rounded up to even for the 16-byte accesses, which
    the emulation checks), short last blocks, clusters and the cp.async staging together."""
    lib = _build(threads)
    t, u = _case(n, n)
    if nops == 3:
        xp = np.stack([np.array([[0.03, 0.05], [0.02, 0.04]]) * s for s in (1.0, 0.7, 1.3)], axis=2)
        maps = [np.asfortranarray(orc.patch_upsample(xp[:, :, k], n, n)) for k in range(3)]
        got, _, _, _ = _run(lib, n, u, t, maps, None, 1e8, (2, 2), 1, nops=3, csize=cs, use_pin=pin)
        lit = sr.sumregs_gradient_reg(maps, u, t, grid_shape=(2, 2), gamma=1e8, refine=3)
    else:
        x = np.array([[0.05, 0.1], [0.08, 0.02]])
        amap = np.asfortranarray(orc.patch_upsample(x, n, n))
        got, _, _, _ = _run(lib, n, u, t, [amap], None, 1e8, (2, 2), 1, nops=1, csize=cs, use_pin=pin)
        got, lit = got[:, :, 0], orc.gradient_reg_patch(amap, (2, 2), u, t, refine=3)
    assert np.all(np.abs(got - lit) <= 2e-9 * np.abs(lit).max()), (got, lit)


def test_band_lu_poisons_the_output_when_a_pivot_vanishes():
    """No silent wrong answer: on a flat image with γ = 2²⁰ and the (unphysical, API-rejected) parameter
    α = −2⁻²¹ the first pivot 1 + 2αγ is exactly zero — the kernels must flag it and return NaN, which the API turns
    into BPLTV_ERR_NUMERIC."""
    lib = _build(64)
    n = 8
    u = np.asfortranarray(np.full((n, n), 0.5))
    t = np.asfortranarray(np.full((n, n), 0.25))
    out = np.zeros(1)
    rr, pf, ld = C.c_double(), C.c_int(), C.c_int()
    a3 = np.array([-2.0 ** -21, 0.0, 0.0])
    rc = lib.emu_lu_gradient(1, 1, 1, n, _ptr(u.flatten(order="F")), _ptr(t.flatten(order="F")), None, _ptr(a3),
                             C.c_double(2.0 ** 20), 1, 1, 3, 1, _ptr(out), C.byref(rr), C.byref(pf), None, C.byref(ld))
    assert rc == 0 and pf.value == 1 and np.isnan(out[0])
