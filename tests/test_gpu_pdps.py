"""Parity of the CUDA lower-level solve with the CPU oracle, through the C ABI.

Tolerances (BASELINE.json north_star): relative L2 ≤ 1e-10 in fp64, ≤ 1e-5 in fp32.
The strict arithmetic mode is held to a stronger bar: bit-identical iterates.
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL64, TOL32 = 1e-10, 1e-5


def _opts(bp, **kw):
    return bp.pdps_opts(**kw)


@pytest.mark.parametrize("kernel", ["generic", "march", "resident", "auto"])
@pytest.mark.parametrize("lam_kind", ["scalar", "map"])
def test_strict_mode_is_bit_identical_to_the_oracle(bp, ctx, oracle, datasets, kernel, lam_kind):
    f = datasets["faces_train_128_10"][1][:, :, :3].copy(order="F")
    kid = dict(generic=bp.KERNEL_GENERIC, march=bp.KERNEL_MARCH, resident=bp.KERNEL_RESIDENT, auto=bp.KERNEL_AUTO)[kernel]
    if lam_kind == "scalar":
        x, alpha = 0.1, 0.1
    else:
        x = np.array([[0.02, 0.1], [0.2, 0.05]])
        alpha = oracle.patch_upsample(x, 128, 128)
    u = ctx.denoise(f, x, _opts(bp, maxiter=300, kernel=kid, arith=bp.STRICT))
    ref = oracle.pdps(f, alpha, maxiter=300)
    assert np.array_equal(u, ref), f"max abs diff {np.abs(u - ref).max()}"


@pytest.mark.parametrize("shape", [(128, 128, 1), (64, 48, 2), (33, 17, 3), (1, 40, 1), (40, 1, 2), (2, 2, 1),
                                   (130, 70, 1), (256, 256, 2), (516, 40, 1)])
def test_ragged_shapes_all_kernels(bp, ctx, oracle, shape):
    M, N, O = shape
    rng = np.random.default_rng(M * 1000 + N)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    ref = oracle.pdps(f, 0.08, maxiter=60)
    for kid in (bp.KERNEL_GENERIC, bp.KERNEL_AUTO):
        u = ctx.denoise(f, 0.08, _opts(bp, maxiter=60, kernel=kid))
        assert np.array_equal(u, ref), (shape, kid)
    for vec in (1, 2, 4):
        import os
        if M % vec:
            continue
        os.environ["BPLTV_MARCH_VEC"] = str(vec)
        bp.reload_env()
        try:
            u = ctx.denoise(f, 0.08, _opts(bp, maxiter=60, kernel=bp.KERNEL_MARCH))
        finally:
            del os.environ["BPLTV_MARCH_VEC"]
            bp.reload_env()
        assert np.array_equal(u, ref), (shape, "march vec", vec)


@pytest.mark.parametrize("shape", [(128, 128, 1), (128, 128, 10), (64, 48, 3), (32, 20, 2), (128, 9, 1), (16, 128, 2),
                                   (256, 64, 1), (2, 2, 1), (128, 128, 40)])
def test_resident_kernel_shapes(bp, ctx, ctx32, oracle, shape):
    """Cluster-resident solve: every cluster size / slot count the planner can pick,
    ragged column splits, more clusters than fit at once (O=40)."""
    M, N, O = shape
    rng = np.random.default_rng(M * 131 + N)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    x = np.array([[0.03, 0.12], [0.07, 0.2]])
    for lam, alpha in ((0.08, 0.08), (x, oracle.patch_upsample(x, M, N))):
        ref = oracle.pdps(f, alpha, maxiter=80)
        u = ctx.denoise(f, lam, _opts(bp, maxiter=80, kernel=bp.KERNEL_RESIDENT))
        assert ctx.stats()["pdps_kernel_used"] == bp.KERNEL_RESIDENT and ctx.stats()["kernel_launches"] <= 2
        assert np.array_equal(u, ref), (shape, np.ndim(lam))
        uf = ctx.denoise(f, lam, _opts(bp, maxiter=80, kernel=bp.KERNEL_RESIDENT, arith=bp.FAST, init_mode=1))
        assert rel_l2(uf, oracle.pdps(f, alpha, maxiter=80, init_mode=1)) <= TOL64
    ref32 = oracle.pdps(f, 0.08, maxiter=80, dtype=np.float32)
    u32 = ctx32.denoise(f, 0.08, _opts(bp, maxiter=80, kernel=bp.KERNEL_RESIDENT))
    assert np.array_equal(u32.astype(np.float32), ref32)


def test_resident_kernel_refuses_what_it_cannot_hold(bp, ctx):
    f = np.zeros((512, 512, 1), order="F")
    with pytest.raises(bp.BpltvError):
        ctx.denoise(f, 0.1, _opts(bp, maxiter=2, kernel=bp.KERNEL_RESIDENT))
    with pytest.raises(bp.BpltvError):
        ctx.denoise(np.zeros((33, 8, 1), order="F"), 0.1, _opts(bp, maxiter=2, kernel=bp.KERNEL_RESIDENT))
    with pytest.raises(bp.BpltvError):
        ctx.denoise(np.zeros((32, 8, 1), order="F"), 0.1, _opts(bp, maxiter=2, kernel=bp.KERNEL_RESIDENT, rho=0.1))


def test_march_chunking_is_invisible(bp, ctx, oracle, datasets):
    import os
    f = datasets["cameraman_128_5"][1]
    ref = oracle.pdps(f, 0.1, maxiter=40)
    for chunk in (1, 3, 8, 127, 128):
        os.environ["BPLTV_MARCH_CHUNK"] = str(chunk)
        bp.reload_env()
        try:
            u = ctx.denoise(f, 0.1, _opts(bp, maxiter=40, kernel=bp.KERNEL_MARCH))
        finally:
            del os.environ["BPLTV_MARCH_CHUNK"]
            bp.reload_env()
        assert np.array_equal(u, ref), chunk


@pytest.mark.parametrize("kw", [dict(rho=0.3), dict(init_mode=1), dict(accel=0), dict(tau0=2.0, sigma0=0.3),
                                dict(opnorm=8.0), dict(maxiter=0), dict(maxiter=1), dict(maxiter=7)])
def test_solver_switches(bp, ctx, oracle, datasets, kw):
    f = datasets["circle_128_10"][1]
    okw = dict(maxiter=50)
    okw.update({k: (bool(v) if k == "accel" else v) for k, v in kw.items()})
    ref = oracle.pdps(f, 0.05, **okw)
    for kid in (bp.KERNEL_GENERIC, bp.KERNEL_MARCH):
        gkw = dict(maxiter=50, kernel=kid)
        gkw.update(kw)
        u = ctx.denoise(f, 0.05, _opts(bp, **gkw))
        assert np.array_equal(u, ref), (kw, kid)


def test_fast_mode_within_tolerance(bp, ctx, oracle, datasets):
    t, f = datasets["cameraman_128_5"]
    ref = oracle.pdps(f, 0.1, maxiter=5000)
    for kid in (bp.KERNEL_GENERIC, bp.KERNEL_MARCH, bp.KERNEL_RESIDENT):
        u = ctx.denoise(f, 0.1, _opts(bp, maxiter=5000, kernel=kid, arith=bp.FAST))
        assert rel_l2(u, ref) <= TOL64, kid


def test_fp32_mode(bp, ctx32, oracle, datasets):
    f = datasets["faces_train_128_10"][1]
    ref64 = oracle.pdps(f, 0.1, maxiter=2000)
    ref32 = oracle.pdps(f, 0.1, maxiter=2000, dtype=np.float32)
    for kid in (bp.KERNEL_GENERIC, bp.KERNEL_MARCH):
        u = ctx32.denoise(f, 0.1, _opts(bp, maxiter=2000, kernel=kid, arith=bp.STRICT))
        assert np.array_equal(u.astype(np.float32), ref32), kid          # bit-identical to the fp32 oracle
        assert rel_l2(u, ref64) <= TOL32                                     # and within 1e-5 of fp64
        uf = ctx32.denoise(f, 0.1, _opts(bp, maxiter=2000, kernel=kid, arith=bp.FAST))
        assert rel_l2(uf, ref64) <= TOL32


def test_full_reference_configs_denoise(bp, ctx, oracle, datasets):
    # BASELINE configs 1-3 at full size and iteration count
    for name, x in (("cameraman_128_5", 0.1), ("faces_train_128_10", 0.1),
                    ("circle_128_10", 1e-4 * np.ones((2, 2)))):
        t, f = datasets[name]
        alpha = x if np.ndim(x) == 0 else oracle.patch_upsample(x, 128, 128)
        ref = oracle.pdps(f, alpha, maxiter=5000)
        u = ctx.denoise(f, x)
        assert np.array_equal(u, ref), name
    # validation solve: TVDenoise = 10000 iterations (/root/reference/src/BPLDenoising.jl:51)
    t, f = datasets["faces_val_128_10"]
    u = bp.TVDenoise(f[:, :, :2], 0.07, ctx=ctx)
    assert np.array_equal(u, oracle.pdps(f[:, :, :2], 0.07, maxiter=10000))


def test_resident_dataset_and_empty_stack(bp, ctx, oracle, datasets):
    t, f = datasets["faces_train_128_10"]
    ctx.set_dataset((t, f))
    u = ctx.denoise(None, 0.1, _opts(bp, maxiter=100))
    assert np.array_equal(u, oracle.pdps(f, 0.1, maxiter=100))
    empty = np.zeros((16, 16, 0), order="F")
    assert ctx.denoise(empty, 0.1, _opts(bp, maxiter=10)).shape == (16, 16, 0)


def test_size_independent_properties_at_config4_size(bp, ctx):
    # 64 × 512×512 (BASELINE config 4): properties that need no oracle run
    truth, noisy = bp.synthetic_dataset(512, 512, 64, seed=20240601)
    # λ = 0 with x⁰ = f is the identity up to rounding (radius 0 ⇒ y ≡ 0 ⇒ x = (x+τf)/(1+τ))
    u = ctx.denoise(noisy, 0.0, _opts(bp, maxiter=20, init_mode=1))
    assert np.abs(u - noisy).max() <= 1e-14
    # the two independent kernels agree bit for bit
    a = ctx.denoise(noisy, 0.1, _opts(bp, maxiter=25, kernel=bp.KERNEL_MARCH))
    b = ctx.denoise(noisy, 0.1, _opts(bp, maxiter=25, kernel=bp.KERNEL_GENERIC))
    assert np.array_equal(a, b)
    # batch independence: image 37 alone gives the same answer as inside the batch
    c = ctx.denoise(noisy[:, :, 37], 0.1, _opts(bp, maxiter=25))
    assert np.array_equal(c[:, :, 0], a[:, :, 37])
    # a constant image is a fixed point; adding a constant shifts the solution by it
    const = np.full((512, 512, 1), 0.25, order="F")
    assert np.abs(ctx.denoise(const, 0.1, _opts(bp, maxiter=25, init_mode=1)) - const).max() <= 1e-14
    # denoising reduces total variation and keeps the range
    def tv(v):
        return np.abs(np.diff(v, axis=0)).sum() + np.abs(np.diff(v, axis=1)).sum()
    long = ctx.denoise(noisy[:, :, :2], 0.1, _opts(bp, maxiter=1000))
    assert tv(long) < 0.5 * tv(noisy[:, :, :2])
    assert long.min() >= noisy.min() - 1e-9 and long.max() <= noisy.max() + 1e-9


def test_error_behaviour(bp, ctx, datasets):
    t, f = datasets["cameraman_128_5"]
    with pytest.raises(bp.BpltvError) as ei:
        ctx.denoise(f, -0.1)
    assert ei.value.code == -1
    with pytest.raises(bp.BpltvError):
        ctx.denoise(f, np.nan)
    with pytest.raises(bp.BpltvError):
        ctx.denoise(f, 0.1, _opts(bp, tau0=50.0))  # τ₀σ₀ < 1 violated
    with pytest.raises(bp.BpltvError):
        ctx.denoise(np.zeros((4, 4, 1), order="F"), 0.1, _opts(bp, kernel=bp.KERNEL_TBLOCK, tblock=9, maxiter=1))
    with pytest.raises(bp.BpltvError):   # the pipelined march has no ρ path
        ctx.denoise(np.zeros((4, 4, 1), order="F"), 0.1, _opts(bp, kernel=bp.KERNEL_TBLOCK, rho=0.2, maxiter=1))
    fresh = bp.Context([0], 64)
    with pytest.raises(bp.BpltvError) as ei:
        fresh.learn_eval(0.1, 0.1)
    assert ei.value.code == -4
    with pytest.raises(bp.BpltvError):
        fresh.denoise(None, 0.1)
    fresh.close()


# ---- kernel C: temporally blocked (pipelined) march --------------------------------------
@pytest.mark.parametrize("depth", [2, 3, 4])
@pytest.mark.parametrize("lam_kind", ["scalar", "map"])
def test_tblock_is_bit_identical_to_the_oracle(bp, ctx, oracle, datasets, depth, lam_kind):
    f = datasets["faces_train_128_10"][1][:, :, :3].copy(order="F")
    if lam_kind == "scalar":
        x, alpha = 0.1, 0.1
    else:
        x = np.array([[0.02, 0.1], [0.2, 0.05]])
        alpha = oracle.patch_upsample(x, 128, 128)
    for maxiter in (301, 24):   # remainders 1 / 1 / 1 and 0 / 0 / 0 of the T-iteration passes
        u = ctx.denoise(f, x, _opts(bp, maxiter=maxiter, kernel=bp.KERNEL_TBLOCK, tblock=depth))
        assert ctx.stats()["pdps_kernel_used"] == bp.KERNEL_TBLOCK
        assert ctx.stats()["kernel_launches"] <= maxiter // depth + maxiter % depth + 2
        ref = oracle.pdps(f, alpha, maxiter=maxiter)
        assert np.array_equal(u, ref), (depth, maxiter, np.abs(u - ref).max())


@pytest.mark.parametrize("depth", [2, 3, 4])
def test_tblock_range_ends_are_invisible(bp, ctx, oracle, datasets, depth):
    """Column ranges shorter than, equal to and longer than the T-1 halo, ranges that
    straddle two images and ranges ending one column before the image edge."""
    import os
    f = np.asfortranarray(datasets["faces_train_128_10"][1][:64, :37, :3])
    ref = oracle.pdps(f, 0.1, maxiter=12 + depth - 1)
    for chunk in (1, 2, 3, 4, 5, 9, 36, 37, 38, 50, 111):
        os.environ["BPLTV_MARCH_CHUNK"] = str(chunk)
        bp.reload_env()
        try:
            u = ctx.denoise(f, 0.1, _opts(bp, maxiter=12 + depth - 1, kernel=bp.KERNEL_TBLOCK, tblock=depth))
        finally:
            del os.environ["BPLTV_MARCH_CHUNK"]
            bp.reload_env()
        assert np.array_equal(u, ref), (depth, chunk, np.abs(u - ref).max())


@pytest.mark.parametrize("shape", [(128, 128, 1), (64, 48, 2), (33, 17, 3), (1, 40, 1), (40, 1, 2), (2, 2, 1),
                                   (130, 70, 1), (256, 256, 2), (512, 40, 1), (6, 3, 5)])
def test_tblock_ragged_shapes_and_precisions(bp, ctx, ctx32, oracle, shape):
    import os
    M, N, O = shape
    rng = np.random.default_rng(M * 977 + N)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    ref = oracle.pdps(f, 0.08, maxiter=30)
    ref32 = oracle.pdps(f, 0.08, maxiter=30, dtype=np.float32)
    for depth in (2, 3, 4):
        for vec in (1, 2, 4):
            if M % vec:
                continue
            os.environ["BPLTV_MARCH_VEC"] = str(vec)
            bp.reload_env()
            try:
                if vec <= 2 and M // vec <= 256:
                    u = ctx.denoise(f, 0.08, _opts(bp, maxiter=30, kernel=bp.KERNEL_TBLOCK, tblock=depth))
                    assert np.array_equal(u, ref), (shape, depth, vec)
                if M // vec <= 256:
                    u32 = ctx32.denoise(f, 0.08, _opts(bp, maxiter=30, kernel=bp.KERNEL_TBLOCK, tblock=depth))
                    assert np.array_equal(u32.astype(np.float32), ref32), (shape, depth, vec, "fp32")
            finally:
                del os.environ["BPLTV_MARCH_VEC"]
                bp.reload_env()


def test_tblock_fast_mode_and_switches(bp, ctx, oracle, datasets):
    t, f = datasets["cameraman_128_5"]
    ref = oracle.pdps(f, 0.1, maxiter=2000)
    for depth in (2, 4):
        u = ctx.denoise(f, 0.1, _opts(bp, maxiter=2000, kernel=bp.KERNEL_TBLOCK, tblock=depth, arith=bp.FAST))
        assert rel_l2(u, ref) <= TOL64, depth
    for kw in (dict(init_mode=1), dict(accel=0), dict(tau0=2.0, sigma0=0.3), dict(maxiter=1), dict(maxiter=0)):
        okw = dict(maxiter=51)
        okw.update({k: (bool(v) if k == "accel" else v) for k, v in kw.items()})
        gkw = dict(maxiter=51, kernel=bp.KERNEL_TBLOCK, tblock=3)
        gkw.update(kw)
        assert np.array_equal(ctx.denoise(f, 0.05, _opts(bp, **gkw)), oracle.pdps(f, 0.05, **okw)), kw


def test_tblock_at_config4_size_matches_the_single_pass_kernel(bp, ctx):
    truth, noisy = bp.synthetic_dataset(512, 512, 64, seed=20240601)
    a = ctx.denoise(noisy, 0.1, _opts(bp, maxiter=25, kernel=bp.KERNEL_MARCH))
    for depth in (2, 3, 4):
        b = ctx.denoise(noisy, 0.1, _opts(bp, maxiter=25, kernel=bp.KERNEL_TBLOCK, tblock=depth))
        assert np.array_equal(a, b), depth


def test_config4_full_size_is_bit_identical_to_the_oracle_pins(bp, ctx, ctx32):
    """BASELINE config 4 at FULL size (64 × 512×512, 1000 iterations) through the default (AUTO)
    dispatch: every denoised image hashes to the oracle's (tests/golden/config4_pins.json, written
    by tools/make_config4_pins.py), in fp64 and in fp32, and so does the loss."""
    import hashlib
    import json
    import os
    from conftest import ROOT
    pins = json.load(open(os.path.join(ROOT, "tests", "golden", "config4_pins.json")))
    truth, noisy = bp.synthetic_dataset(pins["M"], pins["N"], pins["O"], seed=pins["seed"])
    assert hashlib.sha256(noisy.tobytes(order="F")).hexdigest() == pins["noisy_sha256"]
    for c, key, dt in ((ctx, "f64", np.float64), (ctx32, "f32", np.float32)):
        c.set_dataset((truth, noisy))
        u, cost, _ = c.learn_eval(pins["lambda"], 0.1, bp.eval_opts(bp.pdps_opts(maxiter=pins["iterations"]), force_branch=3))
        # AUTO: four iterations per HBM pass in fp64, two for strict arithmetic in fp32
        assert c.stats()["pdps_kernel_used"] == bp.KERNEL_TBLOCK and c.stats()["tblock_depth"] == (4 if key == "f64" else 2)
        got = [hashlib.sha256(np.ascontiguousarray(u[:, :, o].astype(dt).T).tobytes()).hexdigest()
               for o in range(pins["O"])]
        assert got == pins[key]["u_sha256"], key
        # an fp32 context also holds ū in fp32 (k/255 rounded), so its loss differs from the fp64 one in the 8th digit
        assert abs(cost - pins[key]["cost"]) <= (1e-12 if key == "f64" else 1e-6) * cost, (key, cost)


def test_temporally_blocked_resident_kernel_is_bit_identical(bp, ctx, ctx32, oracle, datasets):
    """pdps_resident_tb_kernel (two iterations per halo exchange; opt-in, BPLTV_RESIDENT_TB=1): the
    same bits as the oracle in fp64 and fp32, odd and even iteration counts, scalar λ and λ-map, 1 and 10 images."""
    import os
    t, f = datasets["faces_train_128_10"]
    x = np.array([[0.05, 0.1], [0.08, 0.02]])
    am = oracle.patch_upsample(x, 128, 128)
    os.environ["BPLTV_RESIDENT_TB"] = "1"
    bp.reload_env()
    try:
        for O in (1, 10):
            fo = np.asfortranarray(f[:, :, :O])
            for its in (301, 600):
                o = bp.pdps_opts(maxiter=its, kernel=bp.KERNEL_RESIDENT)
                assert np.array_equal(ctx.denoise(fo, 0.08, o), oracle.pdps(fo, 0.08, maxiter=its)), (O, its)
                u32 = ctx32.denoise(fo, 0.08, o)
                assert np.array_equal(u32.astype(np.float32), oracle.pdps(fo, 0.08, maxiter=its, dtype=np.float32)), (O, its)
            o = bp.pdps_opts(maxiter=400, kernel=bp.KERNEL_RESIDENT)
            assert np.array_equal(ctx.denoise(fo, x, o), oracle.pdps(fo, am, maxiter=400)), O
    finally:
        del os.environ["BPLTV_RESIDENT_TB"]
        bp.reload_env()


@pytest.mark.parametrize("prec", [64, 32])
def test_projection_scale_chain_equals_the_ieee_operations(bp, ctx, ctx32, prec):
    """The strict kernels form the projection scale `α / sqrt(n²)` (two correctly rounded operations in the reference) by
    BallScale's branch-free chain (csrc/common.cuh).  bpltv_selftest runs generated operand pairs through the chain and
    through __ddiv_rn(α, __dsqrt_rn(a)) / __fdiv_rn(α, __fsqrt_rn(a)) on the device: not one pair may differ.  Modes: image
    range, the chain's whole operand range, structured significands (rounding boundaries, perfect squares ± 1 ulp) and —
    exhaustive in fp32 — every `a` bit pattern of the range."""
    c = ctx if prec == 64 else ctx32
    for mode, count in ((0, 1 << 32), (1, 1 << 32), (2, 1 << 32), (3, 1 << 32)):
        r = c.selftest(mode, count, seed=2024 + mode)
        assert r["took"] > 0.7 * count, (mode, r)
        assert r["mismatches"] == 0, (prec, mode, r, hex(r["first_a_bits"]), hex(r["first_alpha_bits"]))
    if prec == 32:
        # every float a of the chain's range (119 binades × 2²³ significands), a fresh α for each, three times over
        span = 119 << 23
        for seed in (1, 2, 3):
            r = c.selftest(3, span, seed=seed)
            assert r["took"] >= span - 2 * 119 and r["mismatches"] == 0, (seed, r)    # two guarded significands per binade


def test_resident_kernel_halo_schemes_agree(bp, ctx, ctx32, oracle, datasets, monkeypatch):
    """Kernel B exchanges its halo columns by st.async + mbarrier (default; csrc/halo_async.cuh) or, with
    BPLTV_RESIDENT_ASYNC=0, by DSMEM stores and two cluster barriers per iteration (round 1): both bit-identical to the
    oracle — 1, 4 and 10 images (16- and 8-CTA clusters), λ-map, fp32, and a ragged shape whose last rank owns fewer columns."""
    t, f = datasets["faces_train_128_10"]
    x = np.array([[0.05, 0.1], [0.08, 0.02]])
    am = oracle.patch_upsample(x, 128, 128)
    fr = np.asfortranarray(f[:64, :45, :3])
    for mode in ("1", "0"):
        monkeypatch.setenv("BPLTV_RESIDENT_ASYNC", mode)
        bp.reload_env()
        for O in (1, 4, 10):
            fo = np.asfortranarray(f[:, :, :O])
            o = bp.pdps_opts(maxiter=257, kernel=bp.KERNEL_RESIDENT)
            assert np.array_equal(ctx.denoise(fo, 0.08, o), oracle.pdps(fo, 0.08, maxiter=257)), (mode, O)
            assert np.array_equal(ctx32.denoise(fo, 0.08, o).astype(np.float32),
                                  oracle.pdps(fo, 0.08, maxiter=257, dtype=np.float32)), (mode, O)
        o = bp.pdps_opts(maxiter=200, kernel=bp.KERNEL_RESIDENT)
        assert np.array_equal(ctx.denoise(np.asfortranarray(f[:, :, :2]), x, o), oracle.pdps(f[:, :, :2], am, maxiter=200)), mode
        assert np.array_equal(ctx.denoise(fr, 0.1, o), oracle.pdps(fr, 0.1, maxiter=200)), mode
        assert np.array_equal(ctx.denoise(fr, 0.1, bp.pdps_opts(maxiter=1, kernel=bp.KERNEL_RESIDENT)), oracle.pdps(fr, 0.1, maxiter=1)), mode


def test_large_pageable_stacks_through_a_two_device_context(bp):
    """Pageable caller buffers of 8 MiB or more travel through the context's pinned slots (several host threads, two
    passes over the devices so that a blocking download does not hold back the other device): a two-device context returns
    the bits of a one-device context for a 32 MiB stack, fp64 and fp32, denoise and learn_eval.  Needs ≥ 2 GPUs."""
    import ctypes
    cuda = ctypes.CDLL("libcuda.so.1")
    n = ctypes.c_int(0)
    cuda.cuInit(0); cuda.cuDeviceGetCount(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    truth, noisy = bp.synthetic_dataset(512, 512, 16, seed=11)          # 32 MiB per array, 16 MiB per device
    for prec in (64, 32):
        with bp.Context([0], prec) as c1, bp.Context([0, 1], prec) as c2:
            o = bp.pdps_opts(maxiter=40)
            u1, u2 = c1.denoise(noisy, 0.1, o), c2.denoise(noisy, 0.1, o)
            assert np.array_equal(u1, u2), prec
            c1.set_dataset((truth, noisy)); c2.set_dataset((truth, noisy))
            eo = bp.eval_opts(o, force_branch=3)
            v1, cost1, _ = c1.learn_eval(0.1, 0.1, eo)
            v2, cost2, _ = c2.learn_eval(0.1, 0.1, eo)
            assert np.array_equal(v1, v2) and np.array_equal(v1, u1) and abs(cost1 - cost2) <= 1e-13 * cost1, prec
    # the one-device fp32 result itself is pinned to the oracle at this shape by the config-4 SHA-256 pins


def test_pipelined_host_copies_do_not_change_the_result(bp):
    """denoise(data, x) with PINNED host buffers: the uploads and downloads of image chunks run on a second stream under
    the first and the last passes of the temporally blocked kernel (PipeIO, bpltv_api.cu).  Same bits as the serial
    upload → solve → download, for x⁰ = 0 and x⁰ = f, and the launch count shows that the chunked passes ran."""
    import os
    torch = pytest.importorskip("torch")
    M = N = 64
    O = 16
    _, f = bp.synthetic_dataset(M, N, O, seed=11)
    h_in = torch.empty((O, N, M), dtype=torch.float64).pin_memory()
    h_out = torch.empty((O, N, M), dtype=torch.float64).pin_memory()
    np_in = h_in.numpy().transpose(2, 1, 0)      # Fortran-ordered M×N×O views of the pinned buffers
    np_out = h_out.numpy().transpose(2, 1, 0)
    np_in[...] = f
    for init_mode in (0, 1):
        o = bp.pdps_opts(maxiter=120, kernel=bp.KERNEL_TBLOCK, init_mode=init_mode)
        res = {}
        for pipe in ("1", "0"):
            os.environ["BPLTV_PIPE_IO"] = pipe
            os.environ["BPLTV_PIPE_MIN_MB"] = "0"
            os.environ["BPLTV_PIPE_CHUNKS"] = "4"
            bp.reload_env()
            try:
                with bp.Context([0], 64) as c:
                    np_out[...] = -1.0
                    c.denoise(np_in, 0.1, o, out=np_out)
                    res[pipe] = (np_out.copy(), c.stats()["kernel_launches"], c.stats()["tblock_depth"])
            finally:
                del os.environ["BPLTV_PIPE_IO"], os.environ["BPLTV_PIPE_MIN_MB"], os.environ["BPLTV_PIPE_CHUNKS"]
                bp.reload_env()
        depth = res["0"][2]
        passes = 120 // depth
        assert res["0"][1] == passes
        assert res["1"][1] == passes + 2 * 6 * (4 - 1), res       # 4 chunks × 6 passes at either end instead of 6 whole ones
        assert np.array_equal(res["1"][0], res["0"][0]), init_mode
    # pageable caller buffers (what a Julia Array is): the same pipeline, the chunks through the context's pinned slots
    # when they are large enough for the staging threads (here they are not: the driver's own pageable path)
    os.environ["BPLTV_PIPE_MIN_MB"] = "0"
    os.environ["BPLTV_PIPE_CHUNKS"] = "4"
    os.environ["BPLTV_PIPE_PASSES"] = "6"            # (pageable buffers take 14 passes per chunk phase by default)
    bp.reload_env()
    try:
        with bp.Context([0], 64) as c:
            u = c.denoise(np.asfortranarray(f), 0.1, bp.pdps_opts(maxiter=120, kernel=bp.KERNEL_TBLOCK, init_mode=1))
            assert c.stats()["kernel_launches"] == 120 // c.stats()["tblock_depth"] + 2 * 6 * 3
            assert np.array_equal(u, res["0"][0])            # (the loop's last result is init_mode = 1)
    finally:
        del os.environ["BPLTV_PIPE_MIN_MB"], os.environ["BPLTV_PIPE_CHUNKS"], os.environ["BPLTV_PIPE_PASSES"]
        bp.reload_env()


def test_pipelined_copies_of_a_large_pageable_stack(bp):
    """64 MiB of pageable input and output: the chunks of the pipelined denoise call travel through the staging threads
    (HostStage) on the copy stream (BPLTV_PIPE_PAGEABLE=0: serial upload → solve → download) — same bits either way."""
    import os
    _, f = bp.synthetic_dataset(256, 256, 128, seed=5)
    f = np.asfortranarray(f)
    o = bp.pdps_opts(maxiter=240)
    res = {}
    for pipe in ("1", "0"):
        os.environ["BPLTV_PIPE_PAGEABLE"] = pipe
        bp.reload_env()
        try:
            with bp.Context([0], 64) as c:
                u = c.denoise(f, 0.1, o)
                res[pipe] = (u, c.stats()["kernel_launches"], c.stats()["pdps_kernel_used"])
        finally:
            del os.environ["BPLTV_PIPE_PAGEABLE"]
            bp.reload_env()
    assert res["0"][2] == bp.KERNEL_TBLOCK and res["1"][1] > res["0"][1]
    assert np.array_equal(res["1"][0], res["0"][0])
