"""Sum-of-regularisers path (/root/reference/src/SumRegsLearningFunction.jl) through the C ABI against
oracle/sumregs.py.  Tolerances as for the TV path: bit-identical iterates in strict arithmetic, relative
L2 ≤ 1e-10 (fp64) / 1e-5 (fp32) otherwise."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sr():
    from oracle import sumregs
    return sumregs


@pytest.mark.parametrize("shape", [(64, 64, 2), (33, 17, 3), (1, 9, 1), (9, 1, 2), (2, 2, 1), (3, 5, 1), (128, 128, 1)])
def test_sumregs_denoise_is_bit_identical_to_the_oracle(bp, ctx, ctx32, sr, shape):
    M, N, O = shape
    rng = np.random.default_rng(M * 31 + N)
    f = np.asfortranarray(np.round(rng.uniform(0, 1, shape) * 255) / 255)
    x = np.array([0.03, 0.012, 0.05])
    ref = sr.sumregs_pdps(f, list(x), maxiter=60)
    u = ctx.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=60))
    assert np.array_equal(u, ref), (shape, np.abs(u - ref).max())
    uf = ctx.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=60, arith=bp.FAST, init_mode=1))
    assert rel_l2(uf, sr.sumregs_pdps(f, list(x), maxiter=60, init_mode=1)) <= 1e-10
    u32 = ctx32.sumregs_denoise(f, x, bp.sumregs_pdps_opts(maxiter=60))
    assert np.array_equal(u32.astype(np.float32), sr.sumregs_pdps(f, list(x), maxiter=60, dtype=np.float32))
    # patch parameter m×n×3 (:62-85): three up-sampled maps
    xp = rng.uniform(0.005, 0.08, (2, 3, 3)) if M >= 2 and N >= 3 else np.tile(x, (1, 1, 1))
    from oracle import oracle as orc
    maps = [orc.patch_upsample(xp[:, :, k], M, N) for k in range(3)]
    up = ctx.sumregs_denoise(f, xp, bp.sumregs_pdps_opts(maxiter=40))
    assert np.array_equal(up, sr.sumregs_pdps(f, maps, maxiter=40)), shape


def test_sumregs_denoise_on_the_reference_dataset(bp, ctx, sr, datasets):
    t, f = datasets["cameraman_128_5"]
    x0 = np.array([0.001, 0.001, 0.001])       # α₀ of scalar_bilevel_sumregs_learn (BPLDenoising.jl:429)
    u = bp.sumregs_denoise(f, x0, ctx=ctx, maxiter=400)
    assert np.array_equal(u, sr.sumregs_pdps(f, list(x0), maxiter=400))
    # one regulariser switched on alone is the corresponding TV model: forward-only = the TV path
    # (same recursion, R_K = √18 on both sides)
    v = ctx.sumregs_denoise(f, np.array([0.08, 0.0, 0.0]), bp.sumregs_pdps_opts(maxiter=300))
    w = ctx.denoise(f, 0.08, bp.pdps_opts(maxiter=300, opnorm=18 ** 0.5))
    assert rel_l2(v, w) <= 1e-12
    # properties: denoising reduces every one of the three total variations
    big = ctx.sumregs_denoise(f, np.array([0.03, 0.03, 0.03]), bp.sumregs_pdps_opts(maxiter=1500))
    for kind in sr.KINDS:
        def tv(a):
            d1, d2 = sr._grad(kind, a)
            return np.sqrt(d1 * d1 + d2 * d2).sum()
        assert tv(big[:, :, 0]) < 0.6 * tv(f[:, :, 0]), kind
    with pytest.raises(ValueError):
        ctx.sumregs_denoise(f, np.array([0.1, 0.1]))
    with pytest.raises(bp.BpltvError):
        ctx.sumregs_denoise(f, np.array([0.1, -0.1, 0.1]))
    with pytest.raises(bp.BpltvError):
        ctx.sumregs_denoise(f, x0, bp.sumregs_pdps_opts(rho=0.5, maxiter=2))
