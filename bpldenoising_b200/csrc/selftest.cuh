// selftest.cuh — arithmetic self-test of the strict kernels' projection scale (bpltv_selftest, include/bpltv.h).
//
// The strict PDPS kernels form `α / sqrt(n²)` (the reference's projection onto the λ-ball, external
// op_denoise_pdps; docs/SEMANTICS.md S6) by BallScale's straight-line chain instead of the compiler's sqrt.rn /
// div.rn expansions (common.cuh).  This kernel runs operand pairs through both and counts the pairs whose bits
// differ: the test-suite asserts zero (tests/test_gpu_pdps.py).  Operands come from a counter-based generator, so
// a run is reproducible from (mode, count, seed).
#pragma once
#include "common.cuh"

namespace bpltv {

static __host__ __device__ __forceinline__ unsigned long long st_mix(unsigned long long x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

template <typename Real> struct StBits;
template <> struct StBits<double> {
    static constexpr int SIG = 52, BIAS = 1023, EA = 499, EL = 199;
    static __device__ __forceinline__ double make(int e, unsigned long long sig)
    {
        return __longlong_as_double((long long)(((unsigned long long)(BIAS + e) << SIG) | (sig & ((1ull << SIG) - 1))));
    }
    static __device__ __forceinline__ unsigned long long bits(double v) { return (unsigned long long)__double_as_longlong(v); }
};
template <> struct StBits<float> {
    static constexpr int SIG = 23, BIAS = 127, EA = 59, EL = 29;
    static __device__ __forceinline__ float make(int e, unsigned long long sig)
    {
        return __uint_as_float(((unsigned)(BIAS + e) << SIG) | (unsigned)(sig & ((1ull << SIG) - 1)));
    }
    static __device__ __forceinline__ unsigned long long bits(float v) { return __float_as_uint(v); }
};

// significands chosen to sit on the rounding boundaries of √ and ÷
template <typename Real>
static __device__ unsigned long long st_hard_sig(unsigned long long h)
{
    constexpr int S = StBits<Real>::SIG;
    const unsigned long long ones = (1ull << S) - 1;
    const unsigned long long r = st_mix(h);
    switch (h & 7) {
    case 0: return ones - ((h >> 3) & 15);                 // all ones, all ones minus a few ulps
    case 1: return (h >> 3) & 15;                          // a power of two plus a few ulps
    case 2: return 1ull << (r % S);                        // one bit set
    case 3: return ones ^ (1ull << (r % S));               // one bit cleared
    case 4: return r & (ones << (S / 2));                  // low half zero
    case 5: return r & (ones >> (S / 2));                  // high half zero
    default: return r;
    }
}

// mode 0: the range image data produces (a = n² ∈ [2⁻²⁴, 2⁶), α ∈ [2⁻¹⁴, 2²)); 1: the chain's whole operand range,
// exponents uniform; 2: structured significands (rounding boundaries, perfect squares ± 1 ulp) over the whole range;
// 3: every `a` pattern of the range in turn (exhaustive for fp32 when count ≥ 2³⁰), α random
template <typename Real>
__global__ void __launch_bounds__(256) selftest_ball_scale_kernel(int mode, unsigned long long count, unsigned long long seed,
                                                                 unsigned long long *result)
{
    typedef StBits<Real> B;
    unsigned long long took = 0, bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long h1 = st_mix(seed + 2 * i), h2 = st_mix(seed + 2 * i + 1);
        Real a, al;
        if (mode == 0) {
            a = B::make(-24 + (int)((h1 >> 54) % 30), h1);
            al = B::make(-14 + (int)((h2 >> 54) % 16), h2);
        } else if (mode == 1) {
            a = B::make(-B::EA + (int)((h1 >> 53) % (2 * B::EA + 1)), h1);
            al = B::make(-B::EL + (int)((h2 >> 53) % (2 * B::EL + 1)), h2);
        } else if (mode == 2) {
            const int ea = -B::EA + (int)((h1 >> 53) % (2 * B::EA + 1));
            if (((h1 >> 40) & 3) == 0) {        // a perfect square of a half-width integer, and its neighbours
                const unsigned long long r = (1ull << (B::SIG / 2)) + (st_mix(h1) & ((1ull << (B::SIG / 2)) - 1));
                const Real sq = (Real)r * (Real)r;                    // exact: r has ≤ SIG/2 + 1 bits
                const long long nb = (long long)((h1 >> 44) % 3) - 1;
                const int e0 = (int)(B::bits(sq) >> B::SIG) - B::BIAS;
                a = B::make(e0 + 2 * ((ea - e0) / 2), B::bits(sq) + nb);   // scaled by a power of 4: still a square
            } else {
                a = B::make(ea, st_hard_sig<Real>(h1));
            }
            al = B::make(-B::EL + (int)((h2 >> 53) % (2 * B::EL + 1)), ((h2 >> 41) & 1) ? st_hard_sig<Real>(h2) : h2);
        } else {
            const unsigned long long span = (unsigned long long)(2 * B::EA + 1) << B::SIG;
            const unsigned long long k = i % span;
            a = B::make(-B::EA + (int)(k >> B::SIG), k);
            al = B::make(-B::EL + (int)((h2 >> 53) % (2 * B::EL + 1)), h2);
        }
        if (!BallScale<Real>::fast_ok(a, al)) continue;    // the kernels send these through the IEEE operations
        ++took;
        const Real got = BallScale<Real>::eval(a, al);
        const Real want = StrictOps<Real>::div(al, StrictOps<Real>::sqrt(a));
        if (B::bits(got) != B::bits(want)) {
            if (bad == 0 && atomicCAS(result + 2, 0ull, B::bits(a) | (1ull << 63)) == 0ull) result[3] = B::bits(al);
            ++bad;
        }
    }
    if (took) atomicAdd(result + 0, took);
    if (bad) atomicAdd(result + 1, bad);
}

}  // namespace bpltv
