// pdps_sumregs.cuh — lower-level solve of the sum-of-regularisers model
//     min_u ½‖u-f‖² + Σ_k ‖α_k ∇_k u‖_{2,1},   ∇₁ forward, ∇₂ backward, ∇₃ centred differences,
// i.e. `sumregs_denoise` of /root/reference/src/SumRegsLearningFunction.jl:38-85 (which calls the
// un-vendored `sumregs_denoise_pdps`; semantics S10-S12 of docs/SEMANTICS.md).
//
// Same accelerated primal-dual recursion as the TV kernels with three dual fields (6 planes).  One
// iteration = two streaming launches that both update in place:
//   primal  x ← prox, x̄ ← over-relaxation   reads x, f, the duals' ±1 neighbours; writes x, x̄
//   dual    y_k ← P_{α_k}(y_k + σ∇_k x̄)       reads x̄'s ±1 neighbours, y; writes y
// (the primal pass writes only x/x̄ and reads only y, the dual pass the reverse, so neither needs a
// second buffer).  Algorithmic traffic: 10 + 13 = 23 words per pixel-iteration; HBM-bound.
// One IEEE operation per operator of the reference expression in strict mode (bit-identical to
// oracle/sumregs.py), FMA / rsqrt in fast mode.
#pragma once
#include "common.cuh"

namespace bpltv {

template <typename Real>
struct SumRegsArgs {
    Real *x, *xb;
    const Real *f;
    Real *y;                 // 6 planes of n = M·N·O: y[(2k+c)·n + pixel], operator k, component c
    const Real *amap;        // 3 maps of M·N (shared by all images) or nullptr
    Real alpha[3];
    StepConsts<Real> sc;
    int M, N, O;
};

template <typename Real, bool STRICT> struct Ar {
    static __device__ __forceinline__ Real add(Real a, Real b) { return STRICT ? StrictOps<Real>::add(a, b) : a + b; }
    static __device__ __forceinline__ Real sub(Real a, Real b) { return STRICT ? StrictOps<Real>::sub(a, b) : a - b; }
    static __device__ __forceinline__ Real mul(Real a, Real b) { return STRICT ? StrictOps<Real>::mul(a, b) : a * b; }
};

template <typename Real, bool STRICT>
__global__ void __launch_bounds__(256) sumregs_primal_kernel(const SumRegsArgs<Real> a)
{
    typedef Ar<Real, STRICT> A;
    const int M = a.M, N = a.N;
    const size_t plane = (size_t)M * N, n = plane * a.O;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int q = (int)(k % plane), i = q % M, j = q / M;
    const Real *y0 = a.y, *y1 = a.y + n, *y2 = a.y + 2 * n, *y3 = a.y + 3 * n, *y4 = a.y + 4 * n, *y5 = a.y + 5 * n;
    const Real z = (Real)0, half = (Real)0.5;
    const bool up = i > 0, dn = i + 1 < M, lf = j > 0, rt = j + 1 < N;
    // ∇ᶠᵀy¹ = (y1(i-1)-y1(i)) + (y2(j-1)-y2(j))
    const Real tF = A::add(A::sub(up ? __ldg(y0 + k - 1) : z, __ldg(y0 + k)), A::sub(lf ? __ldg(y1 + k - M) : z, __ldg(y1 + k)));
    // ∇ᵇᵀy² = (y1(i)-y1(i+1)) + (y2(j)-y2(j+1))
    const Real tB = A::add(A::sub(__ldg(y2 + k), dn ? __ldg(y2 + k + 1) : z), A::sub(__ldg(y3 + k), rt ? __ldg(y3 + k + M) : z));
    // ∇ᶜᵀy³ = ½(y1(i-1)-y1(i+1)) + ½(y2(j-1)-y2(j+1))
    const Real tC = A::add(A::mul(half, A::sub(up ? __ldg(y4 + k - 1) : z, dn ? __ldg(y4 + k + 1) : z)),
                           A::mul(half, A::sub(lf ? __ldg(y5 + k - M) : z, rt ? __ldg(y5 + k + M) : z)));
    const Real dx = A::add(A::add(tF, tB), tC);
    const Real xo = a.x[k], f = __ldg(a.f + k);
    Real xn, xbar;
    if (STRICT) {
        Real t = A::sub(dx, f);
        t = A::mul(a.sc.tau, t);
        t = A::sub(xo, t);
        xn = div_by_const<Real>(t, a.sc.one_p_tau, a.sc.rcp_one_p_tau);
        xbar = A::sub(A::mul(a.sc.one_p_omega, xn), A::mul(a.sc.omega, xo));
    } else {
        xn = fma_(xo, a.sc.inv_one_p_tau, -a.sc.tau_over_one_p_tau * (dx - f));
        xbar = fma_(a.sc.one_p_omega, xn, -a.sc.omega * xo);
    }
    a.x[k] = xn;
    a.xb[k] = xbar;
}

template <typename Real, bool MAP, bool STRICT>
__global__ void __launch_bounds__(256) sumregs_dual_kernel(const SumRegsArgs<Real> a)
{
    typedef Ar<Real, STRICT> A;
    const int M = a.M, N = a.N;
    const size_t plane = (size_t)M * N, n = plane * a.O;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int q = (int)(k % plane), i = q % M, j = q / M;
    const Real z = (Real)0, half = (Real)0.5;
    const bool up = i > 0, dn = i + 1 < M, lf = j > 0, rt = j + 1 < N;
    const Real c = __ldg(a.xb + k);
    const Real xu = up ? __ldg(a.xb + k - 1) : z, xd = dn ? __ldg(a.xb + k + 1) : z;
    const Real xl = lf ? __ldg(a.xb + k - M) : z, xr = rt ? __ldg(a.xb + k + M) : z;
    Real d1[3], d2[3];
    d1[0] = dn ? A::sub(xd, c) : z;                               // ∇ᶠ (S4)
    d2[0] = rt ? A::sub(xr, c) : z;
    d1[1] = up ? A::sub(c, xu) : z;                               // ∇ᵇ (S10)
    d2[1] = lf ? A::sub(c, xl) : z;
    d1[2] = (up && dn) ? A::mul(half, A::sub(xd, xu)) : z;        // ∇ᶜ (S11)
    d2[2] = (lf && rt) ? A::mul(half, A::sub(xr, xl)) : z;
#pragma unroll
    for (int op = 0; op < 3; ++op) {
        Real *p1 = a.y + (size_t)(2 * op) * n + k, *p2 = a.y + (size_t)(2 * op + 1) * n + k;
        Real v1 = *p1, v2 = *p2;
        const Real al = MAP ? __ldg(a.amap + (size_t)op * plane + q) : a.alpha[op];
        dual_update<Real, STRICT, false>(v1, v2, d1[op], d2[op], al, (Real)0, a.sc);
        *p1 = v1; *p2 = v2;
    }
}

}  // namespace bpltv
