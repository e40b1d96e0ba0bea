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
#include "env_switches.h"
#ifndef BPLTV_EMU      // tests/emu/emu_cuda.h supplies cooperative_groups::this_cluster() on OS threads
#include <cooperative_groups.h>
#endif

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "halo_async.cuh"

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
    Real v1[3], v2[3], al[3];
#pragma unroll
    for (int op = 0; op < 3; ++op) {
        v1[op] = a.y[(size_t)(2 * op) * n + k]; v2[op] = a.y[(size_t)(2 * op + 1) * n + k];
        al[op] = MAP ? __ldg(a.amap + (size_t)op * plane + q) : a.alpha[op];
    }
    dual_update_n<Real, STRICT, 3>(v1, v2, d1, d2, al, a.sc);     // the three operators' projection chains interleave
#pragma unroll
    for (int op = 0; op < 3; ++op) { a.y[(size_t)(2 * op) * n + k] = v1[op]; a.y[(size_t)(2 * op + 1) * n + k] = v2[op]; }
}

// ---------------------------------------------------------------------------
// Resident variant: one launch = the complete solve, one image per thread-block cluster (the
// sum-of-regularisers counterpart of pdps_resident.cuh; at the reference's 128×128 the streaming pair above
// is launch-bound: 10 000 launches ≈ 41 ms for 5000 iterations).  CTA `rank` owns NC consecutive columns;
// x and f (and the three λ values of a map) stay in registers, x̄ and the six dual planes in shared memory
// with the halo columns their stencils need: x̄ ±1 column, y[1] (forward, component 2) one to the left,
// y[3] (backward, component 2) one to the right, y[5] (centred, component 2) both.  Owners PUSH their
// boundary columns into the neighbours' halo slots through distributed shared memory.  Round 2: the pushes are
// `st.async` stores counted on mbarriers of the RECEIVER (halo_async.cuh) — four per CTA: what the left / the right
// neighbour sends in the primal phase (one x̄ column each) and in the dual phase (two dual columns each) — and the two
// phases of an iteration are separated by CTA barriers; round 1's two cluster barriers per iteration (a GPU-scope memory
// fence each) are gone.  The readers of a halo column are the threads of this CTA's first / last column, and they are
// the ones that send the columns the neighbour waits for before it overwrites that halo: no slot is double-buffered.
// Same operations in the same order as the streaming kernels: bit-identical.
// ---------------------------------------------------------------------------
constexpr int SRR_THREADS = 512;

template <typename Real>
struct SumRegsResArgs {
    const Real *f;
    Real *u_out;
    const Real *amap;            // 3 maps of M·N or nullptr
    const StepConsts<Real> *steps;
    Real alpha[3];
    int maxiter, M, N, O, init_mode, NC;
};

// bytes of the planes of one CTA (the four halo mbarriers follow them)
template <typename Real>
static __host__ __device__ __forceinline__ size_t sumregs_resident_plane_bytes(int NC, int M)
{
    return (((size_t)(7 * NC + 6) * M * sizeof(Real)) + 15) & ~(size_t)15;
}

template <typename Real, int KP, bool MAP, bool STRICT>
__global__ void __launch_bounds__(SRR_THREADS, 1) sumregs_resident_kernel(const SumRegsResArgs<Real> a)
{
    namespace cg = cooperative_groups;
    typedef Ar<Real, STRICT> A;
#ifdef BPLTV_EMU
    unsigned char *srr_smem = reinterpret_cast<unsigned char *>(emu::dyn_smem());
#else
    extern __shared__ __align__(16) unsigned char srr_smem[];
#endif
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int o = blockIdx.x / CS;
    const int M = a.M, N = a.N, NC = a.NC;
    const int c_begin = rank * NC;
    const int nc = min(NC, N - c_begin);          // ≥ 1 by the plan
    // planes (column pitch M): xb NC+2 columns (slot = local column + 1), y0 NC, y1 NC+1 (slot = c + 1),
    // y2 NC, y3 NC+1 (slot = c), y4 NC, y5 NC+2 (slot = c + 1)
    Real *xb = reinterpret_cast<Real *>(srr_smem);
    Real *y0 = xb + (size_t)(NC + 2) * M;
    Real *y1 = y0 + (size_t)NC * M;
    Real *y2 = y1 + (size_t)(NC + 1) * M;
    Real *y3 = y2 + (size_t)NC * M;
    Real *y4 = y3 + (size_t)(NC + 1) * M;
    Real *y5 = y4 + (size_t)NC * M;
    const int total = (7 * NC + 6) * M;
    for (int k = threadIdx.x; k < total; k += blockDim.x) xb[k] = (Real)0;
    // the neighbours' planes (a left neighbour is never the last rank, so its column count is NC)
    Real *xb_l = rank > 0 ? cluster.map_shared_rank(xb, rank - 1) : nullptr;
    Real *xb_r = rank + 1 < CS ? cluster.map_shared_rank(xb, rank + 1) : nullptr;
    Real *y1_r = rank + 1 < CS ? cluster.map_shared_rank(y1, rank + 1) : nullptr;
    Real *y3_l = rank > 0 ? cluster.map_shared_rank(y3, rank - 1) : nullptr;
    Real *y5_l = rank > 0 ? cluster.map_shared_rank(y5, rank - 1) : nullptr;
    Real *y5_r = rank + 1 < CS ? cluster.map_shared_rank(y5, rank + 1) : nullptr;
    // mbarriers behind the planes: [0] primal-phase bytes from the left neighbour, [1] from the right, [2] dual-phase
    // bytes from the left, [3] from the right
    unsigned long long *hbar = reinterpret_cast<unsigned long long *>(srr_smem + sumregs_resident_plane_bytes<Real>(NC, M));
    unsigned long long *hbar_l = rank > 0 ? cluster.map_shared_rank(hbar, rank - 1) : nullptr;        // I am its RIGHT neighbour: [1], [3]
    unsigned long long *hbar_r = rank + 1 < CS ? cluster.map_shared_rank(hbar, rank + 1) : nullptr;   // I am its LEFT neighbour: [0], [2]
    const bool has_l = rank > 0, has_r = rank + 1 < CS;
    const unsigned colbytes = (unsigned)(M * sizeof(Real));
    if (threadIdx.x == 0)
        for (int b = 0; b < 4; ++b) halo_bar_init(hbar + b);

    const size_t plane = (size_t)M * N;
    const size_t img = (size_t)o * plane;
    const int npix = nc * M;
    Real x[KP], f[KP], al[KP][3];
    int pos[KP];          // c·M + i of the thread's k-th pixel
    unsigned flg[KP];     // 1 valid, 2 up, 4 down, 8 left, 16 right, 32 first local column, 64 last local column
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const int l = threadIdx.x + k * blockDim.x;
        x[k] = f[k] = (Real)0;
        al[k][0] = a.alpha[0]; al[k][1] = a.alpha[1]; al[k][2] = a.alpha[2];
        pos[k] = 0; flg[k] = 0;
        if (l < npix) {
            const int c = l / M, i = l - c * M, jg = c_begin + c;
            pos[k] = l;
            flg[k] = 1u | (i > 0 ? 2u : 0u) | (i + 1 < M ? 4u : 0u) | (jg > 0 ? 8u : 0u) | (jg + 1 < N ? 16u : 0u) |
                     (c == 0 ? 32u : 0u) | (c == nc - 1 ? 64u : 0u);
            const size_t q = (size_t)c_begin * M + l;           // pixel index inside the image
            f[k] = a.f[img + q];
            x[k] = a.init_mode ? f[k] : (Real)0;
            if (MAP) { al[k][0] = a.amap[q]; al[k][1] = a.amap[plane + q]; al[k][2] = a.amap[2 * plane + q]; }
        }
    }
    cluster.sync();   // planes zeroed and mbarriers initialised everywhere before anyone pushes into a halo

    const Real z = (Real)0, half = (Real)0.5;
    for (int it = 0; it < a.maxiter; ++it) {
        const StepConsts<Real> sc = a.steps[it];
        // the neighbours' dual columns of the previous iteration; then this iteration's phases of the four mbarriers
        if (it > 0) {
            if (has_l) halo_wait(hbar + 2, it - 1, 2 * colbytes);
            if (has_r) halo_wait(hbar + 3, it - 1, 2 * colbytes);
        }
        if (threadIdx.x == 0) {
            if (has_l) { halo_bar_arm(hbar + 0, colbytes); halo_bar_arm(hbar + 2, 2 * colbytes); }
            if (has_r) { halo_bar_arm(hbar + 1, colbytes); halo_bar_arm(hbar + 3, 2 * colbytes); }
        }
        // ---- primal: x ← prox, x̄ ← over-relaxation (reads the duals, writes x̄) ----
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const unsigned fl = flg[k];
            if (!(fl & 1u)) continue;
            const bool up = fl & 2u, dn = fl & 4u, lf = fl & 8u, rt = fl & 16u, first = fl & 32u, last = fl & 64u;
            const int p0 = pos[k], p1 = p0 + M;
            const Real tF = A::add(A::sub(up ? y0[p0 - 1] : z, y0[p0]), A::sub(lf ? y1[p0] : z, y1[p1]));
            const Real tB = A::add(A::sub(y2[p0], dn ? y2[p0 + 1] : z), A::sub(y3[p0], rt ? y3[p1] : z));
            const Real tC = A::add(A::mul(half, A::sub(up ? y4[p0 - 1] : z, dn ? y4[p0 + 1] : z)),
                                   A::mul(half, A::sub(lf ? y5[p0] : z, rt ? y5[p1 + M] : z)));
            const Real dx = A::add(A::add(tF, tB), tC);
            const Real xo = x[k];
            Real xn, xbar;
            if (STRICT) {
                Real t = A::sub(dx, f[k]);
                t = A::mul(sc.tau, t);
                t = A::sub(xo, t);
                xn = div_by_const<Real>(t, sc.one_p_tau, sc.rcp_one_p_tau);
                xbar = A::sub(A::mul(sc.one_p_omega, xn), A::mul(sc.omega, xo));
            } else {
                xn = fma_(xo, sc.inv_one_p_tau, -sc.tau_over_one_p_tau * (dx - f[k]));
                xbar = fma_(sc.one_p_omega, xn, -sc.omega * xo);
            }
            x[k] = xn;
            xb[p1] = xbar;
            if (first && xb_l) halo_push1(xb_l + (NC + 1) * M + p0, xbar, hbar_l + 1);        // the left CTA's right halo (first column: p0 = row)
            if (last && xb_r) halo_push1(xb_r + (p0 - (nc - 1) * M), xbar, hbar_r + 0);        // the right CTA's left halo
        }
        __syncthreads();                                      // x̄ of this CTA's own columns
        if (has_l) halo_wait(hbar + 0, it, colbytes);         // x̄ of the columns next to them
        if (has_r) halo_wait(hbar + 1, it, colbytes);
        // ---- dual: y_k ← P_{α_k}(y_k + σ∇_k x̄) (reads x̄, writes the duals) ----
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const unsigned fl = flg[k];
            if (!(fl & 1u)) continue;
            const bool up = fl & 2u, dn = fl & 4u, lf = fl & 8u, rt = fl & 16u, first = fl & 32u, last = fl & 64u;
            const int p0 = pos[k], p1 = p0 + M;
            const Real cc = xb[p1];
            const Real xu = up ? xb[p1 - 1] : z, xd = dn ? xb[p1 + 1] : z;
            const Real xl = lf ? xb[p0] : z, xr = rt ? xb[p1 + M] : z;
            Real d1[3], d2[3];
            d1[0] = dn ? A::sub(xd, cc) : z;                               // ∇ᶠ (S4)
            d2[0] = rt ? A::sub(xr, cc) : z;
            d1[1] = up ? A::sub(cc, xu) : z;                               // ∇ᵇ (S10)
            d2[1] = lf ? A::sub(cc, xl) : z;
            d1[2] = (up && dn) ? A::mul(half, A::sub(xd, xu)) : z;         // ∇ᶜ (S11)
            d2[2] = (lf && rt) ? A::mul(half, A::sub(xr, xl)) : z;
            Real v1[3] = {y0[p0], y2[p0], y4[p0]}, v2[3] = {y1[p1], y3[p0], y5[p1]};
            const Real alk[3] = {al[k][0], al[k][1], al[k][2]};
            dual_update_n<Real, STRICT, 3>(v1, v2, d1, d2, alk, sc);    // the three operators' projection chains interleave
            y0[p0] = v1[0]; y1[p1] = v2[0];
            if (last && y1_r) halo_push1(y1_r + (p0 - (nc - 1) * M), v2[0], hbar_r + 2);      // forward, component 2: the right CTA reads column c0−1
            y2[p0] = v1[1]; y3[p0] = v2[1];
            if (first && y3_l) halo_push1(y3_l + NC * M + p0, v2[1], hbar_l + 3);            // backward, component 2: the left CTA reads column c0+NC
            y4[p0] = v1[2]; y5[p1] = v2[2];
            if (first && y5_l) halo_push1(y5_l + (NC + 1) * M + p0, v2[2], hbar_l + 3);      // centred, component 2: both neighbours
            if (last && y5_r) halo_push1(y5_r + (p0 - (nc - 1) * M), v2[2], hbar_r + 2);
        }
        __syncthreads();                                      // the duals of this CTA's own columns
    }
    // nothing may still be in flight towards this CTA's shared memory when it exits
    if (a.maxiter > 0) {
        if (has_l) halo_wait(hbar + 2, a.maxiter - 1, 2 * colbytes);
        if (has_r) halo_wait(hbar + 3, a.maxiter - 1, 2 * colbytes);
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const int l = threadIdx.x + k * blockDim.x;
        if (l < npix) a.u_out[img + (size_t)c_begin * M + l] = x[k];
    }
}

struct SumRegsResPlan {
    bool ok = false;
    int CS = 0, NC = 0, KP = 0;
    size_t smem = 0;
};

template <typename Real>
static inline SumRegsResPlan sumregs_resident_plan(size_t smem_optin, int M, int N, int cs_max)
{
    SumRegsResPlan p;
    const int cs_cands[5] = {16, 8, 4, 2, 1};
    for (int ci = 0; ci < 5; ++ci) {
        const int CS = cs_cands[ci];
        if (CS > cs_max || CS > N) continue;
        const int NC = (N + CS - 1) / CS;
        if ((CS - 1) * NC >= N) continue;                 // every rank owns at least one column
        const int KP = (NC * M + SRR_THREADS - 1) / SRR_THREADS;
        if (KP > 8) continue;
        const size_t smem = sumregs_resident_plane_bytes<Real>(NC, M) + 32;     // planes + four halo mbarriers
        if (smem > smem_optin) continue;
        p.ok = true; p.CS = CS; p.NC = NC; p.KP = KP <= 1 ? 1 : (KP <= 2 ? 2 : (KP <= 4 ? 4 : 8)); p.smem = smem;
        return p;
    }
    return p;
}

#ifndef BPLTV_EMU      // host-side launch (CUDA runtime)
template <typename Real, int KP>
static inline cudaError_t launch_sumregs_resident_kp(const SumRegsResArgs<Real> &a, const SumRegsResPlan &p, bool map,
                                                     bool strict, cudaStream_t st)
{
    void (*fn)(const SumRegsResArgs<Real>);
    if (map) fn = strict ? sumregs_resident_kernel<Real, KP, true, true> : sumregs_resident_kernel<Real, KP, true, false>;
    else fn = strict ? sumregs_resident_kernel<Real, KP, false, true> : sumregs_resident_kernel<Real, KP, false, false>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    if (p.CS > 8) {
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.O * p.CS));
    cfg.blockDim = dim3(SRR_THREADS);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) != cudaSuccess || nclusters < 1 ||
        (p.CS > 8 && nclusters < a.O)) {     // 16-CTA clusters only when the whole batch is co-resident
        cudaGetLastError();
        return cudaErrorLaunchOutOfResources;
    }
    return cudaLaunchKernelEx(&cfg, fn, a);
}

// cudaSuccess: launched.  cudaErrorInvalidValue / cudaErrorLaunchOutOfResources: not eligible (stream instead).
template <typename Real>
static inline cudaError_t launch_sumregs_resident(SumRegsResArgs<Real> a, size_t smem_optin, bool map, bool strict,
                                                  cudaStream_t st)
{
    const char *cs_env = bpltv::env_get("BPLTV_RESIDENT_CS");
    const int cs_cap = cs_env && *cs_env ? atoi(cs_env) : 16;
    const int caps[2] = {std::min(cs_cap, 16), std::min(cs_cap, 8)};
    cudaError_t last = cudaErrorInvalidValue;
    for (int t = 0; t < 2; ++t) {
        if (t == 1 && caps[1] == caps[0]) break;
        const SumRegsResPlan p = sumregs_resident_plan<Real>(smem_optin, a.M, a.N, caps[t]);
        if (!p.ok) continue;
        if (t == 1 && p.CS > 8) continue;
        a.NC = p.NC;
        last = p.KP == 1 ? launch_sumregs_resident_kp<Real, 1>(a, p, map, strict, st)
             : p.KP == 2 ? launch_sumregs_resident_kp<Real, 2>(a, p, map, strict, st)
             : p.KP == 4 ? launch_sumregs_resident_kp<Real, 4>(a, p, map, strict, st)
                         : launch_sumregs_resident_kp<Real, 8>(a, p, map, strict, st);
        if (last == cudaSuccess) return last;
        cudaGetLastError();
    }
    return last;
}
#endif  // BPLTV_EMU

}  // namespace bpltv
