// lu_band.cuh — node-space band LU for the regularised adjoint systems (fp64).
//
// Written for the patch variant of sumregs_gradient_reg, /root/reference/src/SumRegsLearningFunction.jl:195-262:
//     p = (I + x₁[:] .* G₁ᵀ(B₁−C₁)G₁ + x₂[:] .* G₂ᵀ(B₂−C₂)G₂ + x₃[:] .* G₃ᵀ(B₃−C₃)G₃) \ (ū − u)       (:246)
//     g_k = p ⊙ G_kᵀ(Act_k Den_k G_k u + γ Inact_k G_k u),   gx[:,:,k] = calc_adjoint(pOp, g_k)          (:247-259)
// `x_k[:] .*` scales the ROWS of each term by a different λ-map, so the matrix is not symmetric and has
// no compliance (Cholesky) form.  It is a banded n²×n² matrix on the column-major nodes, though: the
// stencils of G_kᵀ T G_k reach ±2 rows and ±2 columns, half-bandwidth 2n.  The same kernels carry the scalar
// sumregs_gradient_reg (:112-167) and, with the forward operator alone (nops = 1, half-bandwidth n), gradient_reg
// of the TV learning function (/root/reference/src/TVLearningFunctionVec.jl:137-161, :192-215); see gradient_lu.cuh.
// Per image, one CTA or one thread-block cluster:
//   lu3_classify   per (pixel, operator): the 2×2 tensor T = B − C and the functional weights
//   lu3_assemble   one thread per matrix row, 13 stencil offsets, fixed summation order (deterministic)
//   lu_factor      blocked right-looking band LU without pivoting, NB = 16: the 16×16 diagonal block in one
//                  warp's registers (shuffles, Newton reciprocal of the pivots), L21 rows / U12 columns by
//                  substitution (one thread each; inputs staged with cp.async under the diagonal block's
//                  factorisation, panels kept in shared memory, L21 transposed), rank-16 update of the bw×bw
//                  trailing window in 32×32 warp tiles (8×4 per thread), optionally dealt over a cluster
//   lu3_solve      blocked forward / backward substitution with the vector in shared memory and the factor
//                  entries prefetched one block step ahead, iterative refinement against the MATRIX-FREE
//                  residual (stencils + tensors, independent of the assembled band), functional,
//                  PatchOp-adjoint sums
// No pivoting: every term is (positive diagonal)·(PSD); for equal maps the matrix is D·SPD, whose LU needs
// none.  A vanished pivot, or a solve whose normwise backward error stays above 1e-11 after refinement,
// poisons the output with NaN, which the API reports as BPLTV_ERR_NUMERIC — no silent wrong answer.
// Band storage: row i holds columns i−bwx … i+bwx at ab[i·LD + (j−i+bwx)], bwx = bw + NB, so that every
// index a block step forms is inside the row's storage (entries outside the true band stay exactly zero);
// LD = 2·bwx + 1 is odd, which makes every window segment that starts at an even column 16-byte aligned.
//
// This header holds device code only and depends on sumregs_stencils.cuh alone: tests/emu/ compiles it with
// g++ and runs the kernels on OS threads (CPU check of the index arithmetic, the barriers and the 16-byte
// alignment of every vector access; no GPU needed).
#pragma once
#ifndef BPLTV_EMU
#include <cooperative_groups.h>
#include <cuda_pipeline.h>
#endif

#include "sumregs_stencils.cuh"

namespace bpltv {

constexpr int LU_NB = 16;
constexpr int LU_DP = LU_NB + 1;        // padded row length of the 16×16 blocks and of the L panel
#ifndef LU_THREADS
#define LU_THREADS 512
#endif
constexpr int LU_PLANES = 22;
// per-slot planes (N doubles each): [6k+0..3] T11 T12 T21 T22, [6k+4..5] w1 w2 of operator k;
// [18] r = ū − u; [19] p; [20] residual / correction; [21] per-node functional
constexpr int LU_PL_R = 18, LU_PL_P = 19, LU_PL_WORK = 20, LU_PL_F = 21;

struct LuSlots {
    double *ab;   size_t ab_stride;     // per slot: N rows × LD doubles
    double *pix;  size_t pix_stride;    // per slot: LU_PLANES planes
    int *info;                          // per slot: 4 ints, [0] ≠ 0 → a pivot vanished
    int n, N, bw, bwx, LD;
    int use_pin;                        // lu_factor: stage the panel inputs in shared memory with cp.async
    int nops;                           // 3: forward, backward, centred (sum of regularisers); 1: forward only (TV)
};

struct Lu3Params {
    double alpha[3];     // scalar parameters when alpha_maps == nullptr
    double gamma;
    int lm, ln, refine;
};

// 16-byte accesses: the thread emulation (tests/emu, -DBPLTV_EMU) checks the alignment the GPU requires
#ifdef BPLTV_EMU
#define LU_A16(p) (emu::check_aligned16(p), (p))
#else
#define LU_A16(p) (p)
#endif

// dynamic shared memory and the thread-block cluster, as the GPU or as the thread emulation provides them
#ifdef BPLTV_EMU
#define LU_DYN_SMEM(name) double *name = emu::dyn_smem()
static inline int lu_cluster_rank() { return emu::cluster_rank(); }
static inline int lu_cluster_size() { return emu::cluster_size(); }
static inline void lu_cluster_sync() { emu::cluster_sync(); }
#else
#define LU_DYN_SMEM(name) extern __shared__ __align__(16) double name[]
static __device__ __forceinline__ int lu_cluster_rank() { return (int)cooperative_groups::this_cluster().block_rank(); }
static __device__ __forceinline__ int lu_cluster_size() { return (int)cooperative_groups::this_cluster().num_blocks(); }
static __device__ __forceinline__ void lu_cluster_sync() { cooperative_groups::this_cluster().sync(); }
#endif

// asynchronous global → shared copies (cp.async: no registers held while the data is in flight) and a fast
// reciprocal for the pivot chain (hardware seed + two Newton steps: ≤ 1 ulp, deterministic)
#ifdef BPLTV_EMU
static inline void lu_cp_async16(double *dst, const double *src) { emu::check_aligned16(dst); emu::check_aligned16(src); dst[0] = src[0]; dst[1] = src[1]; }
static inline void lu_cp_async8(double *dst, const double *src) { dst[0] = src[0]; }
static inline void lu_cp_async_commit() {}
static inline void lu_cp_async_wait() {}
static inline double lu_rcp(double a) { return 1.0 / a; }
#else
static __device__ __forceinline__ void lu_cp_async16(double *dst, const double *src) { __pipeline_memcpy_async(dst, src, 16); }
static __device__ __forceinline__ void lu_cp_async8(double *dst, const double *src) { __pipeline_memcpy_async(dst, src, 8); }
static __device__ __forceinline__ void lu_cp_async_commit() { __pipeline_commit(); }
static __device__ __forceinline__ void lu_cp_async_wait() { __pipeline_wait_prior(0); }
static __device__ __forceinline__ double lu_rcp(double a)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    if (!(fabs(r) > 0.0) || !(fabs(r) < 1e300)) return 1.0 / a;     // 0, inf, NaN, denormal: the IEEE path
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}
#endif

static __device__ __forceinline__ double lu_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the CTA, valid in every thread; `red` = 33 doubles of shared memory
static __device__ double lu_block_sum(double v, double *red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    v = lu_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double r = lane < nwarps ? red[lane] : 0.0;
        r = lu_warp_sum(r);
        if (lane == 0) red[32] = r;
    }
    __syncthreads();
    return red[32];
}

// max over the CTA, valid in every thread
static __device__ double lu_block_max(double v, double *red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double r = lane < nwarps ? red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
        if (lane == 0) red[32] = r;
    }
    __syncthreads();
    return red[32];
}

// ---------------------------------------------------------------------------
// K1: T_kq = B − C and w_kq of every (pixel, operator); r = ū − u     (:203-241, :246)
// ---------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(256) lu3_classify_kernel(LuSlots ws, double gamma, const Real *u_all,
                                                           const Real *ubar_all, int img0)
{
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const Real *u = u_all + (size_t)(img0 + slot) * N;
    const Real *ub = ubar_all + (size_t)(img0 + slot) * N;
    double *pix = ws.pix + ws.pix_stride * slot;
    for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < N; q += gridDim.y * blockDim.x) {
        const int i = q % n, j = q / n;
        pix[(size_t)LU_PL_R * N + q] = (double)ub[q] - (double)u[q];
        for (int k = 0; k < ws.nops; ++k) {
            double g1, g2;
            op_apply<Real>(k, i, j, n, u, q, g1, g2);
            const double nrm = sqrt(g1 * g1 + g2 * g2);
            const bool act = fmax(0.0, nrm - 1.0 / gamma) != 0.0;     // :207-208
            double t11, t12, t21, t22, w1, w2;
            if (act) {   // −C = Den − prodesc(Gu/den³, Gu) (:211-215); w = Den·Gu
                const double den = nrm, id = 1.0 / den, d3 = den * den * den;
                const double a1 = g1 / d3, a2 = g2 / d3;
                t11 = id - a1 * g1; t12 = -(a1 * g2); t21 = -(a2 * g1); t22 = id - a2 * g2;
                w1 = id * g1; w2 = id * g2;
            } else {     // B = γ·Inact; w = γ·Gu
                t11 = gamma; t12 = 0.0; t21 = 0.0; t22 = gamma;
                w1 = gamma * g1; w2 = gamma * g2;
            }
            double *pk = pix + (size_t)(6 * k) * N;
            pk[q] = t11; pk[(size_t)N + q] = t12; pk[(size_t)2 * N + q] = t21; pk[(size_t)3 * N + q] = t22;
            pk[(size_t)4 * N + q] = w1; pk[(size_t)5 * N + q] = w2;
        }
    }
}

template <typename Real>
static __device__ __forceinline__ double lu3_alpha(const Real *alpha_maps, const Lu3Params &pr, int N, int k, int v)
{
    return alpha_maps ? (double)alpha_maps[(size_t)k * N + v] : pr.alpha[k];
}

// ---------------------------------------------------------------------------
// K2: row v of  I + Σ_k diag(α_k) G_kᵀ T_k G_k  (the band must be zero on entry)
// ---------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(256) lu3_assemble_kernel(LuSlots ws, Lu3Params pr, const Real *alpha_maps)
{
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const double *pix = ws.pix + ws.pix_stride * slot;
    double *ab = ws.ab + ws.ab_stride * slot;
    if (blockIdx.y == 0 && threadIdx.x == 0) ws.info[4 * slot] = 0;
    const int offs[13] = {-2 * n, -n - 1, -n, -n + 1, -2, -1, 0, 1, 2, n - 1, n, n + 1, 2 * n};   // distinct for n ≥ 4
    for (int v = blockIdx.y * blockDim.x + threadIdx.x; v < N; v += gridDim.y * blockDim.x) {
        const int i = v % n, j = v / n;
        double acc[13];
#pragma unroll
        for (int s = 0; s < 13; ++s) acc[s] = 0.0;
        acc[6] = 1.0;
        visit_node(i, j, n, [&](int q, int k, double c1, double c2) {
            if (k >= ws.nops) return;
            const double *tk = pix + (size_t)(6 * k) * N;
            const double a = lu3_alpha<Real>(alpha_maps, pr, N, k, v);
            const double v1 = c1 * tk[q] + c2 * tk[(size_t)2 * N + q];                  // (c1 c2)·T
            const double v2 = c1 * tk[(size_t)N + q] + c2 * tk[(size_t)3 * N + q];
            visit_stencil(k, q % n, q / n, n, [&](int node, double e1, double e2) {
                const double val = a * (v1 * e1 + v2 * e2);
                const int d = node - v;
#pragma unroll
                for (int s = 0; s < 13; ++s)
                    if (offs[s] == d) { acc[s] += val; break; }
            });
        });
        double *row = ab + (size_t)v * ws.LD + ws.bwx;
#pragma unroll
        for (int s = 0; s < 13; ++s) {
            const int col = v + offs[s];
            if (col >= 0 && col < N) row[offs[s]] = acc[s];
        }
    }
}

// ---------------------------------------------------------------------------
// K3: in-place band LU, no pivoting.  Dynamic shared memory (doubles):
//   D[NB·DP] diagonal block | rD[NB] reciprocal pivots | Lt[NB·lsp] L panel, TRANSPOSED (column c of the
//   panel contiguous over the rows) | Us[NB·bwp] U panel (row c contiguous over the columns)
// LD is odd (≡ 1 mod 4), so (row·LD + col − row) has the parity of col: every row segment that starts at an
// even column of the window is 16-byte aligned and moves as double2.
// ---------------------------------------------------------------------------
static inline int lu_panel_pitch(int bw) { return (bw + 32 + 3) & ~3; }
constexpr int LU_PINP = LU_NB + 2;      // pitch of a staged panel item: 16-byte aligned, conflict-free 16-byte reads
static inline size_t lu_factor_smem(int bw, bool use_pin)
{
    return (size_t)(LU_NB * LU_DP + LU_NB + 2 * LU_NB * lu_panel_pitch(bw) + (use_pin ? 2 * bw * LU_PINP : 0)) * sizeof(double);
}

// 16×16 LU without pivoting of the block in D (row pitch LU_DP), by ONE warp, in registers: lane c (< 16;
// lanes 16-31 mirror it) holds column c; per pivot the multipliers of column pv come out of lane pv by
// shuffle — no shared-memory round trip on the chain.  Writes the factors back to D and the reciprocal
// pivots to rD; returns true when a pivot vanished (or is NaN).  Its own function so that its 16 column
// registers do not weigh on the register allocation of the trailing update.
#ifndef BPLTV_EMU
__noinline__
#endif
static __device__ bool lu16_warp(double *D, double *rD)
{
    constexpr int NB = LU_NB, DP = LU_DP;
    const int lane = threadIdx.x & 31, cl = lane & 15;
    bool bad = false;
    double acol[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) acol[r] = D[r * DP + cl];
#pragma unroll
    for (int pv = 0; pv < NB; ++pv) {
        const double piv = __shfl_sync(0xffffffffu, acol[pv], pv);
        if (!(fabs(piv) > 0.0)) bad = true;
        const double rp = lu_rcp(piv);
#pragma unroll
        for (int r = pv + 1; r < NB; ++r) {
            const double l = __shfl_sync(0xffffffffu, acol[r], pv) * rp;
            if (cl > pv) acol[r] -= l * acol[pv];
            else if (cl == pv) acol[r] = l;
        }
    }
    if (lane < NB) {
#pragma unroll
        for (int r = 0; r < NB; ++r) D[r * DP + lane] = acol[r];
    }
    __syncwarp();
    if (lane < NB) rD[lane] = 1.0 / D[lane * DP + lane];
    return bad;
}

// CL: the image is factorised by a thread-block CLUSTER of csize CTAs (one SM each).  Every CTA forms the
// diagonal block and both panels redundantly in its own shared memory (cheap), the 32×32 tiles of the
// trailing window are dealt over all warps of the cluster, and two cluster barriers per block step order the
// traffic through global memory: one after every CTA has READ the step's inputs (then the
// factored block and panels are written back in place, each entry by one CTA), one after the trailing update.  Same operations per entry as the
// single-CTA kernel: identical bits for every cluster size.
template <bool CL>
__global__ void __launch_bounds__(LU_THREADS, 1) lu_factor_kernel(LuSlots ws)
{
    LU_DYN_SMEM(lu_fsm);
    constexpr int NB = LU_NB, DP = LU_DP;
    const int crank = CL ? lu_cluster_rank() : 0, csize = CL ? lu_cluster_size() : 1;
    const int slot = CL ? (int)blockIdx.x / csize : (int)blockIdx.x;
    const int N = ws.N, bw = ws.bw, bwx = ws.bwx, LD = ws.LD;
    double *ab = ws.ab + ws.ab_stride * slot;
    const int lsp = (bw + 32 + 3) & ~3;           // pitch of both panels (32 spare entries: ragged tiles read, never use them)
    double *D = lu_fsm;
    double *rD = D + NB * DP;                     // NB·DP = 272 and NB are even: the panels are 16-byte aligned
    double *Lt = rD + NB;
    double *Us = Lt + NB * lsp;
    double *Pin = Us + NB * lsp;                  // staged panel inputs (use_pin): item t at Pin[t·LU_PINP …]
    const bool use_pin = ws.use_pin != 0;         // host guarantees 2·bw ≤ blockDim.x then
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rg = lane >> 3, cg = lane & 7;      // trailing update: 4 row groups × 8 column groups per warp
    bool bad = false;

    for (int k0 = 0; k0 < N; k0 += NB) {
        const int nb = min(NB, N - k0);
        const int R = min(bw, N - k0 - nb);       // rows below / columns right of the block inside the band
        // item t of the panels: t < R → row k0+nb+t of A21, else column k0+nb+(t−R) of A12
        auto panel_load = [&](int t, double (&v)[NB]) {
            if (t < R) {
                const int gr = k0 + nb + t;
                const double *rowp = ab + (size_t)gr * LD + (k0 - gr + bwx);     // entry (gr, k0+c) at rowp[c]; 16-byte aligned
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    const double2 w = *reinterpret_cast<const double2 *>(LU_A16(rowp + c));
                    v[c] = w.x; v[c + 1] = w.y;
                }
            } else {
                const int gj = k0 + nb + (t - R);
                const double *colp = ab + (size_t)k0 * LD + (gj - k0 + bwx);     // entry (k0+r, gj) at colp[r·(LD−1)]
#pragma unroll
                for (int r = 0; r < NB; ++r) v[r] = colp[(size_t)r * (LD - 1)];
            }
        };
        auto panel_solve = [&](int t, double (&v)[NB]) {
            if (t < R) {
                const int gr = k0 + nb + t;
                double *rowp = ab + (size_t)gr * LD + (k0 - gr + bwx);
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    double s = v[c];
#pragma unroll
                    for (int m = 0; m < c; ++m) s -= v[m] * D[m * DP + c];
                    v[c] = s * rD[c];
                }
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    if (!CL) *reinterpret_cast<double2 *>(LU_A16(rowp + c)) = make_double2(v[c], v[c + 1]);
                    Lt[c * lsp + t] = v[c]; Lt[(c + 1) * lsp + t] = v[c + 1];
                }
            } else {
                const int tt = t - R, gj = k0 + nb + tt;
                double *colp = ab + (size_t)k0 * LD + (gj - k0 + bwx);
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    double s = v[r];
#pragma unroll
                    for (int m = 0; m < r; ++m) s -= D[r * DP + m] * v[m];
                    v[r] = s;
                }
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    if (!CL) colp[(size_t)r * (LD - 1)] = v[r];
                    Us[r * lsp + tt] = v[r];
                }
            }
        };
        // the inputs of the thread's panel item start travelling now and arrive under the factorisation of the
        // diagonal block: cp.async into shared memory, so no register is held across it
        if (use_pin && tid < 2 * R) {
            double *dst = Pin + tid * LU_PINP;
            if (tid < R) {
                const int gr = k0 + nb + tid;
                const double *rowp = ab + (size_t)gr * LD + (k0 - gr + bwx);
#pragma unroll
                for (int c = 0; c < NB; c += 2) lu_cp_async16(dst + c, rowp + c);
            } else {
                const double *colp = ab + (size_t)k0 * LD + (nb + (tid - R) + bwx);
#pragma unroll
                for (int r = 0; r < NB; ++r) lu_cp_async8(dst + r, colp + (size_t)r * (LD - 1));
            }
            lu_cp_async_commit();
        }
        // ---- diagonal block: load, factor in warp 0 (identity padding for a short last block) ----
        for (int e = tid; e < NB * NB; e += blockDim.x) {
            const int r = e / NB, c = e - r * NB;
            D[r * DP + c] = (r < nb && c < nb) ? ab[(size_t)(k0 + r) * LD + (c - r + bwx)] : (r == c ? 1.0 : 0.0);
        }
        __syncthreads();
        if (warp == 0 && lu16_warp(D, rD)) bad = true;
        __syncthreads();
        if (!CL) for (int e = tid; e < NB * NB; e += blockDim.x) {   // the factored block back to the band (cluster: after the read barrier below)
            const int r = e / NB, c = e - r * NB;
            if (r < nb && c < nb) ab[(size_t)(k0 + r) * LD + (c - r + bwx)] = D[r * DP + c];
        }
        // ---- L21 = A21 U11⁻¹ (one thread per row), U12 = L11⁻¹ A12 (one thread per column) ----
        if (use_pin) {
            if (tid < 2 * R) {
                lu_cp_async_wait();
                const double *src = Pin + tid * LU_PINP;
                double v[NB];
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    const double2 w = *reinterpret_cast<const double2 *>(LU_A16(src + c));
                    v[c] = w.x; v[c + 1] = w.y;
                }
                panel_solve(tid, v);
            }
        } else {
            for (int t = tid; t < 2 * R; t += blockDim.x) {
                double v[NB];
                panel_load(t, v);
                panel_solve(t, v);
            }
        }
        if (CL) {
            lu_cluster_sync();     // every CTA of the cluster has read this step's block and panels
            // the factored block (rank 0) and panels (items dealt round-robin over the ranks: every CTA holds
            // the complete panels) back to the band, from shared memory
            if (crank == 0)
                for (int e = tid; e < NB * NB; e += blockDim.x) {
                    const int r = e / NB, c = e - r * NB;
                    if (r < nb && c < nb) ab[(size_t)(k0 + r) * LD + (c - r + bwx)] = D[r * DP + c];
                }
            for (int t = tid * csize + crank; t < 2 * R; t += blockDim.x * csize) {
                if (t < R) {
                    const int gr = k0 + nb + t;
                    double *rowp = ab + (size_t)gr * LD + (k0 - gr + bwx);
#pragma unroll
                    for (int c = 0; c < NB; c += 2)
                        *reinterpret_cast<double2 *>(LU_A16(rowp + c)) = make_double2(Lt[c * lsp + t], Lt[(c + 1) * lsp + t]);
                } else {
                    const int tt = t - R, gj = k0 + nb + tt;
                    double *colp = ab + (size_t)k0 * LD + (gj - k0 + bwx);
#pragma unroll
                    for (int r = 0; r < NB; ++r) colp[(size_t)r * (LD - 1)] = Us[r * lsp + tt];
                }
            }
        } else {
            __syncthreads();
        }
        // ---- trailing window A22 −= L21·U12: warp tiles of 32 rows × 32 columns, 8×4 per thread; per rank
        //      the thread reads 8 panel entries of Lt and 4 of Us as six 16-byte shared-memory loads ----
        const int ct = (R + 31) >> 5, ntile = ct * ct;
        const int g0 = k0 + nb;
        for (int wt = crank * nwarps + warp; wt < ntile; wt += csize * nwarps) {
            const int tr = wt / ct, tc = wt - tr * ct;
            const int r0 = tr * 32 + rg * 8, j0 = tc * 32 + cg * 4;
            if (r0 >= R || j0 >= R) continue;
            const bool full = r0 + 8 <= R && j0 + 4 <= R;
            double acc[8][4];
            if (full) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const double *pp = ab + (size_t)(g0 + r0 + a) * LD + (j0 - r0 - a + bwx);
                    const double2 v0 = *reinterpret_cast<const double2 *>(LU_A16(pp));
                    const double2 v1 = *reinterpret_cast<const double2 *>(LU_A16(pp + 2));
                    acc[a][0] = v0.x; acc[a][1] = v0.y; acc[a][2] = v1.x; acc[a][3] = v1.y;
                }
            } else {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int rr = r0 + a;
                    const double *pp = ab + (size_t)(g0 + rr) * LD + (j0 - rr + bwx);
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = (rr < R && j0 + b < R) ? pp[b] : 0.0;
                }
            }
#pragma unroll 4
            for (int c = 0; c < NB; ++c) {
                const double2 u01 = *reinterpret_cast<const double2 *>(LU_A16(Us + c * lsp + j0));
                const double2 u23 = *reinterpret_cast<const double2 *>(LU_A16(Us + c * lsp + j0 + 2));
                double l[8];
#pragma unroll
                for (int a = 0; a < 8; a += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(LU_A16(Lt + c * lsp + r0 + a));
                    l[a] = v.x; l[a + 1] = v.y;
                }
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    acc[a][0] = fma(-l[a], u01.x, acc[a][0]);
                    acc[a][1] = fma(-l[a], u01.y, acc[a][1]);
                    acc[a][2] = fma(-l[a], u23.x, acc[a][2]);
                    acc[a][3] = fma(-l[a], u23.y, acc[a][3]);
                }
            }
            if (full) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    double *pp = ab + (size_t)(g0 + r0 + a) * LD + (j0 - r0 - a + bwx);
                    *reinterpret_cast<double2 *>(LU_A16(pp)) = make_double2(acc[a][0], acc[a][1]);
                    *reinterpret_cast<double2 *>(LU_A16(pp + 2)) = make_double2(acc[a][2], acc[a][3]);
                }
            } else {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int rr = r0 + a;
                    double *pp = ab + (size_t)(g0 + rr) * LD + (j0 - rr + bwx);
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (rr < R && j0 + b < R) pp[b] = acc[a][b];
                }
            }
        }
        if (CL) lu_cluster_sync(); else __syncthreads();
    }
    if (bad && tid == 0 && crank == 0) ws.info[4 * slot] = 1;
}

// ---------------------------------------------------------------------------
// blocked substitution with the factors of lu_factor_kernel; `vec` (N doubles, shared or global memory)
// is solved in place.  D: 2·NB·DP doubles (double buffer), rhs: NB doubles of shared memory.  The factor
// entries a block step needs do not depend on the vector, so they are PREFETCHED one step ahead (the 16 L21
// entries of a thread's row / the 8 U12 entries of a lane's columns into registers, the diagonal block into
// the other half of D): the critical path of a step is the 16-step triangular solve in warp 0 plus one
// register-fed update, not a global-memory round trip.  Ends with a barrier.
// ---------------------------------------------------------------------------
static __device__ void lu_band_solve(const double *ab, int N, int bw, int bwx, int LD, double *vec, double *D,
                                     double *rhs)
{
    constexpr int NB = LU_NB, DP = LU_DP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int nblk = (N + NB - 1) / NB;
    // ---- forward: L y = b, L unit lower ----
    auto fwd_block = [&](int kb, double *Dd) {      // strictly lower part of diagonal block kb
        const int k0 = kb * NB, nb = min(NB, N - k0);
        for (int e = tid; e < NB * NB; e += blockDim.x) {
            const int r = e / NB, c = e - r * NB;
            Dd[r * DP + c] = (r < nb && c < r) ? ab[(size_t)(k0 + r) * LD + (c - r + bwx)] : 0.0;
        }
    };
    auto fwd_row = [&](int kb, double (&l)[NB]) {   // L21 entries of this thread's row below block kb
        const int k0 = kb * NB, nb = min(NB, N - k0), R = min(bw, N - k0 - nb);
        if (tid < R) {
            const int gr = k0 + nb + tid;
            const double *rowp = ab + (size_t)gr * LD + (k0 - gr + bwx);
#pragma unroll
            for (int c = 0; c < NB; c += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(LU_A16(rowp + c));
                l[c] = v.x; l[c + 1] = v.y;
            }
        }
    };
    double lcur[NB], lnext[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) lcur[c] = lnext[c] = 0.0;
    fwd_block(0, D);
    fwd_row(0, lcur);
    __syncthreads();
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB, nb = min(NB, N - k0), R = min(bw, N - k0 - nb);
        double *Dc = D + (kb & 1) * NB * DP;
        if (kb + 1 < nblk) { fwd_block(kb + 1, D + ((kb + 1) & 1) * NB * DP); fwd_row(kb + 1, lnext); }
        if (warp == 0) {
            double y = lane < nb ? vec[k0 + lane] : 0.0;
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                const double ym = __shfl_sync(0xffffffffu, y, m);
                if (lane > m && lane < NB) y -= Dc[lane * DP + m] * ym;
            }
            if (lane < nb) vec[k0 + lane] = y;
        }
        __syncthreads();
        if (tid < R) {                                  // R > 0 implies a full block (nb = NB)
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) s = fma(lcur[c], vec[k0 + c], s);
            vec[k0 + nb + tid] -= s;
        }
        for (int t = tid + blockDim.x; t < R; t += blockDim.x) {   // bands wider than the CTA: not prefetched
            const int gr = k0 + nb + t;
            const double *rowp = ab + (size_t)gr * LD + (k0 - gr + bwx);
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) s = fma(rowp[c], vec[k0 + c], s);
            vec[gr] -= s;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < NB; ++c) lcur[c] = lnext[c];
    }
    // ---- backward: U x = y ----
    constexpr int UM = 8;                               // prefetched U12 entries per lane (columns lane + 32·m)
    auto bwd_block = [&](int kb, double *Dd) {          // upper part of diagonal block kb, identity padding
        const int k0 = kb * NB, nb = min(NB, N - k0);
        for (int e = tid; e < NB * NB; e += blockDim.x) {
            const int r = e / NB, c = e - r * NB;
            double v = (r < nb && c < nb && c >= r) ? ab[(size_t)(k0 + r) * LD + (c - r + bwx)] : (r == c ? 1.0 : 0.0);
            if (r == c) v = 1.0 / v;          // reciprocal pivot, formed off the critical path
            Dd[r * DP + c] = v;
        }
    };
    auto bwd_row = [&](int kb, double (&u)[UM]) {       // U12 entries of row `warp` of block kb (rows ≥ nwarps: not prefetched)
        const int k0 = kb * NB, nb = min(NB, N - k0), R = min(bw, N - k0 - nb);
        if (warp < nb) {
            const double *rowp = ab + (size_t)(k0 + warp) * LD + (nb - warp + bwx);   // entry (k0+r, k0+nb+j) at rowp[j]
#pragma unroll
            for (int m = 0; m < UM; ++m) u[m] = (lane + 32 * m < R) ? rowp[lane + 32 * m] : 0.0;
        }
    };
    double ucur[UM], unext[UM];
#pragma unroll
    for (int m = 0; m < UM; ++m) ucur[m] = unext[m] = 0.0;
    bwd_block(nblk - 1, D + ((nblk - 1) & 1) * NB * DP);
    bwd_row(nblk - 1, ucur);
    __syncthreads();
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int k0 = kb * NB, nb = min(NB, N - k0), R = min(bw, N - k0 - nb);
        double *Dc = D + (kb & 1) * NB * DP;
        if (kb > 0) { bwd_block(kb - 1, D + ((kb - 1) & 1) * NB * DP); bwd_row(kb - 1, unext); }
        for (int r = warp; r < nb; r += nwarps) {
            const double *rowp = ab + (size_t)(k0 + r) * LD + (nb - r + bwx);
            double s = 0.0;
            if (r == warp) {
#pragma unroll
                for (int m = 0; m < UM; ++m)
                    if (lane + 32 * m < R) s = fma(ucur[m], vec[k0 + nb + lane + 32 * m], s);
                for (int j = lane + 32 * UM; j < R; j += 32) s = fma(rowp[j], vec[k0 + nb + j], s);
            } else {
                for (int j = lane; j < R; j += 32) s = fma(rowp[j], vec[k0 + nb + j], s);
            }
            s = lu_warp_sum(s);
            if (lane == 0) rhs[r] = vec[k0 + r] - s;
        }
        __syncthreads();
        if (warp == 0) {
            double x = lane < nb ? rhs[lane] : 0.0;
#pragma unroll
            for (int m = NB - 1; m >= 0; --m) {
                const double xm = __shfl_sync(0xffffffffu, x, m) * Dc[m * DP + m];
                if (lane == m) x = xm;
                if (lane < m) x -= Dc[lane * DP + m] * xm;
            }
            if (lane < nb) vec[k0 + lane] = x;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < UM; ++m) ucur[m] = unext[m];
    }
}

// (M p)(v), matrix-free: p_v + Σ_k α_k(v) (G_kᵀ T_k G_k p)(v)
template <typename Real>
static __device__ __forceinline__ double lu3_apply(const double *pix, int n, int N, int nops, const Real *alpha_maps,
                                                   const Lu3Params &pr, const double *p, int v)
{
    double sk[3] = {0.0, 0.0, 0.0};
    visit_node(v % n, v / n, n, [&](int q, int k, double c1, double c2) {
        if (k >= nops) return;
        double d1, d2;
        op_apply<double>(k, q % n, q / n, n, p, q, d1, d2);
        const double *tk = pix + (size_t)(6 * k) * N;
        const double z1 = tk[q] * d1 + tk[(size_t)N + q] * d2;
        const double z2 = tk[(size_t)2 * N + q] * d1 + tk[(size_t)3 * N + q] * d2;
        sk[k] += c1 * z1 + c2 * z2;
    });
    double s = p[v];
    for (int k = 0; k < nops; ++k) s += lu3_alpha<Real>(alpha_maps, pr, N, k, v) * sk[k];
    return s;
}

// ---------------------------------------------------------------------------
// K4: p = M⁻¹ r with refinement, g_k = p ⊙ G_kᵀ w_k, patch sums (or plain sums for a scalar parameter).
// Dynamic shared memory: D[2·NB·DP] | rhs[NB] | red[40] | vec[N] when vec_in_smem.
// out_img: nops·lm·ln doubles per image, layout [operator][patch] like the m×n×3 array.
// ---------------------------------------------------------------------------
static inline size_t lu_solve_smem(int N, bool vec_in_smem)
{
    return (size_t)(2 * LU_NB * LU_DP + LU_NB + 40 + (vec_in_smem ? N : 0)) * sizeof(double);
}

template <typename Real>
__global__ void __launch_bounds__(LU_THREADS, 1) lu3_solve_kernel(LuSlots ws, Lu3Params pr, const Real *alpha_maps,
                                                                  double *out_img, double *relres_img, int img0,
                                                                  int vec_in_smem)
{
    LU_DYN_SMEM(lu_ssm);
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const double *ab = ws.ab + ws.ab_stride * slot;
    double *pix = ws.pix + ws.pix_stride * slot;
    const double *r = pix + (size_t)LU_PL_R * N;
    double *p = pix + (size_t)LU_PL_P * N, *work = pix + (size_t)LU_PL_WORK * N, *fk = pix + (size_t)LU_PL_F * N;
    double *D = lu_ssm, *rhs = D + 2 * LU_NB * LU_DP, *red = rhs + LU_NB;
    double *vec = vec_in_smem ? red + 40 : work;
    const int tid = threadIdx.x;

    double bn2 = 0.0;
    for (int v = tid; v < N; v += blockDim.x) { const double x = r[v]; vec[v] = x; bn2 = fma(x, x, bn2); }
    bn2 = lu_block_sum(bn2, red);
    lu_band_solve(ab, N, ws.bw, ws.bwx, ws.LD, vec, D, rhs);
    for (int v = tid; v < N; v += blockDim.x) p[v] = vec[v];
    __syncthreads();
    double relres = 0.0, prev = 1e300, rn2_last = 0.0;
    for (int it = 0;; ++it) {
        double rn2 = 0.0;
        for (int v = tid; v < N; v += blockDim.x) {
            const double res = r[v] - lu3_apply<Real>(pix, n, N, ws.nops, alpha_maps, pr, p, v);
            work[v] = res;
            rn2 = fma(res, res, rn2);
        }
        rn2 = lu_block_sum(rn2, red);
        rn2_last = rn2;
        relres = bn2 > 0.0 ? sqrt(rn2 / bn2) : 0.0;
        // stop at rounding level, or when a step no longer gains a factor 4 (the residual itself is only
        // accurate to eps·‖M‖‖p‖: ≈ 1e-10·‖r‖ at γ = 1e8)
        if (it >= pr.refine || relres <= 1e-14 || !(relres <= 0.25 * prev)) break;
        prev = relres;
        if (vec != work) {
            for (int v = tid; v < N; v += blockDim.x) vec[v] = work[v];
            __syncthreads();
        }
        lu_band_solve(ab, N, ws.bw, ws.bwx, ws.LD, vec, D, rhs);
        for (int v = tid; v < N; v += blockDim.x) p[v] += vec[v];
        __syncthreads();
    }
    // a vanished pivot or an unconverged solve must not pass as a gradient: normwise backward error
    // ‖r − Mp‖ / (‖M‖‖p‖ + ‖r‖) with the bound ‖M‖ ≤ 1 + (8 + 8 + 2)·γ·max α
    double pn2 = 0.0, amax = 0.0;
    for (int v = tid; v < N; v += blockDim.x) {
        pn2 = fma(p[v], p[v], pn2);
        for (int k = 0; k < ws.nops; ++k) amax = fmax(amax, lu3_alpha<Real>(alpha_maps, pr, N, k, v));
    }
    pn2 = lu_block_sum(pn2, red);
    amax = lu_block_max(amax, red);
    const double berr = sqrt(rn2_last) / ((1.0 + 18.0 * pr.gamma * amax) * sqrt(pn2) + sqrt(bn2) + 1e-300);
    const bool failed = ws.info[4 * slot] != 0 || !(berr <= 1e-11);
    const int ng = pr.lm * pr.ln;
    for (int k = 0; k < ws.nops; ++k) {
        const double *w1 = pix + (size_t)(6 * k + 4) * N, *w2 = pix + (size_t)(6 * k + 5) * N;
        for (int v = tid; v < N; v += blockDim.x) {
            double s = 0.0;
            visit_node(v % n, v / n, n, [&](int q, int kk, double c1, double c2) {
                if (kk == k) s += c1 * w1[q] + c2 * w2[q];
            });
            fk[v] = p[v] * s;                                   // p ⊙ G_kᵀ w_k (:247-249; p'·G_kᵀw_k :166)
        }
        __syncthreads();
        for (int g = 0; g < ng; ++g) {
            const int pi = g % pr.lm, pj = g / pr.lm;
            double acc = 0.0;
            for (int v = tid; v < N; v += blockDim.x) {
                const int i = v % n, j = v / n;
                const int qi = (int)(((long long)i * pr.lm) / n), qj = (int)(((long long)j * pr.ln) / n);
                if (qi == pi && qj == pj) acc += fk[v];
            }
            acc = lu_block_sum(acc, red);
            if (tid == 0) out_img[((size_t)(img0 + slot) * ws.nops + k) * ng + g] = failed ? nan("") : acc;
        }
        __syncthreads();
    }
    if (tid == 0) relres_img[img0 + slot] = relres;
}

}  // namespace bpltv
