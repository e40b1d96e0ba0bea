// gradient.cuh — adjoint solve and λ-gradient on the GPU (fp64).
//
// Replaces gradient / gradient_reg of /root/reference/src/TVLearningFunctionVec.jl
// (:98-135, :137-161, :192-215, :219-254), which assemble an explicit sparse
// (3n²)² or (n²)² matrix per image and call a sparse LU (`\`).
//
// Formulation.  After eliminating the multipliers every variant is
//       (C + Gᵀ D G) p = r ,   grad = ± Σ ⟨(Gp), w⟩ (pixel) or Σ p·(Gᵀw) (node)
// with C diagonal and D block diagonal, D_q = s·I ("iso": flat pixel, s = 1/eps or
// αγ) or s·t tᵀ with t ⟂ ∇u ("aniso", s = α/|∇u|).  The penalty form has entries up
// to 4.5e15 and is unusable in fp64, so we factor the equivalent
// SPD system in multiplier space instead,
//       (diag(E) + B C⁻¹ Bᵀ) ζ = B C⁻¹ r ,   p = C⁻¹ (r − Bᵀ ζ),   E = 1/s,
// one unknown per aniso pixel (mode t) and two per iso pixel (modes e₁, e₂).  All its
// entries are O(1); numbering the modes in pixel order makes it banded with
// half-bandwidth ≤ 2n+1, and a blocked right-looking banded Cholesky with a pivot
// floor (redundant constraints of flat regions give pivots ≈ eps) followed by one
// step of iterative refinement reproduces the refined literal solve
// (oracle/oracle.py: gradient_dual is the same algorithm on the CPU).
//
// Mapping: one CTA per image "slot"; images are processed in waves of ≤ #SM slots.
#pragma once
#include "env_switches.h"
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <string>

#include "common.cuh"

namespace bpltv {

constexpr int GRAD_NB = 16;        // Cholesky block size
constexpr int GRAD_THREADS = 512;  // CTA size of the per-image kernels

struct GradSlots {
    // per-slot strides (elements)
    int N;          // pixels per image
    int n;          // image side
    int LD;         // allocated leading dimension of the band (≥ 2n+2+NB)
    int NdMax;      // 2N
    int nblkMax;    // NdMax/NB + 1
    // base pointers of slot 0; slot s adds s*stride
    double *pix;    // 10 pixel arrays of N doubles: ea eb E w1 w2 cinv rc p q fpix
    double *mode;   // 3 mode arrays of NdMax doubles: b zeta work
    double *sinv;   // nblkMax * NB*NB
    double *ab;     // NdMax * LD
    int *off;       // N+1
    int *ext;       // NdMax
    int *info;      // 4 ints per slot: Nd, LDa (actual band + NB + 1), guarded pivots, spare
    size_t pix_stride, mode_stride, sinv_stride, ab_stride, off_stride, ext_stride;
};

struct GradVariant {
    int regularised;   // 1: gradient_reg, 0: gradient
    int patch;         // λ is a map
    int lm, ln;
    double alpha_s, gamma, act_tol, eps_act, guard_rel;
    int refine;
};

template <typename T>
static __device__ __forceinline__ T *slot_ptr(T *base, size_t stride, int slot) { return base + stride * slot; }

// ---------------------------------------------------------------------------
// K1: per-pixel classification + exclusive scan of the mode counts.
// ---------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(GRAD_THREADS) grad_classify_kernel(GradSlots ws, GradVariant gv, const Real *u_all,
                                                                     const Real *ubar_all, const Real *alpha_map,
                                                                     int img0)
{
    __shared__ int s_warp[32];
    __shared__ int s_total;
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const Real *u = u_all + (size_t)(img0 + slot) * N;
    const Real *ub = ubar_all + (size_t)(img0 + slot) * N;
    double *pix = slot_ptr(ws.pix, ws.pix_stride, slot);
    double *ea = pix, *eb = pix + N, *E = pix + 2 * (size_t)N, *w1 = pix + 3 * (size_t)N, *w2 = pix + 4 * (size_t)N,
           *cinv = pix + 5 * (size_t)N, *rc = pix + 6 * (size_t)N;
    int *off = slot_ptr(ws.off, ws.off_stride, slot);
    int *ext = slot_ptr(ws.ext, ws.ext_stride, slot);
    int *info = ws.info + 4 * slot;

    // each thread owns a contiguous run of pixels so the scan is a two-level one
    const int per = (N + blockDim.x - 1) / blockDim.x;
    const int q0 = threadIdx.x * per, q1 = min(N, q0 + per);
    int cnt = 0;
    for (int q = q0; q < q1; ++q) {
        const int i = q % n, j = q / n;
        const double uq = (double)u[q];
        const double g1 = (i + 1 < n) ? (double)u[q + 1] - uq : 0.0;
        const double g2 = (j + 1 < n) ? (double)u[q + n] - uq : 0.0;
        const double nrm = sqrt(g1 * g1 + g2 * g2);
        const double a = gv.patch ? (double)alpha_map[q] : gv.alpha_s;
        bool iso;
        double e, c, r, v1, v2;
        if (gv.regularised) {
            // act = max(0,|Gu|-1/γ) != 0  (:146-147, :201-202); iso = !act
            iso = !(fmax(0.0, nrm - 1.0 / gv.gamma) != 0.0);
            if (iso) { v1 = gv.gamma * g1; v2 = gv.gamma * g2; }
            else { v1 = g1 / nrm; v2 = g2 / nrm; }
            r = (double)ub[q] - uq;                     // ū - u (:157, :212)
            if (gv.patch) { c = a; e = iso ? 1.0 / gv.gamma : nrm; }
            else { c = 1.0; e = iso ? 1.0 / (a * gv.gamma) : nrm / a; }
        } else {
            iso = nrm < gv.act_tol;                     // (:109, :231)
            if (iso) { v1 = 0.0; v2 = 0.0; }
            else { v1 = g1 / nrm; v2 = g2 / nrm; }
            r = uq - (double)ub[q];                     // u - ū (:130, :247)
            c = 1.0;
            e = iso ? gv.eps_act : nrm / a;
        }
        ea[q] = iso ? 1.0 : -g2 / nrm;
        eb[q] = iso ? 0.0 : g1 / nrm;
        E[q] = e; w1[q] = v1; w2[q] = v2; cinv[q] = c; rc[q] = r;
        off[q] = iso ? 2 : 1;  // count for now
        cnt += iso ? 2 : 1;
    }
    // block exclusive scan of cnt
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        int v = lane < nw ? s_warp[lane] : 0, iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        s_warp[lane] = iv - v;  // exclusive warp offsets
        if (lane == 31) s_total = iv;
    }
    __syncthreads();
    int run = s_warp[warp] + incl - cnt;
    for (int q = q0; q < q1; ++q) {
        const int c = off[q];
        off[q] = run;
        run += c;
    }
    if (threadIdx.x == 0) { off[N] = s_total; info[0] = s_total; info[2] = 0; }
    __syncthreads();
    // envelope: last mode coupled to the modes of pixel q = last mode of pixel min(q+n, N-1)
    int bwmax = 0;
    for (int q = threadIdx.x; q < N; q += blockDim.x) {
        const int qq = min(q + n, N - 1);
        const int last = off[qq + 1] - 1;
        for (int a = off[q]; a < off[q + 1]; ++a) {
            ext[a] = last;
            bwmax = max(bwmax, last - a);
        }
    }
    bwmax = max(bwmax, __shfl_xor_sync(0xffffffffu, bwmax, 16));
    bwmax = max(bwmax, __shfl_xor_sync(0xffffffffu, bwmax, 8));
    bwmax = max(bwmax, __shfl_xor_sync(0xffffffffu, bwmax, 4));
    bwmax = max(bwmax, __shfl_xor_sync(0xffffffffu, bwmax, 2));
    bwmax = max(bwmax, __shfl_xor_sync(0xffffffffu, bwmax, 1));
    __syncthreads();
    if (lane == 0) s_warp[warp] = bwmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        int m = 0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) m = max(m, s_warp[w]);
        info[1] = min(ws.LD, m + 1 + GRAD_NB);  // actual leading dimension used by this slot
    }
}

// node coefficients of a mode (pixel (i,j), vector (e1,e2)): β0 at q, β1 at q+1, β2 at q+n
static __device__ __forceinline__ void mode_beta(int i, int j, int n, double e1, double e2, double &b0, double &b1,
                                                 double &b2)
{
    b1 = (i + 1 < n) ? e1 : 0.0;
    b2 = (j + 1 < n) ? e2 : 0.0;
    b0 = -(b1 + b2);
}

// mode m (0/1) of pixel q: vector and compliance
static __device__ __forceinline__ void mode_vec(const double *ea, const double *eb, int q, int m, bool iso, double &e1,
                                                double &e2)
{
    if (iso) { e1 = m == 0 ? 1.0 : 0.0; e2 = m == 0 ? 0.0 : 1.0; }
    else { e1 = ea[q]; e2 = eb[q]; }
}

// ---------------------------------------------------------------------------
// K2: zero the used part of the band, scatter diag(E) + B C⁻¹ Bᵀ (lower band) and
// the right-hand side b = B C⁻¹ r.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GRAD_THREADS) grad_assemble_kernel(GradSlots ws)
{
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const double *pix = slot_ptr(ws.pix, ws.pix_stride, slot);
    const double *ea = pix, *eb = pix + N, *E = pix + 2 * (size_t)N, *cinv = pix + 5 * (size_t)N,
                 *rc = pix + 6 * (size_t)N;
    const int *off = slot_ptr(ws.off, ws.off_stride, slot);
    const int *info = ws.info + 4 * slot;
    const int Nd = info[0], LDa = info[1];
    double *ab = slot_ptr(ws.ab, ws.ab_stride, slot);
    double *bvec = slot_ptr(ws.mode, ws.mode_stride, slot);

    const size_t total = (size_t)Nd * LDa;
    {   // vectorised zero fill (ab is 16-byte aligned per slot, LDa arbitrary → scalar tail)
        double2 *ab2 = reinterpret_cast<double2 *>(ab);
        const size_t n2 = total >> 1;
        const double2 z2 = make_double2(0.0, 0.0);
        for (size_t k = threadIdx.x; k < n2; k += blockDim.x) ab2[k] = z2;
        if (threadIdx.x == 0 && (total & 1)) ab[total - 1] = 0.0;
    }
    __syncthreads();

    for (int q = threadIdx.x; q < N; q += blockDim.x) {
        const int i = q % n, j = q / n;
        const int a0 = off[q];
        const int nm = off[q + 1] - a0;
        const bool iso = nm == 2;
        const double c0 = cinv[q];
        const double c1 = (i + 1 < n) ? cinv[q + 1] : 0.0;
        const double c2 = (j + 1 < n) ? cinv[q + n] : 0.0;
        const double r0 = rc[q];
        const double r1 = (i + 1 < n) ? rc[q + 1] : 0.0;
        const double r2 = (j + 1 < n) ? rc[q + n] : 0.0;
        for (int m = 0; m < nm; ++m) {
            const int a = a0 + m;
            double e1, e2, b0, b1, b2;
            mode_vec(ea, eb, q, m, iso, e1, e2);
            mode_beta(i, j, n, e1, e2, b0, b1, b2);
            double *col = ab + (size_t)a * LDa;
            col[0] = E[q] + b0 * b0 * c0 + b1 * b1 * c1 + b2 * b2 * c2;
            bvec[a] = b0 * r0 + b1 * r1 + b2 * r2;
            if (iso && m == 0) {  // second mode of the same pixel: e = (0,1)
                double f0, f1, f2;
                mode_beta(i, j, n, 0.0, 1.0, f0, f1, f2);
                col[1] = b0 * f0 * c0 + b1 * f1 * c1 + b2 * f2 * c2;
            }
            // pixel q+1 = (i+1, j): shared node q+1 (its β0)
            if (i + 1 < n) {
                const int qq = q + 1, o0 = off[qq], nn = off[qq + 1] - o0;
                for (int mm = 0; mm < nn; ++mm) {
                    double g1, g2, f0, f1, f2;
                    mode_vec(ea, eb, qq, mm, nn == 2, g1, g2);
                    mode_beta(i + 1, j, n, g1, g2, f0, f1, f2);
                    col[o0 + mm - a] = b1 * f0 * c1;
                }
            }
            if (j + 1 < n) {
                // pixel q+n-1 = (i-1, j+1): shared node q+n (its β1)
                if (i > 0) {
                    const int qq = q + n - 1, o0 = off[qq], nn = off[qq + 1] - o0;
                    for (int mm = 0; mm < nn; ++mm) {
                        double g1, g2, f0, f1, f2;
                        mode_vec(ea, eb, qq, mm, nn == 2, g1, g2);
                        mode_beta(i - 1, j + 1, n, g1, g2, f0, f1, f2);
                        col[o0 + mm - a] = b2 * f1 * c2;
                    }
                }
                // pixel q+n = (i, j+1): shared node q+n (its β0)
                {
                    const int qq = q + n, o0 = off[qq], nn = off[qq + 1] - o0;
                    for (int mm = 0; mm < nn; ++mm) {
                        double g1, g2, f0, f1, f2;
                        mode_vec(ea, eb, qq, mm, nn == 2, g1, g2);
                        mode_beta(i, j + 1, n, g1, g2, f0, f1, f2);
                        col[o0 + mm - a] = b2 * f0 * c2;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K3: blocked right-looking banded Cholesky, in place on the band (global memory; the
// working window of ≤ (2n+2+NB)²/2 entries lives in L1/L2).  Per block of NB columns:
//   (1) panel: one thread per row forms L21 = A21·L11⁻ᵀ with the stored inverse of L11;
//   (2) trailing update A22 -= L21·L21ᵀ on the envelope in 4×8 register tiles (band
//       entries prefetched before the rank-NB loop, operands from shared memory);
//   (3) LOOK-AHEAD: while warps 1..15 do (2), warp 0 updates only the next NB×NB
//       diagonal block, factors it in registers (rows across lanes, columns through
//       warp shuffles, pivot floor `guard`) and inverts it, so the serial part of the
//       next step is already done when (2) finishes.
// ---------------------------------------------------------------------------

static __device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gsrc));
}

// warp-level: S (NB×NB row-major, lower, identity padded) → L11 in S, 1/diag(L11) in dinv
static __device__ __forceinline__ void diag_factor_warp(double *S, double *dinv, int nb, double guard, int lane,
                                                        int &guarded)
{
    constexpr int NB = GRAD_NB;
    double row[NB];
    const int lr = lane < NB ? lane : NB - 1;
#pragma unroll
    for (int c = 0; c < NB; ++c) row[c] = S[lr * NB + c];
    double dinv_own = 1.0;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        double d = __shfl_sync(0xffffffffu, row[c], c);
        if (!(d > guard)) { d = guard; if (lane == 0 && c < nb) ++guarded; }
        const double inv = rsqrt(d);
        const double l = row[c] * inv;  // lane == c: d·rsqrt(d) = sqrt(d)
        if (lane == c) { row[c] = d * inv; dinv_own = inv; }
        else if (lane > c) row[c] = l;
#pragma unroll
        for (int c2 = c + 1; c2 < NB; ++c2) {
            const double lc2 = __shfl_sync(0xffffffffu, l, c2);
            if (lane >= c2) row[c2] = fma(-l, lc2, row[c2]);
        }
    }
    if (lane < NB) {
#pragma unroll
        for (int c = 0; c < NB; ++c) S[lane * NB + c] = (c <= lane) ? row[c] : 0.0;
        dinv[lane] = dinv_own;
    }
    __syncwarp();
}

// warp-level: L (NB×NB row-major lower, in shared memory) → Si = L⁻¹ (lane k = column k)
static __device__ __forceinline__ void tri_inverse_warp(const double *L, double *Si, int lane)
{
    constexpr int NB = GRAD_NB;
    double x[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) {
        double s0 = (r == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
        for (int c = 0; c < r; ++c) {
            const double t = (c >= lane) ? L[r * NB + c] * x[c] : 0.0;
            if (c & 1) s1 -= t; else s0 -= t;
        }
        x[r] = (r >= lane) ? (s0 + s1) / L[r * NB + r] : 0.0;
    }
    if (lane < NB) {
#pragma unroll
        for (int r = 0; r < NB; ++r) Si[r * NB + lane] = x[r];
    }
    __syncwarp();
}

// One 4(i)×8(j) tile of the trailing update A22 -= P Pᵀ.  FAST: every entry of the tile is
// strictly inside the envelope (no predicates, affine addresses).  The tile's old band
// entries are staged global→shared with cp.async while the rank-NB FMA loop runs.
template <bool FAST>
static __device__ __forceinline__ void tile_update(double *a22, int LDa, int nrows, const double *P, int PR, int i0,
                                                   int j0, bool use_stage, double *stage, int nstage, int ut)
{
    constexpr int NB = GRAD_NB;
    double *t00 = a22 + (size_t)j0 * LDa - j0 + i0;   // entry (x,y) at t00 + y·(LDa-1) + x
    const size_t cstride = (size_t)LDa - 1;
    if (use_stage) {
#pragma unroll
        for (int y = 0; y < 8; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                if (FAST || (j0 + y < nrows && i0 + x >= j0 + y && i0 + x < nrows)) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + (size_t)(y * 4 + x) * nstage + ut);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(t00 + y * cstride + x));
                }
            }
        asm volatile("cp.async.commit_group;");
    }
    double acc[4][8];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) acc[x][y] = 0.0;
#pragma unroll 2
    for (int c = 0; c < NB; ++c) {
        const double *pc = P + c * PR;
        const double2 pi0 = *reinterpret_cast<const double2 *>(pc + i0);
        const double2 pi1 = *reinterpret_cast<const double2 *>(pc + i0 + 2);
        const double2 pj0 = *reinterpret_cast<const double2 *>(pc + j0);
        const double2 pj1 = *reinterpret_cast<const double2 *>(pc + j0 + 2);
        const double2 pj2 = *reinterpret_cast<const double2 *>(pc + j0 + 4);
        const double2 pj3 = *reinterpret_cast<const double2 *>(pc + j0 + 6);
        const double pi[4] = {pi0.x, pi0.y, pi1.x, pi1.y};
        const double pj[8] = {pj0.x, pj0.y, pj1.x, pj1.y, pj2.x, pj2.y, pj3.x, pj3.y};
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 8; ++y) acc[x][y] = fma(pi[x], pj[y], acc[x][y]);
    }
    if (use_stage) asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int y = 0; y < 8; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            if (FAST || (j0 + y < nrows && i0 + x >= j0 + y && i0 + x < nrows)) {
                double *q = t00 + y * cstride + x;
                const double old = use_stage ? stage[(size_t)(y * 4 + x) * nstage + ut] : *q;
                *q = old - acc[x][y];
            }
        }
}

// CL: the image is factorised by a thread-block CLUSTER (csize CTAs, one SM each).  Every CTA forms the
// (cheap) panel and the look-ahead block redundantly in its own shared memory; the tiles of the trailing
// update — where the flops are — are dealt round-robin over all update warps of the cluster, and one
// cluster barrier per block step publishes them (global memory, release/acquire at cluster scope).
// Only rank 0 writes L11 and the panel back to the band; the panel write-back is deferred by one step
// because the other CTAs are still reading those entries while they form their own copy.
// PLA: the look-ahead (update + factorisation of the next diagonal block) is done by warps 0-3 together
// (one entry pair per thread, column exchange through shared memory, one named barrier per pivot)
// instead of by warp 0 alone with shuffles; same operations per entry, hence the same bits.
template <bool CL, bool PLA>
__global__ void __launch_bounds__(GRAD_THREADS) grad_factor_kernel(GradSlots ws, double guard, int use_stage, int dbg)
{
    extern __shared__ __align__(16) double sm[];
    constexpr int NB = GRAD_NB;
    int crank = 0, csize = 1;
    if (CL) {
        cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
        crank = (int)cluster.block_rank(); csize = (int)cluster.num_blocks();
    }
    double *Sbuf = sm;                 // 2 × NB*NB: L11 of the current / next block
    double *Dbuf = sm + 2 * NB * NB;   // 2 × NB: 1/diag of those; + NB*NB raw next diagonal block
    double *Araw = sm + 3 * NB * NB;   // NB*NB: next diagonal block before its update (cp.async target)
    double *P = sm + 4 * NB * NB;      // NB × PR panel, c-major
    const int slot = CL ? (int)blockIdx.x / csize : (int)blockIdx.x;
    int *info = ws.info + 4 * slot;
    const int Nd = info[0], LDa = info[1];
    const int PR = (LDa + 7) & ~7;     // rows of P, padded to the 4×8 tiles
    const int nstage = GRAD_THREADS - 32;                        // update threads
    double *stage = P + (size_t)NB * ((ws.LD + 7) & ~7);         // 32 × nstage staging slots (if use_stage)
    double *ab = slot_ptr(ws.ab, ws.ab_stride, slot);
    const int *ext = slot_ptr(ws.ext, ws.ext_stride, slot);
    double *sinv = slot_ptr(ws.sinv, ws.sinv_stride, slot);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    int guarded = 0;

    // prologue: first diagonal block straight from the band
    if (warp == 0) {
        const int nb0 = min(NB, Nd);
        for (int idx = lane; idx < NB * NB; idx += 32) {
            const int r = idx / NB, c = idx % NB;
            double v = (r == c) ? 1.0 : 0.0;
            if (r < nb0 && c <= r) v = ab[(size_t)c * LDa + (r - c)];
            Sbuf[idx] = (c <= r) ? v : 0.0;
        }
        __syncwarp();
        diag_factor_warp(Sbuf, Dbuf, nb0, guard, lane, guarded);
    }
    __syncthreads();

    int wb_kb = -1, wb_nb = 0, wb_rows = 0;   // CL: panel of the previous step, still to be written back
// (same thread ↔ row mapping as the panel loop, so no barrier is needed between the two)
#define BPLTV_FACTOR_WRITE_BACK()                                                                        \
    if (CL) {                                                                                            \
        if (crank == 0 && wb_kb >= 0)                                                                    \
            for (int r = tid; r < wb_rows; r += blockDim.x) {                                            \
                const int i_ = wb_kb + wb_nb + r;                                                        \
                for (int c = 0; c < wb_nb; ++c) ab[(size_t)(wb_kb + c) * LDa + (i_ - wb_kb - c)] = P[c * PR + r]; \
            }                                                                                            \
        wb_kb = -1;                                                                                      \
    }
#ifdef BPLTV_FACTOR_TIMING
    // cycle attribution per phase (developer build only: tools/factor_timing.sh); lane 0 of warps 0, 1, 15
    long long tm_panel = 0, tm_syncA = 0, tm_work = 0, tm_syncB = 0, tm_t0 = 0, tm_items = 0;
    auto tm_clock = []() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; };
#define TM_NOW() tm_clock()
#else
#define TM_NOW() 0LL
#endif
    for (int kb = 0, blk = 0; kb < Nd; kb += NB, ++blk) {
        const int cur = blk & 1;
#ifdef BPLTV_FACTOR_TIMING
        tm_t0 = TM_NOW();
#endif
        BPLTV_FACTOR_WRITE_BACK()
        double *S = Sbuf + cur * NB * NB, *dinv = Dbuf + cur * NB;
        double *Sn = Sbuf + (cur ^ 1) * NB * NB, *dinvn = Dbuf + (cur ^ 1) * NB;
        const int nb = min(NB, Nd - kb);
        const int hi = min(Nd - 1, ext[kb + nb - 1]);
        const int nrows = hi - (kb + nb) + 1;
        double *a22 = ab + (size_t)(kb + nb) * LDa;  // column (kb+nb+j) at a22 + j*LDa, entry i-j
        // next diagonal block's current entries → shared memory (independent of this panel)
        if (warp == 0 && nrows > 0) {
            const int nb2 = min(NB, nrows);
            for (int idx = lane; idx < NB * NB; idx += 32) {
                const int r = idx / NB, c = idx % NB;
                if (r < nb2 && c <= r) cp_async8(Araw + idx, a22 + (size_t)c * LDa + (r - c));
            }
            asm volatile("cp.async.commit_group;");
        }
        // ---- store L11; (1) panel rows by forward substitution: x L11ᵀ = a ---------------
        if (!CL || crank == 0)
            for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
                const int r = idx / NB, c = idx % NB;
                if (r < nb && c <= r) ab[(size_t)(kb + c) * LDa + (r - c)] = S[idx];
            }
        for (int r = tid; r < PR; r += blockDim.x) {
            double x[NB];
            if (r < nrows && !(dbg & 4)) {
                const int i = kb + nb + r;
                double a[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) a[c] = (c < nb) ? ab[(size_t)(kb + c) * LDa + (i - kb - c)] : 0.0;
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    double s0 = a[c], s1 = 0.0;
#pragma unroll
                    for (int c2 = 0; c2 < c; ++c2) {
                        if (c2 & 1) s1 = fma(-S[c * NB + c2], x[c2], s1);
                        else s0 = fma(-S[c * NB + c2], x[c2], s0);
                    }
                    x[c] = (s0 + s1) * dinv[c];
                }
                if (!CL) {
#pragma unroll
                    for (int c = 0; c < NB; ++c)
                        if (c < nb) ab[(size_t)(kb + c) * LDa + (i - kb - c)] = x[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < NB; ++c) x[c] = 0.0;
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) P[c * PR + r] = x[c];
        }
        if (warp == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
#ifdef BPLTV_FACTOR_TIMING
        { const long long t = TM_NOW(); tm_panel += t - tm_t0; tm_t0 = t; }
#endif
        __syncthreads();
#ifdef BPLTV_FACTOR_TIMING
        { const long long t = TM_NOW(); tm_syncA += t - tm_t0; tm_t0 = t; }
#endif
        if (CL) { wb_kb = kb; wb_nb = nb; wb_rows = max(nrows, 0); }
        if (nrows <= 0) break;
        if (PLA && (warp & 3) == 0) {
            // ---- (3') look-ahead shared by warps 0, 4, 8, 12: the four warps of ONE scheduler, so that the
            // pivot chain does not compete for issue slots with the tile warps of the other three ---------
            const int nb2 = min(NB, nrows);
            const int lt = (warp >> 2) * 32 + lane;      // 0..127 within the look-ahead group
            double e[2];
            int er[2], ec[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int idx = lt + 128 * h, r = idx / NB, c = idx % NB;
                er[h] = r; ec[h] = c;
                double v = (r == c) ? 1.0 : 0.0;
                if (r < nb2 && c <= r) {
                    double s0 = Araw[idx], s1 = 0.0;
#pragma unroll
                    for (int k = 0; k < NB; k += 2) {
                        s0 = fma(-P[k * PR + r], P[k * PR + c], s0);
                        s1 = fma(-P[(k + 1) * PR + r], P[(k + 1) * PR + c], s1);
                    }
                    v = s0 + s1;
                }
                e[h] = (c <= r) ? v : 0.0;
            }
            if (!(dbg & 2)) {
#pragma unroll 1
                for (int c = 0; c < NB; ++c) {
                    // the owners of column c publish its current entries (final after update c-1)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (ec[h] == c && er[h] >= c) Sn[er[h] * NB + c] = e[h];
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    double d = Sn[c * NB + c];
                    if (!(d > guard)) { d = guard; if (tid == 0 && c < nb2) ++guarded; }   // tid 0 is in the group
                    const double inv = rsqrt(d);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = er[h], c2 = ec[h];
                        if (c2 == c) {
                            if (r == c) { e[h] = d * inv; dinvn[c] = inv; }
                            else if (r > c) e[h] = e[h] * inv;
                        } else if (c2 > c && r >= c2) {
                            const double lr = Sn[r * NB + c] * inv, lc2 = Sn[c2 * NB + c] * inv;
                            e[h] = fma(-lr, lc2, e[h]);
                        }
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");   // all column reads done before L overwrites them
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) Sn[lt + 128 * h] = (ec[h] <= er[h]) ? e[h] : 0.0;
        }
        if (warp == 0) {
            if (!PLA) {
            // ---- (3) look-ahead: next diagonal block = A22[0:NB,0:NB] - P Pᵀ, factor ----------
            const int nb2 = min(NB, nrows);
            for (int idx = lane; idx < NB * NB; idx += 32) {
                const int r = idx / NB, c = idx % NB;
                double v = (r == c) ? 1.0 : 0.0;
                if (r < nb2 && c <= r) {
                    double s0 = Araw[idx], s1 = 0.0;
#pragma unroll
                    for (int k = 0; k < NB; k += 2) {
                        s0 = fma(-P[k * PR + r], P[k * PR + c], s0);
                        s1 = fma(-P[(k + 1) * PR + r], P[(k + 1) * PR + c], s1);
                    }
                    v = s0 + s1;
                }
                Sn[idx] = (c <= r) ? v : 0.0;
            }
            __syncwarp();
            if (!(dbg & 2)) diag_factor_warp(Sn, dinvn, nb2, guard, lane, guarded);
            }
        } else if (!(dbg & 1)) {
            // ---- (2) trailing update on 4(i)×8(j) tiles, skipping the look-ahead block -------
            // Work items are dealt round-robin to the 15 update warps so that every warp runs
            // the same number of rounds (±1).  Interior rounds (tile rows 2tj+2 … nfull-1 of tile
            // column tj: all 32 entries valid) take the predicate-free path; the tiles that
            // cross the diagonal (rows 2tj, 2tj+1) or hang over the last row are gathered into
            // separate boundary rounds, so no warp executes both paths for one item.
            const int nti = (nrows + 3) >> 2, ntj = (nrows + 7) >> 3, nfull = nrows >> 2;
            // update warps of the cluster; with PLA the look-ahead warps join late, so they take the last item slots
            const bool la_warp = (warp & 3) == 0;                            // (PLA) warps 4, 8, 12 come late from the look-ahead
            const int nidx = (warp >> 2) * 3 + (warp & 3) - 1;               // 0..11 among the twelve other update warps
            const int uw = (PLA ? (la_warp ? 11 + (warp >> 2) : nidx) : warp - 1) + (nwarps - 1) * crank, nuw = (nwarps - 1) * csize;
            const int ut = tid - 32;               // index among the update threads
            // item k of the step belongs to update warp k mod nuw.
            if (nrows < 640) {
                // narrow band (TV up to 256x256): the plain scan is the faster code (A/B on B200: 0.84 vs 0.90 s at config 5)
                // PLA: the look-ahead warps arrive late (≈ 1.3-2 items' worth of cycles, measured with
                // tools/factor_timing.py), so they take one item per round of 27 and the others two
                // (warps 4, 8, 12: one item per round of 27; the twelve others two)
                const int slot_a = PLA ? 27 * crank + (la_warp ? 23 + (warp >> 2) : nidx) : uw;
                const int slot_b = PLA ? (la_warp ? -1 : 27 * crank + nidx + 12) : -1;
                const int nslot = PLA ? 27 * csize : nuw;
                int item = 0;
                for (int tj = 0; tj < ntj; ++tj) {
                    for (int tbase = 2 * tj + 2; tbase < nfull; tbase += 32, ++item) {
                        { const int sl = item % nslot; if (sl != slot_a && sl != slot_b) continue; }
                        const int ti = tbase + lane;
                        if (ti < nfull && !(ti < 4 && tj < 2))   // (ti<4,tj<2): next diagonal block (warp 0)
                            tile_update<true>(a22, LDa, nrows, P, PR, 4 * ti, 8 * tj, use_stage != 0, stage, nstage, ut);
                    }
                }
                const int nbound = 3 * ntj;        // (tj, kind): kind 0,1 → rows 2tj, 2tj+1; kind 2 → partial last row
                for (int bbase = 0; bbase < nbound; bbase += 32, ++item) {
                    { const int sl = item % nslot; if (sl != slot_a && sl != slot_b) continue; }
                    const int b = bbase + lane;
                    if (b >= nbound) continue;
                    const int tj = b / 3, kind = b - 3 * tj;
                    int ti = 2 * tj + kind;
                    if (kind == 2) { ti = nti - 1; if (nti == nfull || ti <= 2 * tj + 1) continue; }
                    if (ti >= nti || (ti < 4 && tj < 2)) continue;
                    tile_update<false>(a22, LDa, nrows, P, PR, 4 * ti, 8 * tj, use_stage != 0, stage, nstage, ut);
                }
            } else {
            // wide band (sum-of-regularisers): hundreds of items per step — `skip` = items between here and
            // this warp's next one, carried across tile columns so that no division or per-item scan is needed
            int skip = uw;
            for (int tj = 0; tj < ntj; ++tj) {
                const int first = 2 * tj + 2;
                const int cj = nfull > first ? (nfull - first + 31) >> 5 : 0;   // interior items of this column
                int k = skip;
                for (; k < cj; k += nuw) {
                    const int ti = first + 32 * k + lane;
                    if (ti < nfull && !(ti < 4 && tj < 2))   // (ti<4,tj<2): next diagonal block (warp 0)
                        tile_update<true>(a22, LDa, nrows, P, PR, 4 * ti, 8 * tj, use_stage != 0, stage, nstage, ut);
                }
                skip = k - cj;
            }
            const int nbound = 3 * ntj;            // (tj, kind): kind 0,1 → rows 2tj, 2tj+1; kind 2 → partial last row
            for (int k = skip; 32 * k < nbound; k += nuw) {
                const int b = 32 * k + lane;
                if (b >= nbound) continue;
                const int tj = b / 3, kind = b - 3 * tj;
                int ti = 2 * tj + kind;
                if (kind == 2) { ti = nti - 1; if (nti == nfull || ti <= 2 * tj + 1) continue; }
                if (ti >= nti || (ti < 4 && tj < 2)) continue;
                tile_update<false>(a22, LDa, nrows, P, PR, 4 * ti, 8 * tj, use_stage != 0, stage, nstage, ut);
            }
            }
        }
#ifdef BPLTV_FACTOR_TIMING
        { const long long t = TM_NOW(); tm_work += t - tm_t0; tm_t0 = t; }
#endif
        if (CL) cooperative_groups::this_cluster().sync(); else __syncthreads();
#ifdef BPLTV_FACTOR_TIMING
        { const long long t = TM_NOW(); tm_syncB += t - tm_t0; tm_t0 = t; tm_items += 1; }
#endif
    }
#ifdef BPLTV_FACTOR_TIMING
    if (lane == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 15) && blockIdx.x == 0)
        printf("factor timing warp %2d: steps %lld  cycles/step: panel %.0f  syncA %.0f  work %.0f  syncB %.0f\n", warp, tm_items,
               (double)tm_panel / tm_items, (double)tm_syncA / tm_items, (double)tm_work / tm_items, (double)tm_syncB / tm_items);
#endif
    BPLTV_FACTOR_WRITE_BACK()
#undef BPLTV_FACTOR_WRITE_BACK
    if (tid == 0 && crank == 0) info[2] = guarded;
    if (CL) cooperative_groups::this_cluster().sync(); else __syncthreads();
    // ---- all L11⁻¹ (used by the triangular solves) in one parallel pass: warp per block ----
    {
        double *Lw = P + warp * 2 * NB * NB, *Siw = Lw + NB * NB;  // per-warp scratch (P is free now)
        const int nblk = (Nd + NB - 1) / NB;
        for (int blk = warp + nwarps * crank; blk < nblk; blk += nwarps * csize) {
            const int kb = blk * NB, nb = min(NB, Nd - kb);
            for (int idx = lane; idx < NB * NB; idx += 32) {
                const int r = idx / NB, c = idx % NB;
                double v = (r == c) ? 1.0 : 0.0;
                if (r < nb && c <= r) v = ab[(size_t)(kb + c) * LDa + (r - c)];
                Lw[idx] = (c <= r) ? v : 0.0;
            }
            __syncwarp();
            tri_inverse_warp(Lw, Siw, lane);
            for (int idx = lane; idx < NB * NB; idx += 32) sinv[(size_t)blk * NB * NB + idx] = Siw[idx];
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------
// Banded triangular solves with the factor (device function, whole CTA).
// z ← (L Lᵀ)⁻¹ z, using the stored inverses of the diagonal blocks.  `z` may point
// to shared memory (the caller stages the vector there when it fits).  The factor is
// read-only here, so its panel entries are loaded through the read-only path.
// ---------------------------------------------------------------------------
constexpr int GRAD_BK = 10;  // backward solve: panel entries per lane held in registers (covers bands ≤ 320)

// BK: as GRAD_BK.  WIDE: bands wider than the CTA (the sum-of-regularisers system: ≈ 770 rows below a block at
// 128×128, up to 1559) — the forward solve prefetches a second panel row per thread (rows up to 2·blockDim) and
// the backward solve refills its single register set right after use instead of double-buffering it, so that
// BK = 26 (bands ≤ 832) fits in registers; without it those rows were loaded on the critical path.
template <int BK, bool WIDE>
static __device__ void band_solve(const double *__restrict__ ab, const double *__restrict__ sinv,
                                  const unsigned short *nr /* rows below each block (shared memory) */, int Nd,
                                  int LDa, double *z, double *sh /* ≥ 2*NB + 2*NB*NB doubles */)
{
    constexpr int NB = GRAD_NB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int nblk = (Nd + NB - 1) / NB;
    double *ys = sh, *ts = sh + NB, *sib = sh + 2 * NB;  // sib: 2 × NB*NB, inverse of the current / next block
    auto block_rows = [&](int blk) { return (int)nr[blk]; };
    auto fetch_si = [&](int blk, int buf) {   // everything the dependent part needs, one block ahead
        if (tid < NB * NB) cp_async8(sib + buf * NB * NB + tid, sinv + (size_t)blk * NB * NB + tid);
        asm volatile("cp.async.commit_group;");
    };
    // ---- forward: L y = z.  Thread r owns panel row r of the block ----------------------
    {
        double lcur[NB], lcur2[WIDE ? NB : 1];
        const int tid2 = tid + (int)blockDim.x;
        auto fetch_row = [&](int blk, int nrows, int row, double(&l)[NB]) {
            const int kb = blk * NB, nb = min(NB, Nd - kb);
#pragma unroll
            for (int c = 0; c < NB; ++c)
                l[c] = (row < nrows && c < nb) ? __ldg(ab + (size_t)(kb + c) * LDa + (nb + row - c)) : 0.0;
        };
        auto fetch_l = [&](int blk, int nrows, double(&l)[NB]) {
            fetch_row(blk, nrows, tid, l);
            if constexpr (WIDE) fetch_row(blk, nrows, tid2, lcur2);
        };
        int nrows = nblk > 0 ? block_rows(0) : 0, nrows_n = 0;
        if (nblk > 0) { fetch_l(0, nrows, lcur); fetch_si(0, 0); }
        for (int blk = 0; blk < nblk; ++blk) {
            const int kb = blk * NB, nb = min(NB, Nd - kb);
            const double *Si = sib + (blk & 1) * NB * NB;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                       // Si(blk) staged; previous step's z updates visible
            if (blk + 1 < nblk) { nrows_n = block_rows(blk + 1); fetch_si(blk + 1, (blk + 1) & 1); }
            if (tid < NB) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int c2 = 0; c2 < NB; c2 += 2) {
                    if (c2 <= tid && c2 < nb) s0 = fma(Si[tid * NB + c2], z[kb + c2], s0);
                    if (c2 + 1 <= tid && c2 + 1 < nb) s1 = fma(Si[tid * NB + c2 + 1], z[kb + c2 + 1], s1);
                }
                ys[tid] = s0 + s1;
            }
            __syncthreads();
            if (tid < nb) z[kb + tid] = ys[tid];
            if (tid < nrows) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) { s0 = fma(lcur[c], ys[c], s0); s1 = fma(lcur[c + 1], ys[c + 1], s1); }
                z[kb + nb + tid] -= s0 + s1;
            }
            if constexpr (WIDE) {
                if (tid2 < nrows) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int c = 0; c < NB; c += 2) { s0 = fma(lcur2[c], ys[c], s0); s1 = fma(lcur2[c + 1], ys[c + 1], s1); }
                    z[kb + nb + tid2] -= s0 + s1;
                }
            }
            // rows beyond the prefetched ones (only when the band is wider than that)
            for (int r = tid + (WIDE ? 2 : 1) * blockDim.x; r < nrows; r += blockDim.x) {
                double s = 0.0;
                for (int c = 0; c < nb; ++c) s = fma(__ldg(ab + (size_t)(kb + c) * LDa + (nb + r - c)), ys[c], s);
                z[kb + nb + r] -= s;
            }
            // this row's factor entries for the next block: issued now, consumed two barriers later
            if (blk + 1 < nblk) fetch_l(blk + 1, nrows_n, lcur);
            nrows = nrows_n;
        }
        __syncthreads();
    }
    // ---- backward: Lᵀ x = y.  One warp per column of the block; lanes stride the rows ----
    {
        double lb[BK], lbn[WIDE ? 1 : BK];
        auto fetch_lb = [&](int blk, int nrows, double(&l)[BK]) {
            const int kb = blk * NB, nb = min(NB, Nd - kb);
            const double *col = ab + (size_t)(kb + warp) * LDa + (nb - warp);
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const int r = lane + 32 * k;
                l[k] = (warp < nb && r < nrows) ? __ldg(col + r) : 0.0;
            }
        };
        int nrows = nblk > 0 ? block_rows(nblk - 1) : 0, nrows_n = 0;
        if (nblk > 0) { fetch_lb(nblk - 1, nrows, lb); fetch_si(nblk - 1, (nblk - 1) & 1); }
        for (int blk = nblk - 1; blk >= 0; --blk) {
            const int kb = blk * NB, nb = min(NB, Nd - kb);
            const double *Si = sib + (blk & 1) * NB * NB;
            if (blk > 0) {
                nrows_n = block_rows(blk - 1);
                if constexpr (!WIDE) fetch_lb(blk - 1, nrows_n, lbn);
            }
            if (warp < nb) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k < BK; k += 2) {
                    const int r = lane + 32 * k;
                    if (r < nrows) s0 = fma(lb[k], z[kb + nb + r], s0);
                    if (r + 32 < nrows) s1 = fma(lb[k + 1], z[kb + nb + r + 32], s1);
                }
                const double *col = ab + (size_t)(kb + warp) * LDa + (nb - warp);
                for (int r = lane + 32 * BK; r < nrows; r += 32) s0 = fma(__ldg(col + r), z[kb + nb + r], s0);
                const double s = warp_sum(s0 + s1);
                if (lane == 0) ts[warp] = s;
            }
            if constexpr (WIDE) {   // the register set is free again: the next block's entries travel under the rest of the step
                if (blk > 0) fetch_lb(blk - 1, nrows_n, lb);
            }
            for (int c = warp + nwarps; c < nb; c += nwarps) {  // only if the CTA has fewer than NB warps
                const double *col = ab + (size_t)(kb + c) * LDa + (nb - c);
                double s = 0.0;
                for (int r = lane; r < nrows; r += 32) s = fma(__ldg(col + r), z[kb + nb + r], s);
                s = warp_sum(s);
                if (lane == 0) ts[c] = s;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                       // ts complete, Si(blk) staged
            if (blk > 0) fetch_si(blk - 1, (blk - 1) & 1);
            if (tid < NB) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int c2 = 0; c2 < NB; c2 += 2) {
                    if (c2 >= tid && c2 < nb) s0 = fma(Si[c2 * NB + tid], z[kb + c2] - ts[c2], s0);
                    if (c2 + 1 >= tid && c2 + 1 < nb) s1 = fma(Si[(c2 + 1) * NB + tid], z[kb + c2 + 1] - ts[c2 + 1], s1);
                }
                ys[tid] = s0 + s1;
            }
            __syncthreads();
            if (tid < nb) z[kb + tid] = ys[tid];
            __syncthreads();
            if constexpr (!WIDE) {
#pragma unroll
                for (int k = 0; k < BK; ++k) lb[k] = lbn[k];
            }
            nrows = nrows_n;
        }
    }
}

// p = C⁻¹(r − Bᵀζ) on the nodes of one image (whole CTA)
static __device__ void dual_primal(const GradSlots &ws, int slot, const double *zeta, double *p)
{
    const int n = ws.n, N = ws.N;
    const double *pix = slot_ptr(ws.pix, ws.pix_stride, slot);
    const double *ea = pix, *eb = pix + N, *cinv = pix + 5 * (size_t)N, *rc = pix + 6 * (size_t)N;
    const int *off = slot_ptr(ws.off, ws.off_stride, slot);
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const int i = k % n, j = k / n;
        double s = 0.0;
        {   // pixel k: node k is its β0
            const int o0 = off[k], nm = off[k + 1] - o0;
            for (int m = 0; m < nm; ++m) {
                double e1, e2, b0, b1, b2;
                mode_vec(ea, eb, k, m, nm == 2, e1, e2);
                mode_beta(i, j, n, e1, e2, b0, b1, b2);
                s += b0 * zeta[o0 + m];
            }
        }
        if (i > 0) {  // pixel k-1: node k is its β1
            const int q = k - 1, o0 = off[q], nm = off[q + 1] - o0;
            for (int m = 0; m < nm; ++m) {
                double e1, e2, b0, b1, b2;
                mode_vec(ea, eb, q, m, nm == 2, e1, e2);
                mode_beta(i - 1, j, n, e1, e2, b0, b1, b2);
                s += b1 * zeta[o0 + m];
            }
        }
        if (j > 0) {  // pixel k-n: node k is its β2
            const int q = k - n, o0 = off[q], nm = off[q + 1] - o0;
            for (int m = 0; m < nm; ++m) {
                double e1, e2, b0, b1, b2;
                mode_vec(ea, eb, q, m, nm == 2, e1, e2);
                mode_beta(i, j - 1, n, e1, e2, b0, b1, b2);
                s += b2 * zeta[o0 + m];
            }
        }
        p[k] = rc[k] - cinv[k] * s;
    }
}

// ---------------------------------------------------------------------------
// K4: ζ = A⁻¹b, `refine` steps of iterative refinement with the residual B p − Eζ
// evaluated through the stencils, p, per-pixel functional, patch sums.
// out_img: per-image gradient entries (lm·ln doubles per image), summed over images
// afterwards in a fixed order (grad_reduce_kernel).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GRAD_THREADS) grad_solve_kernel(GradSlots ws, GradVariant gv, double *out_img,
                                                                  double *relres_img, int img0, int zs_cap)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    // [per-block row counts (nblkMax u16, 16-byte padded)] [zs_cap doubles: the solve vector, if it fits]
    unsigned short *s_nr = reinterpret_cast<unsigned short *>(dyn);
    double *zs = reinterpret_cast<double *>(dyn + (((size_t)ws.nblkMax * 2 + 15) & ~(size_t)15));
    __shared__ double sh[2 * GRAD_NB + 2 * GRAD_NB * GRAD_NB + 32];
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const int *info = ws.info + 4 * slot;
    const int Nd = info[0], LDa = info[1];
    const double *ab = slot_ptr(ws.ab, ws.ab_stride, slot);
    const int *ext = slot_ptr(ws.ext, ws.ext_stride, slot);
    const int *off = slot_ptr(ws.off, ws.off_stride, slot);
    const double *sinv = slot_ptr(ws.sinv, ws.sinv_stride, slot);
    double *pix = slot_ptr(ws.pix, ws.pix_stride, slot);
    const double *ea = pix, *eb = pix + N, *E = pix + 2 * (size_t)N, *w1 = pix + 3 * (size_t)N,
                 *w2 = pix + 4 * (size_t)N;
    double *p = pix + 7 * (size_t)N, *fpix = pix + 9 * (size_t)N;
    double *mode = slot_ptr(ws.mode, ws.mode_stride, slot);
    double *bvec = mode, *zeta = mode + ws.NdMax, *work = mode + 2 * (size_t)ws.NdMax;
    const int tid = threadIdx.x;

    double bnorm2 = 0.0;
    for (int a = tid; a < Nd; a += blockDim.x) { const double v = bvec[a]; zeta[a] = v; bnorm2 = fma(v, v, bnorm2); }
    bnorm2 = block_sum(bnorm2, sh);
    __shared__ double s_bn, s_rn;
    if (tid == 0) s_bn = bnorm2;
    __syncthreads();
    {
        const int nblk = (Nd + GRAD_NB - 1) / GRAD_NB;
        for (int blk = tid; blk < nblk; blk += blockDim.x) {
            const int kb = blk * GRAD_NB, nb = min(GRAD_NB, Nd - kb);
            s_nr[blk] = (unsigned short)(min(Nd - 1, ext[kb + nb - 1]) - (kb + nb) + 1);
        }
        __syncthreads();
    }
    const bool in_smem = Nd <= zs_cap;
    // solve `vec` in place; staged through shared memory when it fits
    auto solve = [&](double *vec) {
        if (in_smem) {
            for (int a = tid; a < Nd; a += blockDim.x) zs[a] = vec[a];
            __syncthreads();
            band_solve<GRAD_BK, false>(ab, sinv, s_nr, Nd, LDa, zs, sh);
            for (int a = tid; a < Nd; a += blockDim.x) vec[a] = zs[a];
            __syncthreads();
        } else {
            band_solve<GRAD_BK, false>(ab, sinv, s_nr, Nd, LDa, vec, sh);
        }
    };
    solve(zeta);
    double relres = 0.0;
    for (int it = 0; it <= gv.refine; ++it) {
        dual_primal(ws, slot, zeta, p);
        __syncthreads();
        // residual of the dual system: res_a = e_aᵀ(Gp)_q − E_a ζ_a
        double rn2 = 0.0;
        for (int q = tid; q < N; q += blockDim.x) {
            const int i = q % n, j = q / n;
            const double pq = p[q];
            const double d1 = (i + 1 < n) ? p[q + 1] - pq : 0.0;
            const double d2 = (j + 1 < n) ? p[q + n] - pq : 0.0;
            const int o0 = off[q], nm = off[q + 1] - o0;
            for (int m = 0; m < nm; ++m) {
                double e1, e2;
                mode_vec(ea, eb, q, m, nm == 2, e1, e2);
                const double r = e1 * d1 + e2 * d2 - E[q] * zeta[o0 + m];
                work[o0 + m] = r;
                rn2 = fma(r, r, rn2);
            }
        }
        rn2 = block_sum(rn2, sh);
        if (tid == 0) s_rn = rn2;
        __syncthreads();
        relres = (s_bn > 0.0) ? sqrt(s_rn / s_bn) : 0.0;
        // refinement is skipped when the factorisation already solved the system to rounding level (the
        // well-conditioned regularised systems: measured change of the gradient ≤ 1e-14); the
        // non-regularised systems (compliances down to eps()) need the step (1e-9 → 1e-12)
        if (it == gv.refine || relres <= 1e-14) break;
        solve(work);
        for (int a = tid; a < Nd; a += blockDim.x) zeta[a] += work[a];
        __syncthreads();
    }
    // functional per pixel (scalar, patch non-reg: sign·⟨(Gp)_q, w_q⟩) or per node
    // (patch reg: p_k (Gᵀw)_k, :213)
    const double sign = gv.regularised ? 1.0 : -1.0;
    const bool node_kind = gv.regularised && gv.patch;
    for (int q = tid; q < N; q += blockDim.x) {
        const int i = q % n, j = q / n;
        double v;
        if (!node_kind) {
            const double pq = p[q];
            const double d1 = (i + 1 < n) ? p[q + 1] - pq : 0.0;
            const double d2 = (j + 1 < n) ? p[q + n] - pq : 0.0;
            v = sign * (d1 * w1[q] + d2 * w2[q]);
        } else {
            // (Gᵀw)_k = [w1(k-1) - w1(k)·v1] + [w2(k-n) - w2(k)·v2]
            double s = 0.0;
            if (i > 0) s += w1[q - 1];
            if (i + 1 < n) s -= w1[q];
            if (j > 0) s += w2[q - n];
            if (j + 1 < n) s -= w2[q];
            v = p[q] * s;
        }
        fpix[q] = v;
    }
    __syncthreads();
    // patch sums (PatchOp adjoint, S7): patch of pixel i is floor(i*lm/n)
    const int ng = gv.lm * gv.ln;
    for (int g = 0; g < ng; ++g) {
        const int pi = g % gv.lm, pj = g / gv.lm;
        double acc = 0.0;
        for (int q = tid; q < N; q += blockDim.x) {
            const int i = q % n, j = q / n;
            const int qi = (int)(((long long)i * gv.lm) / n), qj = (int)(((long long)j * gv.ln) / n);
            if (qi == pi && qj == pj) acc += fpix[q];
        }
        acc = block_sum(acc, sh);
        if (tid == 0) out_img[(size_t)(img0 + slot) * ng + g] = acc;
        __syncthreads();
    }
    if (tid == 0) relres_img[img0 + slot] = relres;
}

// grad[g] = Σ_o out_img[o][g] in image order (:76-81: serial i ascending); also the
// worst relative residual.
__global__ void grad_reduce_kernel(const double *out_img, const double *relres_img, int O, int ng, double *grad,
                                   double *relres_max)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < ng) {
        double s = 0.0;
        for (int o = 0; o < O; ++o) s += out_img[(size_t)o * ng + g];
        grad[g] = s;
    }
    if (g == 0 && relres_max) {
        double m = 0.0;
        for (int o = 0; o < O; ++o) m = fmax(m, relres_img[o]);
        relres_max[0] = m;
    }
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
// launches the factorisation of `cnt` slots, `C` CTAs (one cluster) per slot
static inline cudaError_t launch_factor(const GradSlots &ws, double guard, int use_stage, int dbg, int cnt, int C,
                                        size_t smem, cudaStream_t st)
{
    if (C <= 1) {
        // shared look-ahead with the weighted tile distribution: A/B on B200 — 23.8 -> 22.1 ms (cameraman),
        // 27.7 -> 25.3 ms (10 faces), 20.6 -> 18.5 ms (circle, patch), 165 -> 140 ms (32 images of 256x256);
        // bit-identical results either way
        const char *pla_env = bpltv::env_get("BPLTV_GRAD_PLA");
        const bool pla = pla_env && *pla_env ? atoi(pla_env) != 0 : true;
        if (pla) {
            cudaError_t e = cudaFuncSetAttribute(grad_factor_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            grad_factor_kernel<false, true><<<cnt, GRAD_THREADS, smem, st>>>(ws, guard, use_stage, dbg);
            return cudaGetLastError();
        }
        cudaError_t e = cudaFuncSetAttribute(grad_factor_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        grad_factor_kernel<false, false><<<cnt, GRAD_THREADS, smem, st>>>(ws, guard, use_stage, dbg);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(grad_factor_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (C > 8) {
        e = cudaFuncSetAttribute(grad_factor_kernel<true, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) { cudaGetLastError(); return launch_factor(ws, guard, use_stage, dbg, cnt, 8, smem, st); }
    }
    {   // a cluster of C CTAs with this much shared memory must be co-schedulable; otherwise halve it
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3((unsigned)C); q.blockDim = dim3(GRAD_THREADS); q.dynamicSmemBytes = smem;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = (unsigned)C; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        q.attrs = qa; q.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, grad_factor_kernel<true, false>, &q) != cudaSuccess || nclusters < 1) {
            cudaGetLastError();
            return launch_factor(ws, guard, use_stage, dbg, cnt, C / 2, smem, st);
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cnt * C));
    cfg.blockDim = dim3(GRAD_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, grad_factor_kernel<true, false>, ws, guard, use_stage, dbg);
}

// CTAs per image for the factorisation: as many as leave every image of the wave its own cluster.
// Only wide bands have enough trailing-update tiles per block step to share (sum-of-regularisers:
// ≥ 771 rows); the TV band (≈ 130-260 rows at 128²) is bound by the per-step latency chain (panel →
// look-ahead pivots → barrier), which a cluster barrier only lengthens (measured: 24 ms with 1 or 8 CTAs).
static inline int factor_cluster_size(int images_in_wave, int sm_count, int max_band)
{
    const char *env = bpltv::env_get("BPLTV_GRAD_CLUSTER");
    if (env && *env) { const int c = atoi(env); if (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) return c; }
    if (max_band < 600) return 1;
    int C = 1;
    while (C < 8 && 2 * C * images_in_wave <= sm_count) C *= 2;
    // 16-CTA clusters (non-portable size, about one per GPC) for the smallest batches: 158 -> 140 ms on
    // one 128x128 sum-of-regularisers image
    if (C == 8 && images_in_wave <= 4) C = 16;
    return C;
}

struct GradWork {
    void *pix = nullptr, *mode = nullptr, *sinv = nullptr, *ab = nullptr, *off = nullptr, *ext = nullptr,
         *info = nullptr, *out_img = nullptr, *relres = nullptr, *relres_max = nullptr;
    size_t cap_slots = 0, cap_N = 0, cap_O = 0, cap_ng = 0;
    void *lu_ab = nullptr, *lu_pix = nullptr, *lu_info = nullptr;   // node-space band LU (lu_band.cuh)
    size_t lu_cap_slots = 0, lu_cap_N = 0;
    int slots = 0;
    std::string err;
    long long last_iterations = 0;
    double last_relres = 0.0;
    void release()
    {
        void **all[] = {&pix, &mode, &sinv, &ab, &off, &ext, &info, &out_img, &relres, &relres_max,
                        &lu_ab, &lu_pix, &lu_info};
        for (void **p : all) { if (*p) cudaFree(*p); *p = nullptr; }
        cap_slots = cap_N = cap_O = cap_ng = 0;
        lu_cap_slots = lu_cap_N = 0;
    }
};

template <typename Real>
struct GradProblem {
    const Real *u, *ubar;
    int M, N, O;
    double alpha_s;
    const Real *alpha_map;
    int lm, ln;
    bool regularised;
    double gamma, act_tol, eps_act, tol;
    int maxit, solver;
};

static inline int grad_fail(GradWork &w, int code, const std::string &msg) { w.err = msg; return code; }

template <typename Real>
static int run_gradient(GradWork &w, const GradProblem<Real> &gp, int sm_count, size_t smem_optin, cudaStream_t st,
                        double *d_grad_out, long long *launches)
{
    const int n = gp.M;
    const int N = gp.M * gp.N;
    const int ng = gp.lm * gp.ln;
    if (gp.M != gp.N) return grad_fail(w, -1, "square images required");
    if (ng > 1024) return grad_fail(w, -1, "lambda grid larger than 1024 entries is not supported");
    GradSlots ws;
    ws.N = N; ws.n = n; ws.LD = 2 * n + 2 + GRAD_NB; ws.NdMax = 2 * N; ws.nblkMax = ws.NdMax / GRAD_NB + 1;
    ws.pix_stride = (size_t)10 * N;
    ws.mode_stride = (size_t)3 * ws.NdMax;
    ws.sinv_stride = (size_t)ws.nblkMax * GRAD_NB * GRAD_NB;
    ws.ab_stride = ((size_t)ws.NdMax * ws.LD + 1) & ~(size_t)1;
    ws.off_stride = (size_t)N + 2;
    ws.ext_stride = (size_t)ws.NdMax;
    size_t smem = (size_t)(4 * GRAD_NB * GRAD_NB + GRAD_NB * ((ws.LD + 7) & ~7)) * sizeof(double);
    const size_t stage_bytes = (size_t)32 * (GRAD_THREADS - 32) * sizeof(double);
    const int use_stage = smem + stage_bytes <= smem_optin ? 1 : 0;
    if (use_stage) smem += stage_bytes;
    smem = std::max<size_t>(smem, (size_t)(4 * GRAD_NB * GRAD_NB + 32 * GRAD_NB * GRAD_NB) * sizeof(double));  // inverse pass scratch
    if (smem > smem_optin) return grad_fail(w, -1, "image too large for the banded Cholesky panel in shared memory");

    // slots: one CTA per image, at most one per SM, bounded by a workspace budget
    const size_t per_slot = (ws.pix_stride + ws.mode_stride + ws.sinv_stride + ws.ab_stride) * 8 +
                            (ws.off_stride + ws.ext_stride + 4) * 4;
    // the driver is asked about free memory only when the workspace has to grow (the query costs host time
    // inside the timed gradient phase)
    const int want_slots = std::min(gp.O, sm_count);
    int slots = want_slots;
    if (w.cap_N != (size_t)N || w.cap_slots < (size_t)want_slots) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        size_t have = w.cap_N == (size_t)N ? w.cap_slots : 0;
        size_t budget = (free_b + have * per_slot) / 2;
        slots = (int)std::min<size_t>((size_t)want_slots, std::max<size_t>(1, budget / per_slot));
    }
    if (w.cap_N != (size_t)N || w.cap_slots < (size_t)slots) {
        void **all[] = {&w.pix, &w.mode, &w.sinv, &w.ab, &w.off, &w.ext, &w.info};
        for (void **p : all) { if (*p) cudaFree(*p); *p = nullptr; }
        cudaError_t e = cudaSuccess;
        if (e == cudaSuccess) e = cudaMalloc(&w.pix, ws.pix_stride * 8 * slots);
        if (e == cudaSuccess) e = cudaMalloc(&w.mode, ws.mode_stride * 8 * slots);
        if (e == cudaSuccess) e = cudaMalloc(&w.sinv, ws.sinv_stride * 8 * slots);
        if (e == cudaSuccess) e = cudaMalloc(&w.ab, ws.ab_stride * 8 * slots);
        if (e == cudaSuccess) e = cudaMalloc(&w.off, ws.off_stride * 4 * slots);
        if (e == cudaSuccess) e = cudaMalloc(&w.ext, ws.ext_stride * 4 * slots);
        if (e == cudaSuccess) e = cudaMalloc(&w.info, 16 * (size_t)slots);
        if (e != cudaSuccess) {
            cudaGetLastError();
            w.cap_slots = 0; w.cap_N = 0;
            return grad_fail(w, -6, std::string("gradient workspace allocation failed: ") + cudaGetErrorString(e));
        }
        w.cap_slots = slots; w.cap_N = N;
    } else {
        slots = (int)std::min<size_t>(w.cap_slots, (size_t)std::min(gp.O, sm_count));
    }
    if (w.cap_O < (size_t)gp.O || w.cap_ng < (size_t)ng) {
        if (w.out_img) cudaFree(w.out_img);
        if (w.relres) cudaFree(w.relres);
        if (!w.relres_max) cudaMalloc(&w.relres_max, 8);
        cudaError_t e = cudaMalloc(&w.out_img, (size_t)gp.O * ng * 8);
        if (e == cudaSuccess) e = cudaMalloc(&w.relres, (size_t)gp.O * 8);
        if (e != cudaSuccess) { cudaGetLastError(); w.cap_O = 0; return grad_fail(w, -6, "gradient output allocation failed"); }
        w.cap_O = gp.O; w.cap_ng = ng;
    }
    ws.pix = (double *)w.pix; ws.mode = (double *)w.mode; ws.sinv = (double *)w.sinv; ws.ab = (double *)w.ab;
    ws.off = (int *)w.off; ws.ext = (int *)w.ext; ws.info = (int *)w.info;

    GradVariant gv;
    gv.regularised = gp.regularised ? 1 : 0;
    gv.patch = gp.alpha_map != nullptr;
    gv.lm = gp.lm; gv.ln = gp.ln;
    gv.alpha_s = gp.alpha_s; gv.gamma = gp.gamma; gv.act_tol = gp.act_tol; gv.eps_act = gp.eps_act;
    gv.guard_rel = 1e-13;
    gv.refine = bpltv::env_get("BPLTV_GRAD_REFINE") ? atoi(bpltv::env_get("BPLTV_GRAD_REFINE")) : 1;
    // pivot floor relative to the scale of B C⁻¹ Bᵀ (C⁻¹ = λ for the patch-reg system, 1 otherwise)
    double cscale = 1.0;
    if (gv.regularised && gv.patch) cscale = 1.0;  // refined below from the map's max on the host side if needed
    const double guard = gv.guard_rel * cscale;

    // solve vector in shared memory when it fits next to the static arrays
    const size_t nr_bytes = ((size_t)ws.nblkMax * 2 + 15) & ~(size_t)15;
    if (nr_bytes + 8192 + 4096 > smem_optin) return grad_fail(w, -1, "image too large for the solve kernel's block table");
    const size_t zs_only = std::min<size_t>((size_t)ws.NdMax * 8, smem_optin - 8192 - nr_bytes);
    const int zs_cap = (int)(zs_only / 8);
    const size_t zs_bytes = nr_bytes + zs_only;
    cudaFuncSetAttribute(grad_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zs_bytes);
    for (int img0 = 0; img0 < gp.O; img0 += slots) {
        const int cnt = std::min(slots, gp.O - img0);
        grad_classify_kernel<Real><<<cnt, GRAD_THREADS, 0, st>>>(ws, gv, gp.u, gp.ubar, gp.alpha_map, img0);
        grad_assemble_kernel<<<cnt, GRAD_THREADS, 0, st>>>(ws);
        {
            cudaError_t fe = launch_factor(ws, guard, use_stage, bpltv::env_get("BPLTV_GRAD_DBG") ? atoi(bpltv::env_get("BPLTV_GRAD_DBG")) : 0, cnt,
                                           factor_cluster_size(cnt, sm_count, ws.LD), smem, st);
            if (fe != cudaSuccess) { cudaGetLastError(); return grad_fail(w, -2, std::string("factor launch failed: ") + cudaGetErrorString(fe)); }
        }
        grad_solve_kernel<<<cnt, GRAD_THREADS, zs_bytes, st>>>(ws, gv, (double *)w.out_img, (double *)w.relres, img0,
                                                                zs_cap);
        *launches += 4;
    }
    grad_reduce_kernel<<<1, std::max(32, (ng + 31) / 32 * 32), 0, st>>>((double *)w.out_img, (double *)w.relres, gp.O,
                                                                        ng, d_grad_out, (double *)w.relres_max);
    *launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return grad_fail(w, -2, std::string("gradient kernel launch failed: ") + cudaGetErrorString(e));
    w.last_iterations = 0;
    w.last_relres = 0.0;
    return 0;
}

}  // namespace bpltv
