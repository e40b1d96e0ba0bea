// pdps_resident.cuh — kernel B: whole image resident on chip for ALL iterations.
//
// One launch = the complete lower-level solve.  One image per thread-block cluster
// of CS CTAs; CTA `rank` owns the NC = ceil(N/CS) consecutive columns starting at
// rank·NC (a contiguous slab of the column-major image).  Per pixel, x and f (and λ
// for a map) live in REGISTERS of the owning thread for the whole solve; the dual
// field y1, y2 and the over-relaxed x̄ live in SHARED MEMORY planes because their
// row/column neighbours are read by other threads.  The two columns a CTA needs from
// its neighbours — y2(:, c0-1) from the left, x̄(:, c0+NC) from the right — are
// PUSHED by their owners into halo slots of the neighbour's planes through
// distributed shared memory (remote stores, made visible by the cluster barrier), so
// every read in the hot loop is a local shared-memory read.  HBM traffic: f once in,
// u once out.  Two cluster barriers per iteration:
//     phase A  primal prox + over-relaxation  (reads y planes, writes x̄ plane)
//     barrier.cluster
//     phase B  dual ascent + projection       (reads x̄ plane, writes y planes)
//     barrier.cluster
// The path is bounded by barrier latency + shared-memory bandwidth + fp64 issue, not
// by HBM.  Same operation order as the other kernels: strict mode is bit-identical.
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

namespace cg = cooperative_groups;

constexpr int RES_THREADS = 512;

template <typename Real>
struct ResidentArgs {
    const Real *f;
    Real *u_out;
    const Real *alpha_map;
    const StepConsts<Real> *steps;
    int maxiter, M, N, O, init_mode;
    int NC;  // columns per CTA
    Real alpha_s;
    BatchMap<Real> bm;
};

template <typename Real> struct Vec2T;
template <> struct Vec2T<double> { typedef double2 type; };
template <> struct Vec2T<float> { typedef float2 type; };
template <typename Real>
static __device__ __forceinline__ void ld2(const Real *p, Real &a, Real &b)
{
    const typename Vec2T<Real>::type v = *reinterpret_cast<const typename Vec2T<Real>::type *>(p);
    a = v.x; b = v.y;
}
template <typename Real>
static __device__ __forceinline__ void st2(Real *p, Real a, Real b)
{
    typename Vec2T<Real>::type v; v.x = a; v.y = b;
    *reinterpret_cast<typename Vec2T<Real>::type *>(p) = v;
}

// bytes of the planes of one CTA, and of the two halo mbarriers behind them (ASYNC kernels)
template <typename Real>
static __host__ __device__ __forceinline__ size_t resident_plane_bytes(int NC, int M)
{
    return (((size_t)(3 * NC + 2) * M * sizeof(Real)) + 15) & ~(size_t)15;
}

// KC = column slots per thread (each slot = 2 consecutive rows of one column)
// ASYNC: halo columns travel by st.async + mbarrier (above) and the two phases of an iteration are separated by
// CTA barriers; otherwise by plain DSMEM stores and two cluster barriers per iteration (round 1).
template <typename Real, int KC, bool MAP, bool STRICT, bool ASYNC>
__global__ void __launch_bounds__(RES_THREADS, 1) pdps_resident_kernel(const ResidentArgs<Real> a)
{
#ifdef BPLTV_EMU
    unsigned char *smem_raw = reinterpret_cast<unsigned char *>(emu::dyn_smem());
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int o = blockIdx.x / CS;  // image
    const int M = a.M, N = a.N, NC = a.NC;

    // planes: y1[NC][M]; y2[1+NC][M] (slot 0 = left halo); xb[NC+1][M] (slot NC = right halo)
    Real *y1p = reinterpret_cast<Real *>(smem_raw);
    Real *y2p = y1p + (size_t)NC * M;
    Real *xbp = y2p + (size_t)(NC + 1) * M;
    const int plane_total = (3 * NC + 2) * M;
    for (int k = threadIdx.x; k < plane_total; k += blockDim.x) y1p[k] = (Real)0;

    // neighbours' planes through DSMEM
    Real *xb_left = rank > 0 ? cluster.map_shared_rank(xbp, rank - 1) : nullptr;        // their right halo
    Real *y2_right = rank + 1 < CS ? cluster.map_shared_rank(y2p, rank + 1) : nullptr;  // their left halo
    // ASYNC: hbar[0] counts the x̄ halo column arriving from the right neighbour, hbar[1] the y2 halo column from the left
    unsigned long long *hbar = reinterpret_cast<unsigned long long *>(smem_raw + resident_plane_bytes<Real>(NC, M));
    unsigned long long *hbar_left = (ASYNC && rank > 0) ? cluster.map_shared_rank(hbar, rank - 1) : nullptr;       // its [0]
    unsigned long long *hbar_right = (ASYNC && rank + 1 < CS) ? cluster.map_shared_rank(hbar, rank + 1) : nullptr;  // its [1]
    const unsigned colbytes = (unsigned)(M * sizeof(Real));
    if (ASYNC && threadIdx.x == 0) { halo_bar_init(hbar); halo_bar_init(hbar + 1); }

    const int tpc = M >> 1;                 // threads per column (2 rows each)
    const int CG = blockDim.x / tpc;        // column groups
    const int cgrp = threadIdx.x / tpc;
    const int r0 = (threadIdx.x - cgrp * tpc) * 2;
    const bool t_ok = cgrp < CG;            // threads beyond CG·tpc idle (still hit the barriers)
    const int c_begin = rank * NC;
    const size_t img = (size_t)o * M * N;
    const size_t fimg = (size_t)a.bm.f_image(o) * M * N;
    const Real *amap = MAP ? a.alpha_map + (size_t)a.bm.lam_set(o) * a.bm.map_stride : nullptr;
    const Real alpha_s = a.bm.scalar(o, a.alpha_s);

    const int lane = threadIdx.x & 31;
    Real x[KC][2], f[KC][2], al[KC][2];
    bool ok[KC], okw[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = cgrp + CG * k;        // local column
        const int jg = c_begin + c;         // image column
        ok[k] = t_ok && c < NC && jg < N;
        okw[k] = __any_sync(0xffffffffu, ok[k]);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            f[k][v] = 0; x[k][v] = 0; al[k][v] = alpha_s;
            if (ok[k]) {
                const size_t idx = (size_t)jg * M + r0 + v;
                f[k][v] = a.f[fimg + idx];
                if (MAP) al[k][v] = amap[idx];
                x[k][v] = a.init_mode ? f[k][v] : (Real)0;
            }
        }
    }
    cluster.sync();  // planes zeroed (and the halo mbarriers initialised) everywhere before anyone pushes into a halo

    for (int it = 0; it < a.maxiter; ++it) {
        const StepConsts<Real> sc = a.steps[it];
        Real xb[KC][2], y1o[KC][2], y2o[KC][2];
        if (ASYNC) {
            // y2 of the column left of this CTA, sent by the left neighbour in phase B of the previous iteration
            if (hbar_left && it > 0) halo_wait(hbar + 1, it - 1, colbytes);
            // this iteration's phases of both mbarriers (their previous phases are complete: thread 0 waited on them)
            if (threadIdx.x == 0) {
                if (hbar_right) halo_bar_arm(hbar, colbytes);
                if (hbar_left) halo_bar_arm(hbar + 1, colbytes);
            }
        }
        // ---- phase A: x ← prox, x̄ ← over-relaxation ----------------------------------
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (!okw[k]) continue;              // warp-uniform: whole warp in or out (shuffles below)
            const int c = cgrp + CG * k;
            const Real *py1 = y1p + (size_t)c * M + r0;
            const Real *py2 = y2p + (size_t)(c + 1) * M + r0;  // own column (slot c+1), left = slot c
            const bool valid = ok[k];           // lanes of a mixed warp (M < 64) may sit on a dead slot
            Real l0 = 0, l1 = 0;
            y1o[k][0] = y1o[k][1] = y2o[k][0] = y2o[k][1] = 0;
            if (valid) {
                ld2(py1, y1o[k][0], y1o[k][1]); // one 2-element vector access per plane: conflict-free
                ld2(py2, y2o[k][0], y2o[k][1]);
                ld2(py2 - M, l0, l1);
            }
            // row above: the previous lane's second row (same column); lane 0 reads the plane
            Real up = __shfl_up_sync(0xffffffffu, y1o[k][1], 1);
            if (valid && lane == 0 && r0 > 0) up = py1[-1];
            if (r0 == 0) up = (Real)0;
            const Real xn0 = primal_update<Real, STRICT>(x[k][0], f[k][0], up, y1o[k][0], l0, y2o[k][0], sc, xb[k][0]);
            const Real xn1 = primal_update<Real, STRICT>(x[k][1], f[k][1], y1o[k][0], y1o[k][1], l1, y2o[k][1], sc, xb[k][1]);
            x[k][0] = xn0; x[k][1] = xn1;
            if (valid) {
                st2(xbp + (size_t)c * M + r0, xb[k][0], xb[k][1]);
                if (c == 0 && xb_left) {  // first local column: the left CTA needs it as x̄(:, its NC)
                    if (ASYNC) halo_push2(xb_left + (size_t)NC * M + r0, xb[k][0], xb[k][1], hbar_left);
                    else st2(xb_left + (size_t)NC * M + r0, xb[k][0], xb[k][1]);
                }
            }
        }
        if (ASYNC) {
            __syncthreads();                                      // x̄ of this CTA's own columns
            if (hbar_right) halo_wait(hbar, it, colbytes);        // x̄ of the column right of it
        } else {
            cluster.sync();
        }
        // ---- phase B: y ← P_λ(y + σ∇x̄) ---------------------------------------------------
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (!okw[k]) continue;
            const int c = cgrp + CG * k;
            const int jg = c_begin + c;
            const Real *pxb = xbp + (size_t)c * M + r0;
            // row below: the next lane's first row (same column); lane 31 reads the plane
            Real below = __shfl_down_sync(0xffffffffu, xb[k][0], 1);
            const bool valid = ok[k];
            if (valid && lane == 31 && r0 + 2 < M) below = pxb[2];
            Real d1_0, d1_1, d2_0 = 0, d2_1 = 0;
            if (STRICT) {
                d1_0 = StrictOps<Real>::sub(xb[k][1], xb[k][0]);
                d1_1 = (r0 + 2 < M) ? StrictOps<Real>::sub(below, xb[k][1]) : (Real)0;
            } else {
                d1_0 = xb[k][1] - xb[k][0];
                d1_1 = (r0 + 2 < M) ? below - xb[k][1] : (Real)0;
            }
            if (valid && jg + 1 < N) {
                Real rt0, rt1;
                ld2(pxb + M, rt0, rt1);                     // column c+1 (slot NC = pushed halo)
                if (STRICT) { d2_0 = StrictOps<Real>::sub(rt0, xb[k][0]); d2_1 = StrictOps<Real>::sub(rt1, xb[k][1]); }
                else { d2_0 = rt0 - xb[k][0]; d2_1 = rt1 - xb[k][1]; }
            }
            Real v1, v2, w1, w2;
            {   // both rows of the thread in one block: their projection chains interleave
                Real ya[2] = {y1o[k][0], y1o[k][1]}, yb[2] = {y2o[k][0], y2o[k][1]};
                const Real da[2] = {d1_0, d1_1}, db[2] = {d2_0, d2_1}, aa[2] = {al[k][0], al[k][1]};
                dual_update_n<Real, STRICT, 2>(ya, yb, da, db, aa, sc);
                v1 = ya[0]; w1 = ya[1]; v2 = yb[0]; w2 = yb[1];
            }
            Real *py1 = y1p + (size_t)c * M + r0;
            Real *py2 = y2p + (size_t)(c + 1) * M + r0;
            if (valid) {
                st2(py1, v1, w1);
                st2(py2, v2, w2);
                if (c == NC - 1 && y2_right) {  // last local column: the right CTA needs it as y2(:, c0-1)
                    if (ASYNC) halo_push2(y2_right + r0, v2, w2, hbar_right + 1);
                    else st2(y2_right + r0, v2, w2);
                }
            }
        }
        if (ASYNC) __syncthreads();                               // the duals of this CTA's own columns
        else cluster.sync();
    }
    // nothing may still be in flight towards this CTA's shared memory when it exits
    if (ASYNC && hbar_left && a.maxiter > 0) halo_wait(hbar + 1, a.maxiter - 1, colbytes);

#pragma unroll
    for (int k = 0; k < KC; ++k) {
        if (!ok[k]) continue;
        const int jg = c_begin + cgrp + CG * k;
        const size_t idx = img + (size_t)jg * M + r0;
        a.u_out[idx] = x[k][0];
        a.u_out[idx + 1] = x[k][1];
    }
}

// ---------------------------------------------------------------------------
// Kernel B, temporally blocked (T = 2): TWO iterations per halo exchange.
//
// In kernel B above every half-iteration ends in a cluster barrier, because phase A needs y2 of the column to the left
// and phase B needs x̄ of the column to the right — across the CTA boundary.  Here every CTA carries two extra columns
// on each side (E = NC + 4 columns of state: x, f in registers, y1, y2, x̄ in shared-memory planes) and recomputes
// them redundantly: the validity of the redundant columns shrinks by one column per half-iteration,
//     A₁ on [1, E)   B₁ on [1, E-1)   A₂ on [2, E-1)   B₂ on [2, E-2) = the owned columns,
// so two full iterations need only CTA-local barriers.  Then the owned boundary columns (x, y1, y2 of two columns
// per side) are pushed into the neighbours' halo columns through distributed shared memory.  Per two iterations:
// three __syncthreads, one SPLIT cluster barrier (arrive after A₂ — "I have read my halo planes for the last time" —,
// wait after B₂, so its latency hides under B₂) and one full cluster barrier after the pushes; kernel B spends four
// full cluster barriers on the same two iterations.  Every pixel still sees the identical operation sequence, so the
// result is bit-identical to kernel B and to the oracle (tests/test_emu_resident.py, tests/test_gpu_pdps.py).
// ---------------------------------------------------------------------------
constexpr int RES_TB_THREADS = 768;
// CTA size limit by slots per thread: more slots need more registers per thread (65536 / threads)
constexpr int res_tb_max_threads(int KC) { return KC == 1 ? 768 : (KC == 2 ? 640 : 512); }

#ifdef BPLTV_EMU
static inline void res_cluster_arrive() {}
static inline void res_cluster_wait() { emu::cluster_sync(); }
#else
static __device__ __forceinline__ void res_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
static __device__ __forceinline__ void res_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
#endif

template <typename Real, int KC, bool MAP, bool STRICT>
__global__ void __launch_bounds__(res_tb_max_threads(KC), 1) pdps_resident_tb_kernel(const ResidentArgs<Real> a)
{
#ifdef BPLTV_EMU
    unsigned char *smem_raw = reinterpret_cast<unsigned char *>(emu::dyn_smem());
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int o = blockIdx.x / CS;
    const int M = a.M, N = a.N, NC = a.NC, E = NC + 4;

    // planes over the extended columns e = 0..E-1 (image column c_begin - 2 + e), and the staging columns for the
    // x of the four halo columns (0, 1: left halo; 2, 3: right halo)
    Real *y1p = reinterpret_cast<Real *>(smem_raw);
    Real *y2p = y1p + (size_t)E * M;
    Real *xbp = y2p + (size_t)E * M;
    Real *xst = xbp + (size_t)E * M;
    for (int k = threadIdx.x; k < (3 * E + 4) * M; k += blockDim.x) y1p[k] = (Real)0;

    const bool has_left = rank > 0, has_right = rank + 1 < CS;
    Real *L_y1 = has_left ? cluster.map_shared_rank(y1p, rank - 1) : nullptr;
    Real *L_y2 = has_left ? cluster.map_shared_rank(y2p, rank - 1) : nullptr;
    Real *L_xs = has_left ? cluster.map_shared_rank(xst, rank - 1) : nullptr;
    Real *R_y1 = has_right ? cluster.map_shared_rank(y1p, rank + 1) : nullptr;
    Real *R_y2 = has_right ? cluster.map_shared_rank(y2p, rank + 1) : nullptr;
    Real *R_xs = has_right ? cluster.map_shared_rank(xst, rank + 1) : nullptr;

    const int tpc = M >> 1;
    const int CG = blockDim.x / tpc;
    const int cgrp = threadIdx.x / tpc;
    const int r0 = (threadIdx.x - cgrp * tpc) * 2;
    const bool t_ok = cgrp < CG;
    const int c_begin = rank * NC;
    const size_t img = (size_t)o * M * N;
    const size_t fimg = (size_t)a.bm.f_image(o) * M * N;
    const Real *amap = MAP ? a.alpha_map + (size_t)a.bm.lam_set(o) * a.bm.map_stride : nullptr;
    const Real alpha_s = a.bm.scalar(o, a.alpha_s);
    const int lane = threadIdx.x & 31;

    Real x[KC][2], f[KC][2], al[KC][2];
    bool ok[KC], okw[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int e = cgrp + CG * k;
        const int jg = c_begin - 2 + e;
        ok[k] = t_ok && e < E && jg >= 0 && jg < N;
        okw[k] = __any_sync(0xffffffffu, ok[k]);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            f[k][v] = 0; x[k][v] = 0; al[k][v] = alpha_s;
            if (ok[k]) {
                const size_t idx = (size_t)jg * M + r0 + v;
                f[k][v] = a.f[fimg + idx];
                if (MAP) al[k][v] = amap[idx];
                x[k][v] = a.init_mode ? f[k][v] : (Real)0;
            }
        }
    }
    cluster.sync();

    Real xb[KC][2], y1o[KC][2], y2o[KC][2];
    // phase A on the extended columns [lo, hi): x ← prox, x̄ ← over-relaxation (reads the y planes, writes the x̄ plane)
    auto phase_a = [&](const StepConsts<Real> &sc, int lo, int hi) {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (!okw[k]) continue;
            const int e = cgrp + CG * k;
            const bool valid = ok[k] && e >= lo && e < hi;
            const Real *py1 = y1p + (size_t)e * M + r0;
            const Real *py2 = y2p + (size_t)e * M + r0;
            Real l0 = 0, l1 = 0;
            y1o[k][0] = y1o[k][1] = y2o[k][0] = y2o[k][1] = 0;
            if (valid) {
                ld2(py1, y1o[k][0], y1o[k][1]);
                ld2(py2, y2o[k][0], y2o[k][1]);
                if (e > 0) ld2(py2 - M, l0, l1);      // (the plane of an out-of-image column stays zero: column 0 has no left term)
            }
            Real up = __shfl_up_sync(0xffffffffu, y1o[k][1], 1);
            if (valid && lane == 0 && r0 > 0) up = py1[-1];
            if (r0 == 0) up = (Real)0;
            if (valid) {
                const Real xn0 = primal_update<Real, STRICT>(x[k][0], f[k][0], up, y1o[k][0], l0, y2o[k][0], sc, xb[k][0]);
                const Real xn1 = primal_update<Real, STRICT>(x[k][1], f[k][1], y1o[k][0], y1o[k][1], l1, y2o[k][1], sc, xb[k][1]);
                x[k][0] = xn0; x[k][1] = xn1;
                st2(xbp + (size_t)e * M + r0, xb[k][0], xb[k][1]);
            }
        }
    };
    // phase B on [lo, hi): y ← P_λ(y + σ∇x̄) (reads the x̄ plane, writes the y planes)
    auto phase_b = [&](const StepConsts<Real> &sc, int lo, int hi) {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (!okw[k]) continue;
            const int e = cgrp + CG * k;
            const int jg = c_begin - 2 + e;
            const bool valid = ok[k] && e >= lo && e < hi;
            const Real *pxb = xbp + (size_t)e * M + r0;
            Real below = __shfl_down_sync(0xffffffffu, xb[k][0], 1);
            if (valid && lane == 31 && r0 + 2 < M) below = pxb[2];
            if (!valid) continue;
            Real d1_0, d1_1, d2_0 = 0, d2_1 = 0;
            if (STRICT) {
                d1_0 = StrictOps<Real>::sub(xb[k][1], xb[k][0]);
                d1_1 = (r0 + 2 < M) ? StrictOps<Real>::sub(below, xb[k][1]) : (Real)0;
            } else {
                d1_0 = xb[k][1] - xb[k][0];
                d1_1 = (r0 + 2 < M) ? below - xb[k][1] : (Real)0;
            }
            if (jg + 1 < N) {
                Real rt0, rt1;
                ld2(pxb + M, rt0, rt1);
                if (STRICT) { d2_0 = StrictOps<Real>::sub(rt0, xb[k][0]); d2_1 = StrictOps<Real>::sub(rt1, xb[k][1]); }
                else { d2_0 = rt0 - xb[k][0]; d2_1 = rt1 - xb[k][1]; }
            }
            Real v1, v2, w1, w2;
            {   // both rows of the thread in one block: their projection chains interleave
                Real ya[2] = {y1o[k][0], y1o[k][1]}, yb[2] = {y2o[k][0], y2o[k][1]};
                const Real da[2] = {d1_0, d1_1}, db[2] = {d2_0, d2_1}, aa[2] = {al[k][0], al[k][1]};
                dual_update_n<Real, STRICT, 2>(ya, yb, da, db, aa, sc);
                v1 = ya[0]; w1 = ya[1]; v2 = yb[0]; w2 = yb[1];
            }
            st2(y1p + (size_t)e * M + r0, v1, w1);
            st2(y2p + (size_t)e * M + r0, v2, w2);
        }
    };

    for (int it = 0; it < a.maxiter; it += 2) {
        const bool two = it + 1 < a.maxiter;
        const StepConsts<Real> sc1 = a.steps[it];
        phase_a(sc1, 1, E);
        __syncthreads();
        phase_b(sc1, 1, E - 1);
        if (!two) break;                    // odd iteration count: the owned columns [2, E-2) are complete
        __syncthreads();
        const StepConsts<Real> sc2 = a.steps[it + 1];
        phase_a(sc2, 2, E - 1);
        __syncthreads();
        res_cluster_arrive();               // this CTA will not read its halo y planes again in this super-step
        phase_b(sc2, 2, E - 2);
        res_cluster_wait();                 // … and neither will its neighbours read theirs
        if (it + 2 >= a.maxiter) break;
        // refresh the neighbours' halo columns with this CTA's owned boundary columns
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int e = cgrp + CG * k;
            if (!(t_ok && e < E)) continue;
            const Real *py1 = y1p + (size_t)e * M + r0, *py2 = y2p + (size_t)e * M + r0;
            if (has_left && (e == 2 || e == 3) && e < E - 2) {           // their extended columns E-2, E-1
                const int eh = E - 2 + (e - 2);
                st2(L_y1 + (size_t)eh * M + r0, py1[0], py1[1]);
                st2(L_y2 + (size_t)eh * M + r0, py2[0], py2[1]);
                st2(L_xs + (size_t)(2 + (e - 2)) * M + r0, x[k][0], x[k][1]);
            }
            if (has_right && (e == E - 4 || e == E - 3) && e >= 2) {     // their extended columns 0, 1
                const int eh = e - (E - 4);
                st2(R_y1 + (size_t)eh * M + r0, py1[0], py1[1]);
                st2(R_y2 + (size_t)eh * M + r0, py2[0], py2[1]);
                st2(R_xs + (size_t)eh * M + r0, x[k][0], x[k][1]);
            }
        }
        cluster.sync();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int e = cgrp + CG * k;
            if (!ok[k]) continue;
            if (e < 2 && has_left) ld2(xst + (size_t)e * M + r0, x[k][0], x[k][1]);
            else if (e >= E - 2 && has_right) ld2(xst + (size_t)(2 + e - (E - 2)) * M + r0, x[k][0], x[k][1]);
        }
    }

#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int e = cgrp + CG * k;
        if (!ok[k] || e < 2 || e >= E - 2) continue;
        const int jg = c_begin - 2 + e;
        const size_t idx = img + (size_t)jg * M + r0;
        a.u_out[idx] = x[k][0];
        a.u_out[idx + 1] = x[k][1];
    }
}

// ---------------------------------------------------------------------------
// host side: shape → (cluster size, columns per CTA, slots per thread)
// ---------------------------------------------------------------------------
struct ResidentPlan {
    bool ok = false;
    int CS = 0, NC = 0, KC = 0;
    size_t smem = 0;
};

template <typename Real>
static inline ResidentPlan resident_plan(size_t smem_optin, int M, int N, int cs_max = 8)
{
    ResidentPlan p;
    if (M < 2 || (M & 1) || (M / 2) > RES_THREADS || RES_THREADS % (M / 2) != 0) return p;
    const int CG = RES_THREADS / (M / 2);
    const int cs_cands[5] = {16, 8, 4, 2, 1};   // 16 = non-portable cluster size (launch_resident decides)
    for (int ci = 0; ci < 5; ++ci) {
        const int CS = cs_cands[ci];
        if (CS > cs_max) continue;
        if (CS > N) continue;
        const int NC = (N + CS - 1) / CS;
        if ((CS - 1) * NC >= N) continue;              // every rank must own at least one column
        const int KC = (NC + CG - 1) / CG;
        if (KC > 4) continue;
        const size_t smem = resident_plane_bytes<Real>(NC, M) + 16;      // planes + the two halo mbarriers
        if (smem > smem_optin) continue;
        p.ok = true; p.CS = CS; p.NC = NC; p.KC = KC <= 1 ? 1 : (KC <= 2 ? 2 : 4); p.smem = smem;
        return p;
    }
    return p;
}

template <typename Real>
static inline bool resident_eligible(size_t smem_optin, int M, int N) { return resident_plan<Real>(smem_optin, M, N).ok; }

template <typename Real>
static inline int resident_cluster_size(size_t smem_optin, int M, int N) { return resident_plan<Real>(smem_optin, M, N).CS; }

// temporally blocked variant: E = NC + 4 columns per CTA on tpc·ceil(E/KC) threads (≤ RES_TB_THREADS)
struct ResidentTbPlan {
    bool ok = false;
    int CS = 0, NC = 0, KC = 0, threads = 0;
    size_t smem = 0;
};

template <typename Real>
static inline ResidentTbPlan resident_tb_plan(size_t smem_optin, int M, int N, int cs_max)
{
    ResidentTbPlan p;
    if (M < 2 || (M & 1) || (M / 2) > RES_TB_THREADS) return p;
    const int tpc = M / 2;
    const int cs_cands[5] = {16, 8, 4, 2, 1};
    for (int ci = 0; ci < 5; ++ci) {
        const int CS = cs_cands[ci];
        if (CS > cs_max || CS > N) continue;
        const int NC = (N + CS - 1) / CS;
        if (NC < 2 || (CS - 1) * NC >= N) continue;    // halos come from the adjacent CTA only; every rank owns a column
        const int E = NC + 4;
        int KC = 0;
        const int kcs[3] = {1, 2, 4};
        for (int q = 0; q < 3 && !KC; ++q)
            if (((tpc * ((E + kcs[q] - 1) / kcs[q]) + 31) & ~31) <= res_tb_max_threads(kcs[q])) KC = kcs[q];
        if (!KC) continue;
        const size_t smem = (size_t)(3 * E + 4) * M * sizeof(Real);
        if (smem > smem_optin) continue;
        p.ok = true; p.CS = CS; p.NC = NC; p.KC = KC; p.smem = smem;
        p.threads = (tpc * ((E + KC - 1) / KC) + 31) & ~31;
        return p;
    }
    return p;
}

#ifndef BPLTV_EMU      // host-side launch (CUDA runtime)
template <typename Real, int KC>
static inline cudaError_t launch_resident_tb_kc(const ResidentArgs<Real> &a, const ResidentTbPlan &p, bool map, bool strict,
                                                cudaStream_t st)
{
    void (*fn)(const ResidentArgs<Real>);
    if (map) fn = strict ? pdps_resident_tb_kernel<Real, KC, true, true> : pdps_resident_tb_kernel<Real, KC, true, false>;
    else fn = strict ? pdps_resident_tb_kernel<Real, KC, false, true> : pdps_resident_tb_kernel<Real, KC, false, false>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    if (p.CS > 8) {
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.O * p.CS));
    cfg.blockDim = dim3((unsigned)p.threads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (p.CS > 8) {   // all clusters of the batch must be co-resident (≈ one 16-CTA cluster per GPC)
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) != cudaSuccess || nclusters < a.O) {
            cudaGetLastError();
            return cudaErrorLaunchOutOfResources;
        }
    }
    return cudaLaunchKernelEx(&cfg, fn, a);
}

// the temporally blocked kernel with the largest cluster that is co-resident for the whole batch; cudaErrorInvalidValue
// when no plan takes the shape
template <typename Real>
static inline cudaError_t launch_resident_tb(ResidentArgs<Real> a, size_t smem_optin, bool map, bool strict, int cs_cap,
                                             cudaStream_t st)
{
    for (int cs_max = std::min(cs_cap, 16); cs_max >= 1; cs_max = cs_max > 8 ? 8 : 0) {
        const ResidentTbPlan p = resident_tb_plan<Real>(smem_optin, a.M, a.N, cs_max);
        if (!p.ok) return cudaErrorInvalidValue;
        a.NC = p.NC;
        const cudaError_t e = p.KC == 1 ? launch_resident_tb_kc<Real, 1>(a, p, map, strict, st)
                              : p.KC == 2 ? launch_resident_tb_kc<Real, 2>(a, p, map, strict, st)
                                          : launch_resident_tb_kc<Real, 4>(a, p, map, strict, st);
        if (e == cudaSuccess || p.CS <= 8) return e;
        cudaGetLastError();       // the 16-CTA clusters are not co-resident for this batch: the portable size
    }
    return cudaErrorInvalidValue;
}

template <typename Real, int KC>
static inline cudaError_t launch_resident_kc(const ResidentArgs<Real> &a, const ResidentPlan &p, bool map, bool strict,
                                             cudaStream_t st)
{
    void (*fn)(const ResidentArgs<Real>);
    // halo exchange by st.async + mbarrier (default) or by DSMEM stores + two cluster barriers per iteration (BPLTV_RESIDENT_ASYNC=0)
    const char *as_env = bpltv::env_get("BPLTV_RESIDENT_ASYNC");
    const bool async = as_env && *as_env ? atoi(as_env) != 0 : true;
    if (async) {
        if (map) fn = strict ? pdps_resident_kernel<Real, KC, true, true, true> : pdps_resident_kernel<Real, KC, true, false, true>;
        else fn = strict ? pdps_resident_kernel<Real, KC, false, true, true> : pdps_resident_kernel<Real, KC, false, false, true>;
    } else {
        if (map) fn = strict ? pdps_resident_kernel<Real, KC, true, true, false> : pdps_resident_kernel<Real, KC, true, false, false>;
        else fn = strict ? pdps_resident_kernel<Real, KC, false, true, false> : pdps_resident_kernel<Real, KC, false, false, false>;
    }
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    if (p.CS > 8) {
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.O * p.CS));
    cfg.blockDim = dim3(RES_THREADS);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (p.CS > 8) {   // all clusters of the batch must be co-resident (≈ one 16-CTA cluster per GPC)
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) != cudaSuccess || nclusters < a.O) {
            cudaGetLastError();
            return cudaErrorLaunchOutOfResources;
        }
    }
    return cudaLaunchKernelEx(&cfg, fn, a);
}

template <typename Real>
static inline cudaError_t launch_resident(ResidentArgs<Real> a, size_t smem_optin, bool map, bool strict, cudaStream_t st)
{
    // Few images: 16-CTA clusters (non-portable size) halve the columns per CTA — 5000 iterations of one
    // 128×128 image in 8.0 ms instead of 11.9 ms; they fit about one per GPC, so larger batches (10 images:
    // 16.0 vs 11.9 ms) keep the 8-CTA plan.  BPLTV_RESIDENT_CS caps the cluster size.
    const char *cs_env = bpltv::env_get("BPLTV_RESIDENT_CS");
    const int cs_cap = cs_env && *cs_env ? atoi(cs_env) : 16;
    // Two iterations per halo exchange (pdps_resident_tb_kernel).  Measured on B200, 5000 iterations of 128×128 (tools/
    // time_resident.py), kernel B vs blocked: fp32 6.62 → 5.76 ms (1 image), 6.49 → 5.77 (4); fp64 7.98 → 8.52 (1 image),
    // 11.95 → 15.13 (10 images): in fp64 the 50 % of redundant halo columns cost as much fp64 issue as the two saved
    // cluster barriers give back, in fp32 the arithmetic is cheap enough to win.  Hence: fp32 single-wave batches by
    // default; BPLTV_RESIDENT_TB=1 forces it for every precision, =0 disables it.
    const char *tb_env = bpltv::env_get("BPLTV_RESIDENT_TB");
    // Round 2, later: with the halo exchange by st.async + mbarrier kernel B itself went 7.52 → 4.99 ms (fp64) and 6.24 →
    // 3.32 ms (fp32) for one image and beats the blocked variant (8.21 / 5.44 ms) everywhere: the blocked kernel is
    // opt-in only (BPLTV_RESIDENT_TB=1).
    const bool tb_on = tb_env && *tb_env ? atoi(tb_env) != 0 : false;
    if (tb_on && a.maxiter >= 4) {
        const cudaError_t etb = launch_resident_tb<Real>(a, smem_optin, map, strict, cs_cap, st);
        if (etb == cudaSuccess) return etb;
        cudaGetLastError();
    }
    if (cs_cap >= 16) {
        const ResidentPlan p16 = resident_plan<Real>(smem_optin, a.M, a.N, 16);
        if (p16.ok && p16.CS == 16) {
            ResidentArgs<Real> a16 = a;
            a16.NC = p16.NC;
            cudaError_t e16 = p16.KC == 1 ? launch_resident_kc<Real, 1>(a16, p16, map, strict, st)
                              : p16.KC == 2 ? launch_resident_kc<Real, 2>(a16, p16, map, strict, st)
                                            : launch_resident_kc<Real, 4>(a16, p16, map, strict, st);
            if (e16 == cudaSuccess) return e16;
            cudaGetLastError();   // not co-resident (or refused): the portable plan below
        }
    }
    const ResidentPlan p = resident_plan<Real>(smem_optin, a.M, a.N, std::min(cs_cap, 8));
    if (!p.ok) return cudaErrorInvalidValue;
    a.NC = p.NC;
    if (p.KC == 1) return launch_resident_kc<Real, 1>(a, p, map, strict, st);
    if (p.KC == 2) return launch_resident_kc<Real, 2>(a, p, map, strict, st);
    return launch_resident_kc<Real, 4>(a, p, map, strict, st);
}
#endif  // BPLTV_EMU

}  // namespace bpltv
