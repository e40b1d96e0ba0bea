// pdps_march.cuh — kernel A: HBM-streaming fused PDPS iteration ("column march").
//
// One launch = one primal-dual iteration over the whole M×N×O stack.  The
// forward-difference gradient, the dual ascent, the pointwise projection onto
// the λ-ball, the divergence, the primal prox and the over-relaxation are fused;
// x̄, Δx, Δy never touch memory.  Algorithmic traffic: read {x, y1, y2, f},
// write {x, y1, y2} = 7 words per pixel-iteration (8 with a λ-map).
//
// Work decomposition: the N·O image columns of the stack form one global column
// index space that is cut into gridDim.x equal contiguous ranges (±1 column), one
// per CTA; the grid is exactly (#SM × resident CTAs per SM), so the launch is a
// single balanced wave.  A CTA spans the whole column height (blockDim.x·VEC ≥ M
// rows, VEC consecutive rows per thread → 16-byte coalesced accesses along Julia's
// fastest axis) and marches along j through its range, one image segment at a time:
// at column c it loads (x,f,y1,y2)(:,c), forms x_new and x̄ for that column, then
// finishes the dual update of column c-1 (which needed x̄(:,c)).  Column c-1's old
// duals and x̄ ride in registers, row neighbours come from warp shuffles (one
// shared-memory slot per warp boundary), so the only redundant global reads are
// y2(:,c0-1) and the look-ahead column at the two ends of a segment.  Columns ahead
// are pulled into L2 with prefetch.global.L2 (no registers held), which keeps the
// register count low enough for ≥ 1024 resident threads per SM.  Input and output
// state are separate buffers (ping-pong), so ranges are independent within a launch.
#pragma once
#include "common.cuh"

namespace bpltv {

template <typename Real>
struct MarchArgs {
    const Real *x_in, *y1_in, *y2_in, *f;
    Real *x_out, *y1_out, *y2_out;
    const Real *alpha_map;            // M×N (shared by all images) or nullptr
    StepConsts<Real> sc;              // this iteration's constants (kernel-parameter constant bank)
    int M, N, O;
    long long total_cols;             // N·O
    int prefetch_dist;                // columns of L2 look-ahead (0 = off)
    Real alpha_s, rho;
    BatchMap<Real> bm;
};

static __device__ __forceinline__ void prefetch_l2(const void *p)
{
#ifndef BPLTV_EMU      // a hint only; the thread emulation of tests/emu has no L2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

template <typename Real, int VEC, bool MAP, bool STRICT, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) pdps_march_kernel(const MarchArgs<Real> a)
{
    typedef VecIO<Real, VEC> IO;
    __shared__ Real s_xb[2][33];  // x̄ of the first row of each warp, double-buffered by column parity

    const int M = a.M, N = a.N;
    const int r0 = threadIdx.x * VEC;
    const bool rows_ok = r0 < M;          // M % VEC == 0 → all VEC rows valid together
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool multi_warp = blockDim.x > 32;
    const bool has_rho = a.rho != (Real)0;
    const StepConsts<Real> &sc = a.sc;

    // this CTA's range of global columns g = o·N + j
    const long long per = (a.total_cols + gridDim.x - 1) / gridDim.x;
    long long g = (long long)blockIdx.x * per;
    const long long g_end = min(a.total_cols, g + per);

    while (g < g_end) {
        // segment = the part of the range inside one image
        const int o = (int)(g / N);
        const int c0 = (int)(g - (long long)o * N);
        const int c1 = (int)min((long long)N, (long long)c0 + (g_end - g));  // exclusive
        g += c1 - c0;
        const size_t img = (size_t)o * M * N;
        const Real *xin = a.x_in + img, *y1in = a.y1_in + img, *y2in = a.y2_in + img;
        const Real *fin = a.f + (size_t)a.bm.f_image(o) * M * N;
        const Real *amap = MAP ? a.alpha_map + (size_t)a.bm.lam_set(o) * a.bm.map_stride : nullptr;
        const Real alpha_s = a.bm.scalar(o, a.alpha_s);
        Real *xout = a.x_out + img, *y1out = a.y1_out + img, *y2out = a.y2_out + img;

        // state of the previous column (c-1): x̄, Δy1 = x̄(i+1)-x̄(i), old duals, λ
        Real xb_p[VEC], d1_p[VEC], y1_p[VEC], y2_p[VEC], al_p[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { xb_p[v] = d1_p[v] = y1_p[v] = y2_p[v] = 0; al_p[v] = alpha_s; }
        // y2 of column c0-1 (zero left of the image)
        if (rows_ok && c0 > 0) IO::ld(y2in + (size_t)(c0 - 1) * M + r0, y2_p);

        const int c_last = min(c1, N - 1);  // last column whose x̄ we need (look-ahead column if c1 < N)

        for (int c = c0; c <= c_last; ++c) {
            Real x_c[VEC], f_c[VEC], y1_c[VEC], y2_c[VEC], al_c[VEC], up_c = 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) { x_c[v] = f_c[v] = y1_c[v] = y2_c[v] = 0; al_c[v] = alpha_s; }
            if (rows_ok) {
                const size_t off = (size_t)c * M + r0;
                IO::ld(xin + off, x_c);
                IO::ld(fin + off, f_c);
                IO::ld(y1in + off, y1_c);
                IO::ld(y2in + off, y2_c);
                if (MAP) IO::ld(amap + off, al_c);
                // y1 of the row above this warp's first row belongs to another warp: read it from
                // memory (it is an input of this launch, so any copy is current)
                if (lane == 0 && r0 > 0) up_c = __ldg(y1in + off - 1);
                // pull a later column towards L2; one 32-byte sector per prefetch
                if (a.prefetch_dist > 0 && c + a.prefetch_dist <= c_last && ((r0 * (int)sizeof(Real)) & 31) == 0) {
                    const size_t pf = (size_t)(c + a.prefetch_dist) * M + r0;
                    prefetch_l2(xin + pf); prefetch_l2(fin + pf); prefetch_l2(y1in + pf); prefetch_l2(y2in + pf);
                }
            }
            // ---- primal update of column c ---------------------------------------
            Real xb_c[VEC], xn_c[VEC];
            {
                Real up = __shfl_up_sync(0xffffffffu, y1_c[VEC - 1], 1);
                if (lane == 0) up = up_c;  // 0 at the top row (r0 == 0 → up_c stays 0)
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const Real y1up = (v == 0) ? up : y1_c[v - 1];
                    xn_c[v] = primal_update<Real, STRICT>(x_c[v], f_c[v], y1up, y1_c[v], y2_p[v], y2_c[v], sc,
                                                          xb_c[v]);
                }
                if (rows_ok && c < c1) IO::st(xout + (size_t)c * M + r0, xn_c);
            }
            // ---- Δy1 of column c: needs x̄ of the next row ------------------------
            Real d1_c[VEC];
            {
                Real dn = __shfl_down_sync(0xffffffffu, xb_c[0], 1);
                if (multi_warp) {
                    if (lane == 0) s_xb[c & 1][warp] = xb_c[0];
                    __syncthreads();
                    if (lane == 31) dn = s_xb[c & 1][warp + 1];
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const Real nxt = (v == VEC - 1) ? dn : xb_c[v + 1];
                    const bool last_row = (r0 + v == M - 1);
                    d1_c[v] = last_row ? (Real)0 : (STRICT ? StrictOps<Real>::sub(nxt, xb_c[v]) : nxt - xb_c[v]);
                }
            }
            // ---- dual update of column c-1 (Δy2 = x̄(:,c) - x̄(:,c-1)) --------------
            if (c > c0) {
                // (per-pixel projections on purpose: this kernel is bound by HBM, not by the √/÷ chain, and the joint
                // projection of a thread's rows — one convergence region around all of them — kept the compiler from
                // hoisting the next column's loads: 104.9 → 68.8 Gpixel-iter/s, measured)
                Real o1[VEC], o2[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const Real d2 = STRICT ? StrictOps<Real>::sub(xb_c[v], xb_p[v]) : xb_c[v] - xb_p[v];
                    o1[v] = y1_p[v]; o2[v] = y2_p[v];
                    const Real al = MAP ? al_p[v] : alpha_s;
                    if (has_rho) dual_update_rho<Real, STRICT>(o1[v], o2[v], d1_p[v], d2, al, a.rho, sc);
                    else dual_update<Real, STRICT, false>(o1[v], o2[v], d1_p[v], d2, al, a.rho, sc);
                }
                if (rows_ok) {
                    IO::st(y1out + (size_t)(c - 1) * M + r0, o1);
                    IO::st(y2out + (size_t)(c - 1) * M + r0, o2);
                }
            }
            // ---- carry column c -------------------------------------------------
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                xb_p[v] = xb_c[v]; d1_p[v] = d1_c[v]; y1_p[v] = y1_c[v]; y2_p[v] = y2_c[v];
                if (MAP) al_p[v] = al_c[v];
            }
        }

        // last image column: Δy2 = 0 there, nobody looks ahead
        if (c1 == N) {
            Real o1[VEC], o2[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                o1[v] = y1_p[v]; o2[v] = y2_p[v];
                const Real al = MAP ? al_p[v] : alpha_s;
                if (has_rho) dual_update_rho<Real, STRICT>(o1[v], o2[v], d1_p[v], (Real)0, al, a.rho, sc);
                else dual_update<Real, STRICT, false>(o1[v], o2[v], d1_p[v], (Real)0, al, a.rho, sc);
            }
            if (rows_ok) {
                IO::st(y1out + (size_t)(N - 1) * M + r0, o1);
                IO::st(y2out + (size_t)(N - 1) * M + r0, o2);
            }
        }
        if (multi_warp) __syncthreads();  // s_xb parity slots are reused by the next segment
    }
}

}  // namespace bpltv
