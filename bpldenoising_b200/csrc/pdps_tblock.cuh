// pdps_tblock.cuh — kernel C: temporally blocked HBM-streaming PDPS ("pipelined march").
//
// One launch = T consecutive primal-dual iterations over the whole M×N×O stack for the
// HBM traffic of ONE: read {x, y1, y2, f}, write {x, y1, y2} = 7 words per pixel per
// launch (8 with a λ-map), i.e. 7/T words per pixel-iteration, plus a halo of T-1 columns
// per column range.
//
// Same decomposition as kernel A (pdps_march.cuh): the N·O image columns are cut into one
// contiguous range per CTA, a CTA spans the whole column height and marches along j.  The
// T iterations are software-pipelined along the march, entirely in registers: at march
// step c
//     stage 0  forms x¹, x̄¹ of column c       and finishes y¹ of column c-1,
//     stage 1  forms x², x̄² of column c-1     and finishes y² of column c-2,   …
//     stage s  forms x^{s+1} of column c-s    and finishes y^{s+1} of column c-s-1,
// each stage consuming the duals the previous stage finished in the same step, the x the
// previous stage formed one step earlier and its own carried column (x̄, Δy1, old duals).
// Only stage 0 reads state from memory and only stage T-1 writes it.  Row neighbours come
// from warp shuffles; the first/last row of each warp crosses through one shared-memory
// slot per warp, stage and direction.
//
// Range ends: a range [c0,c1) of an image starts marching at c0-(T-1) and loads up to
// c1+T-1 (clipped to the image), recomputing the T-1 halo columns on each side; stage s is
// exact from column cs+s on (its left neighbour y2^s(cs+s-1) is unknown before that), which
// is exactly what stage T-1 needs at c0.  At the image's right edge each stage "flushes"
// the dual of column N-1 with Δy2 = 0, as kernel A does.
//
// The arithmetic is common.cuh's primal_update / dual_update in the same order per pixel
// and iteration, so strict mode stays bit-identical to the oracle.
#pragma once
#include "common.cuh"

namespace bpltv {

template <typename Real, int T>
struct TBlockArgs {
    const Real *x_in, *y1_in, *y2_in, *f;
    Real *x_out, *y1_out, *y2_out;
    const Real *alpha_map;            // M×N (shared by all images) or nullptr
    StepConsts<Real> sc[T];           // constants of the T iterations of this launch
    int M, N, O;
    long long total_cols;             // N·O
    Real alpha_s;
};

template <typename Real, int VEC, int T, bool MAP, bool STRICT, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) pdps_tblock_kernel(const TBlockArgs<Real, T> a)
{
    typedef VecIO<Real, VEC> IO;
    // slot [warp] of s_dn: x̄ of the warp's first row (read by the warp above it);
    // slot [warp+1] of s_up: the finished y1 of the warp's last row (read by the warp below)
    __shared__ Real s_dn[T][2][33];
    __shared__ Real s_up[T][2][34];

    const int M = a.M, N = a.N;
    const int r0 = threadIdx.x * VEC;
    const bool rows_ok = r0 < M;          // M % VEC == 0 → all VEC rows valid together
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool multi_warp = blockDim.x > 32;

    const long long per = (a.total_cols + gridDim.x - 1) / gridDim.x;
    long long g = (long long)blockIdx.x * per;
    const long long g_end = min(a.total_cols, g + per);

    while (g < g_end) {
        const int o = (int)(g / N);
        const int c0 = (int)(g - (long long)o * N);
        const int c1 = (int)min((long long)N, (long long)c0 + (g_end - g));  // exclusive
        g += c1 - c0;
        const size_t img = (size_t)o * M * N;
        const Real *xin = a.x_in + img, *y1in = a.y1_in + img, *y2in = a.y2_in + img, *fin = a.f + img;
        Real *xout = a.x_out + img, *y1out = a.y1_out + img, *y2out = a.y2_out + img;

        const int cs = max(0, c0 - (T - 1));                  // first marched column
        const int c_end = (c1 == N) ? N + T - 1 : c1 + T - 1;  // last march step

        // per-stage carried column: x̄, Δy1, the stage's input duals; x input of the next step;
        // delay line of f
        Real xb_p[T][VEC], d1_p[T][VEC], y1_p[T][VEC], y2_p[T][VEC], xh[T][VEC], fd[T][VEC];
#pragma unroll
        for (int s = 0; s < T; ++s)
#pragma unroll
            for (int v = 0; v < VEC; ++v) xb_p[s][v] = d1_p[s][v] = y1_p[s][v] = y2_p[s][v] = xh[s][v] = fd[s][v] = 0;
        if (rows_ok && cs > 0) IO::ld(y2in + (size_t)(cs - 1) * M + r0, y2_p[0]);

        for (int c = cs; c <= c_end; ++c) {
            const int par = c & 1;
            Real cy1[VEC], cy2[VEC], cx[VEC];  // hand-over from stage s-1 to stage s within this step
#pragma unroll
            for (int v = 0; v < VEC; ++v) cy1[v] = cy2[v] = cx[v] = 0;

#pragma unroll
            for (int s = 0; s < T; ++s) {
                const int p = c - s;                              // this stage's primal column
                const int pl = min(c1 + (T - 1 - s), N - 1);      // last primal column this stage needs
                const bool do_primal = p >= cs && p <= pl;        // all three are CTA-uniform
                const bool do_dual = p - 1 >= cs && p <= pl;      // finishes column p-1 with x̄(:,p)
                const bool do_flush = p == N && pl == N - 1;      // finishes column N-1 with Δy2 = 0
                const bool last = s == T - 1;
                const StepConsts<Real> &sc = a.sc[s];

                // ---- inputs of the primal update of column p ---------------------------
                Real x_c[VEC], f_c[VEC], y1_c[VEC], y2_c[VEC], up_c = 0;
                if (s == 0) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x_c[v] = f_c[v] = y1_c[v] = y2_c[v] = 0;
                    if (do_primal && rows_ok) {
                        const size_t off = (size_t)p * M + r0;
                        IO::ld(xin + off, x_c);
                        IO::ld(fin + off, f_c);
                        IO::ld(y1in + off, y1_c);
                        IO::ld(y2in + off, y2_c);
                        if (lane == 0 && r0 > 0) up_c = __ldg(y1in + off - 1);
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) fd[0][v] = f_c[v];
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        x_c[v] = xh[s][v]; xh[s][v] = cx[v];     // x^s(:,p), formed one step ago
                        f_c[v] = fd[s][v];
                        y1_c[v] = cy1[v]; y2_c[v] = cy2[v];      // y^s(:,p), finished in this step
                    }
                    if (multi_warp) {
                        __syncthreads();
                        if (lane == 0 && warp > 0) up_c = s_up[s][par][warp];
                    }
                }

                // ---- primal update + over-relaxation of column p -----------------------
                Real xb_c[VEC], xn_c[VEC], d1_c[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) xb_c[v] = xn_c[v] = d1_c[v] = 0;
                if (do_primal) {
                    Real up = __shfl_up_sync(0xffffffffu, y1_c[VEC - 1], 1);
                    if (lane == 0) up = up_c;  // 0 at the top row
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const Real y1up = (v == 0) ? up : y1_c[v - 1];
                        xn_c[v] = primal_update<Real, STRICT>(x_c[v], f_c[v], y1up, y1_c[v], y2_p[s][v], y2_c[v], sc,
                                                              xb_c[v]);
                    }
                    if (last && rows_ok && p >= c0 && p < c1) IO::st(xout + (size_t)p * M + r0, xn_c);
                    if (multi_warp && lane == 0) s_dn[s][par][warp] = xb_c[0];
                }
                if (multi_warp) __syncthreads();
                if (do_primal) {
                    Real dn = __shfl_down_sync(0xffffffffu, xb_c[0], 1);
                    if (multi_warp && lane == 31) dn = s_dn[s][par][warp + 1];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const Real nxt = (v == VEC - 1) ? dn : xb_c[v + 1];
                        const bool last_row = (r0 + v == M - 1);
                        d1_c[v] = last_row ? (Real)0 : (STRICT ? StrictOps<Real>::sub(nxt, xb_c[v]) : nxt - xb_c[v]);
                    }
                }

                // ---- dual update of column p-1 ------------------------------------------
                if (do_dual || do_flush) {
                    Real o1[VEC], o2[VEC], al[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) al[v] = a.alpha_s;
                    if (MAP && rows_ok) IO::ld(a.alpha_map + (size_t)(p - 1) * M + r0, al);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        Real d2 = (Real)0;
                        if (do_dual) d2 = STRICT ? StrictOps<Real>::sub(xb_c[v], xb_p[s][v]) : xb_c[v] - xb_p[s][v];
                        o1[v] = y1_p[s][v]; o2[v] = y2_p[s][v];
                        dual_update<Real, STRICT, false>(o1[v], o2[v], d1_p[s][v], d2, al[v], (Real)0, sc);
                    }
                    if (last) {
                        if (rows_ok && p - 1 >= c0) {
                            IO::st(y1out + (size_t)(p - 1) * M + r0, o1);
                            IO::st(y2out + (size_t)(p - 1) * M + r0, o2);
                        }
                    } else {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) { cy1[v] = o1[v]; cy2[v] = o2[v]; }
                        if (multi_warp && lane == 31) s_up[s + 1][par][warp + 1] = o1[VEC - 1];
                    }
                }

                // ---- carry column p ---------------------------------------------------
                if (do_primal) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        xb_p[s][v] = xb_c[v]; d1_p[s][v] = d1_c[v]; y1_p[s][v] = y1_c[v]; y2_p[s][v] = y2_c[v];
                    }
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) cx[v] = xn_c[v];
            }
            // f delay line: stage s reads f of column c-s
#pragma unroll
            for (int s = T - 1; s >= 1; --s)
#pragma unroll
                for (int v = 0; v < VEC; ++v) fd[s][v] = fd[s - 1][v];
        }
        if (multi_warp) __syncthreads();  // exchange slots are reused by the next segment
    }
}

}  // namespace bpltv
