// pdps_march.cuh — kernel A: HBM-streaming fused PDPS iteration ("column march").
//
// One launch = one primal-dual iteration over the whole M×N×O stack.  The
// forward-difference gradient, the dual ascent, the pointwise projection onto
// the λ-ball, the divergence, the primal prox and the over-relaxation are fused;
// x̄, Δx, Δy never touch memory.  Algorithmic traffic: read {x, y1, y2, f},
// write {x, y1, y2} = 7 words per pixel-iteration (8 with a λ-map).
//
// Work decomposition: a CTA owns `chunk` consecutive columns of one image and
// spans the whole column height (blockDim.x·VEC ≥ M rows, VEC consecutive rows
// per thread → 16-byte coalesced loads along Julia's fastest axis).  It marches
// along j: at column c it loads (x,f,y1,y2)(:,c), forms x_new and x̄ for that
// column, then finishes the dual update of column c-1 (which needed x̄(:,c)).
// Column c-1's old duals and x̄ ride in registers, row neighbours come from warp
// shuffles (one shared-memory slot per warp boundary), so the only redundant
// global reads are y2(:,c0-1) and the look-ahead column c1 at the chunk edges.
// Input and output state are separate buffers (ping-pong), so chunks and images
// are independent within a launch.
#pragma once
#include "common.cuh"

namespace bpltv {

template <typename Real>
struct MarchArgs {
    const Real *x_in, *y1_in, *y2_in, *f;
    Real *x_out, *y1_out, *y2_out;
    const Real *alpha_map;            // M×N (shared by all images) or nullptr
    const StepConsts<Real> *steps;    // device array, one entry per iteration
    int it;                           // iteration index into steps
    int M, N, O;
    int chunk, chunks_per_image;
    Real alpha_s, rho;
};

template <typename Real, int VEC, bool MAP, bool STRICT, int MAXT>
__global__ void __launch_bounds__(MAXT) pdps_march_kernel(const MarchArgs<Real> a)
{
    typedef VecIO<Real, VEC> IO;
    __shared__ Real s_xb[2][33];  // x̄ of the first row of each warp, double-buffered by column parity

    const int unit = blockIdx.x;
    const int o = unit / a.chunks_per_image;
    const int ch = unit - o * a.chunks_per_image;
    const int M = a.M, N = a.N;
    const int c0 = ch * a.chunk;
    const int c1 = min(N, c0 + a.chunk);  // exclusive
    const int r0 = threadIdx.x * VEC;
    const bool rows_ok = r0 < M;          // M % VEC == 0 → all VEC rows valid together
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool multi_warp = blockDim.x > 32;
    const bool has_rho = a.rho != (Real)0;

    const StepConsts<Real> sc = a.steps[a.it];
    const size_t img = (size_t)o * M * N;
    const Real *xin = a.x_in + img, *y1in = a.y1_in + img, *y2in = a.y2_in + img, *fin = a.f + img;
    Real *xout = a.x_out + img, *y1out = a.y1_out + img, *y2out = a.y2_out + img;

    // state of the previous column (c-1): x̄, Δy1 = x̄(i+1)-x̄(i), old duals, λ
    Real xb_p[VEC], d1_p[VEC], y1_p[VEC], y2_p[VEC], al_p[VEC];
    // data of the current column c and the prefetched column c+1
    Real x_c[VEC], f_c[VEC], y1_c[VEC], y2_c[VEC], al_c[VEC], up_c = 0;
    Real x_n[VEC], f_n[VEC], y1_n[VEC], y2_n[VEC], al_n[VEC], up_n = 0;

#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        xb_p[v] = d1_p[v] = y1_p[v] = y2_p[v] = 0; al_p[v] = a.alpha_s;
        x_c[v] = f_c[v] = y1_c[v] = y2_c[v] = 0; al_c[v] = a.alpha_s;
        x_n[v] = f_n[v] = y1_n[v] = y2_n[v] = 0; al_n[v] = a.alpha_s;
    }

    // y2 of column c0-1 (zero left of the image)
    if (rows_ok && c0 > 0) IO::ld(y2in + (size_t)(c0 - 1) * M + r0, y2_p);

    const int c_last = min(c1, N - 1);  // last column whose x̄ we need (look-ahead column if c1 < N)

    auto load_col = [&](int c, Real(&x)[VEC], Real(&f)[VEC], Real(&y1)[VEC], Real(&y2)[VEC],
                        Real(&al)[VEC], Real &up) {
        if (rows_ok) {
            const size_t off = (size_t)c * M + r0;
            IO::ld(xin + off, x);
            IO::ld(fin + off, f);
            IO::ld(y1in + off, y1);
            IO::ld(y2in + off, y2);
            if (MAP) IO::ld(a.alpha_map + off, al);
            // y1 of the row above this warp's first row comes from another warp: read it
            // (it is an input of this launch, so any copy is current).
            if (lane == 0 && r0 > 0) up = __ldg(y1in + off - 1);
        }
    };

    load_col(c0, x_c, f_c, y1_c, y2_c, al_c, up_c);

    for (int c = c0; c <= c_last; ++c) {
        if (c + 1 <= c_last) load_col(c + 1, x_n, f_n, y1_n, y2_n, al_n, up_n);

        // ---- primal update of column c -------------------------------------------
        Real xb_c[VEC], xn_c[VEC];
        {
            Real up = __shfl_up_sync(0xffffffffu, y1_c[VEC - 1], 1);
            if (lane == 0) up = up_c;  // 0 at the top row (r0 == 0 → up_c stays 0)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const Real y1up = (v == 0) ? up : y1_c[v - 1];
                xn_c[v] = primal_update<Real, STRICT>(x_c[v], f_c[v], y1up, y1_c[v], y2_p[v], y2_c[v],
                                                      sc, xb_c[v]);
            }
            if (rows_ok && c < c1) IO::st(xout + (size_t)c * M + r0, xn_c);
        }
        // ---- Δy1 of column c: needs x̄ of the next row ----------------------------
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
                d1_c[v] = last_row ? (Real)0
                                   : (STRICT ? StrictOps<Real>::sub(nxt, xb_c[v]) : nxt - xb_c[v]);
            }
        }
        // ---- dual update of column c-1 (Δy2 = x̄(:,c) - x̄(:,c-1)) ------------------
        if (c > c0) {
            Real o1[VEC], o2[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const Real d2 = STRICT ? StrictOps<Real>::sub(xb_c[v], xb_p[v]) : xb_c[v] - xb_p[v];
                o1[v] = y1_p[v]; o2[v] = y2_p[v];
                const Real al = MAP ? al_p[v] : a.alpha_s;
                if (has_rho) dual_update<Real, STRICT, true>(o1[v], o2[v], d1_p[v], d2, al, a.rho, sc);
                else dual_update<Real, STRICT, false>(o1[v], o2[v], d1_p[v], d2, al, a.rho, sc);
            }
            if (rows_ok) {
                IO::st(y1out + (size_t)(c - 1) * M + r0, o1);
                IO::st(y2out + (size_t)(c - 1) * M + r0, o2);
            }
        }
        // ---- rotate registers ----------------------------------------------------
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            xb_p[v] = xb_c[v]; d1_p[v] = d1_c[v]; y1_p[v] = y1_c[v]; y2_p[v] = y2_c[v]; al_p[v] = al_c[v];
            x_c[v] = x_n[v]; f_c[v] = f_n[v]; y1_c[v] = y1_n[v]; y2_c[v] = y2_n[v]; al_c[v] = al_n[v];
        }
        up_c = up_n;
    }

    // last image column: Δy2 = 0 there, nobody looks ahead
    if (c1 == N) {
        Real o1[VEC], o2[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            o1[v] = y1_p[v]; o2[v] = y2_p[v];
            const Real al = MAP ? al_p[v] : a.alpha_s;
            if (has_rho) dual_update<Real, STRICT, true>(o1[v], o2[v], d1_p[v], (Real)0, al, a.rho, sc);
            else dual_update<Real, STRICT, false>(o1[v], o2[v], d1_p[v], (Real)0, al, a.rho, sc);
        }
        if (rows_ok) {
            IO::st(y1out + (size_t)(N - 1) * M + r0, o1);
            IO::st(y2out + (size_t)(N - 1) * M + r0, o2);
        }
    }
}

}  // namespace bpltv
