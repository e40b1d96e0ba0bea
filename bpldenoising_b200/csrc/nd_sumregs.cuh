// nd_sumregs.cuh — the scalar `sumregs_gradient_reg` system (/root/reference/src/SumRegsLearningFunction.jl:112-167) in the
// form nd_solver.cuh factorises (fp64, node space).
//
//     (I + Σ_k α_k G_kᵀ(B_k − C_k)G_k) p = ū − u,      g_k = pᵀ G_kᵀ(Act_k Den_k G_k u + γ Inact_k G_k u)        (:160-166)
//
// with the three difference operators (forward, backward, centred).  For SCALAR α_k the matrix is symmetric positive
// definite, one unknown per node, and couples nodes at most two pixels apart (G_kᵀ T G_k reaches ±2 rows / columns): the
// nested-dissection multifrontal Cholesky of the TV learning function takes it with coupling radius W = 2.  (The PATCH
// variant, :195-262, scales the rows of each term by a different map — not symmetric — and stays on the band LU.)
// The per-(pixel, operator) tensors and weights are lu3_classify's, the matrix-free residual is lu3_apply, the functional
// and the backward-error rule are those of lu3_solve (lu_band.cuh): the band LU remains the second implementation.
//
// Kernels: nd3_stencil (the matrix in pixel-stencil form + the right-hand side), nd3_residual (r − M p through the stencils
// and tensors, independent of the assembled matrix), nd3_axpy, nd3_finish (backward error, functional, sums).
#pragma once
#include "lu_band.cuh"
#include "nd_solver.cuh"

namespace bpltv {

constexpr int ND3_W = 2;
constexpr int ND3_NH = 13;      // nd_nh(2): the node itself + 12 forward offsets

// forward offset index h of nd_symbolic.h (W = 2) for (di, dj), dj > 0 or (dj == 0 and di > 0)
static __host__ __device__ __forceinline__ int nd3_h(int di, int dj) { return dj == 0 ? di : 3 + 5 * (dj - 1) + (di + 2); }

// Row v of I + Σ_k α_k G_kᵀ T_k G_k (13 stencil offsets, the summation order of lu3_assemble), written as the entries
// A[v + d, v] = A[v, v + d] of the forward offsets d; p ← r (the solve runs in place).  The slot's `ast` must be zero.
// grid (slots, chunks)
__global__ void __launch_bounds__(256) nd3_stencil_kernel(LuSlots ws, Lu3Params pr, double *ast_all, size_t ast_stride)
{
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    double *pix = ws.pix + ws.pix_stride * slot;
    double *ast = ast_all + ast_stride * slot;
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
            const double a = pr.alpha[k];
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
        double *row = ast + (size_t)v * ND3_NH;
        row[0] = acc[6];
        row[nd3_h(1, 0)] = acc[7];
        row[nd3_h(2, 0)] = acc[8];
        row[nd3_h(-1, 1)] = acc[9];
        row[nd3_h(0, 1)] = acc[10];
        row[nd3_h(1, 1)] = acc[11];
        row[nd3_h(0, 2)] = acc[12];
        pix[(size_t)LU_PL_P * N + v] = pix[(size_t)LU_PL_R * N + v];
    }
}

// work = r − M p (matrix-free), relres = ‖work‖ / ‖r‖.  One CTA per image.
template <typename Real>
__global__ void __launch_bounds__(512) nd3_residual_kernel(LuSlots ws, Lu3Params pr, double *relres_img, int img0)
{
    __shared__ double red[40];
    const int slot = blockIdx.x, tid = threadIdx.x;
    const int n = ws.n, N = ws.N;
    double *pix = ws.pix + ws.pix_stride * slot;
    const double *r = pix + (size_t)LU_PL_R * N, *p = pix + (size_t)LU_PL_P * N;
    double *work = pix + (size_t)LU_PL_WORK * N;
    double bn2 = 0.0, rn2 = 0.0;
    for (int v = tid; v < N; v += blockDim.x) {
        const double res = r[v] - lu3_apply<Real>(pix, n, N, ws.nops, nullptr, pr, p, v);
        work[v] = res;
        rn2 = fma(res, res, rn2);
        bn2 = fma(r[v], r[v], bn2);
    }
    rn2 = lu_block_sum(rn2, red);
    bn2 = lu_block_sum(bn2, red);
    if (tid == 0) relres_img[img0 + slot] = bn2 > 0.0 ? sqrt(rn2 / bn2) : 0.0;
}

// p += work.  grid (slots, chunks)
__global__ void __launch_bounds__(256) nd3_axpy_kernel(LuSlots ws)
{
    const int slot = blockIdx.x, N = ws.N;
    double *pix = ws.pix + ws.pix_stride * slot;
    double *p = pix + (size_t)LU_PL_P * N;
    const double *work = pix + (size_t)LU_PL_WORK * N;
    for (int v = blockIdx.y * blockDim.x + threadIdx.x; v < N; v += gridDim.y * blockDim.x) p[v] += work[v];
}

// Normwise backward error ‖r − Mp‖ / (‖M‖‖p‖ + ‖r‖) with the bound ‖M‖ ≤ 1 + 18·γ·max α (as lu3_solve) → relres_img; a
// factorisation that broke down or an error above `tol` poisons the result with NaN (→ BPLTV_ERR_NUMERIC).
// g_k = Σ_v p_v (G_kᵀ w_k)_v per operator (:166), summed per patch of the lm×ln grid (1×1 for the scalar parameter).
// `work` must hold the residual of the final p.  One CTA per image.
__global__ void __launch_bounds__(512) nd3_finish_kernel(LuSlots ws, Lu3Params pr, const int *nd_info, double tol, double *out_img,
                                                         double *relres_img, int img0)
{
    __shared__ double red[40];
    const int slot = blockIdx.x, tid = threadIdx.x;
    const int n = ws.n, N = ws.N;
    double *pix = ws.pix + ws.pix_stride * slot;
    const double *r = pix + (size_t)LU_PL_R * N, *p = pix + (size_t)LU_PL_P * N, *work = pix + (size_t)LU_PL_WORK * N;
    double *fk = pix + (size_t)LU_PL_F * N;
    double pn2 = 0.0, bn2 = 0.0, rn2 = 0.0;
    for (int v = tid; v < N; v += blockDim.x) {
        pn2 = fma(p[v], p[v], pn2);
        bn2 = fma(r[v], r[v], bn2);
        rn2 = fma(work[v], work[v], rn2);
    }
    pn2 = lu_block_sum(pn2, red);
    bn2 = lu_block_sum(bn2, red);
    rn2 = lu_block_sum(rn2, red);
    double amax = 0.0;
    for (int k = 0; k < ws.nops; ++k) amax = fmax(amax, pr.alpha[k]);
    const double berr = sqrt(rn2) / ((1.0 + 18.0 * pr.gamma * amax) * sqrt(pn2) + sqrt(bn2) + 1e-300);
    const bool failed = nd_info[4 * slot + 1] != 0 || !(berr <= tol);
    const int ng = pr.lm * pr.ln;
    for (int k = 0; k < ws.nops; ++k) {
        const double *w1 = pix + (size_t)(6 * k + 4) * N, *w2 = pix + (size_t)(6 * k + 5) * N;
        for (int v = tid; v < N; v += blockDim.x) {
            double s = 0.0;
            visit_node(v % n, v / n, n, [&](int q, int kk, double c1, double c2) {
                if (kk == k) s += c1 * w1[q] + c2 * w2[q];
            });
            fk[v] = p[v] * s;
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
    if (tid == 0) relres_img[img0 + slot] = berr;
}

}  // namespace bpltv
