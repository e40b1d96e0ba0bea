// nd_tv.cuh — the adjoint systems of the TV learning function in the form nd_solver.cuh factorises (fp64).
//
// gradient / gradient_reg of /root/reference/src/TVLearningFunctionVec.jl (:98-135, :137-161 scalar λ;
// :219-254, :192-215 patch λ) solve one sparse system per image.  Two formulations, both SPD with the
// unknowns on the pixel grid and couplings across one pixel (W = 1):
//
//  MULT  (gradient, the non-regularised branch).  The reference's block system [I −Gᵀ; Act·G + Inact·α(Den−prodKuKu)G,
//        Inact + eps·Act] is, after eliminating p, (diag(E) + B Bᵀ) ζ = B r in MULTIPLIER space: one unknown per sloped
//        pixel (direction t ⟂ ∇u, compliance E = |∇u|/α), two per flat pixel (E = eps()); p = r − Bᵀζ.  All entries are
//        O(1), where the node-space form carries 1/eps.  (Same formulation as round 1's banded Cholesky, gradient.cuh,
//        and as oracle.gradient_dual.)
//  NODE  (gradient_reg).  (I + α Gᵀ(B−C)G) p = ū − u as the reference writes it (:157); the patch variant
//        `I + α[:] .* Gᵀ(B−C)G` (:212) scales ROW v by α_v, so it is diag(α)·(diag(1/α) + Gᵀ(B−C)G): we factor the
//        symmetric bracket with right-hand side (ū − u)/α.  One unknown per node whatever the image looks like.
//
// Kernels: classify (per-pixel tensors / modes), stencil (the matrix in pixel-stencil form + right-hand side),
// residual (matrix-free, through the difference stencils — independent of the assembled matrix), finish
// (per-pixel functional, PatchOp-adjoint sums).  Device code only; emulation-compatible (-DBPLTV_EMU).
#pragma once
#include "nd_solver.cuh"

namespace bpltv {

constexpr int NDTV_PLANES = 10;
// MULT planes: 0 ea 1 eb 2 E 3 w1 4 w2 6 r = u − ū 7 p
// NODE planes: 0 t11 1 t12 2 t22 3 w1 4 w2 5 c (diagonal) 6 rhs 7 p 8 work
struct NdTvSlots {
    int n, N;
    double *pix; size_t pix_stride;
    int *off; size_t off_stride;         // MULT: N+1 per slot
    double *vec; size_t vec_stride;      // MULT: 3 vectors of 2N per slot — b, ζ, work
    int *info;                           // the solver's: 4 per slot
};

struct NdTvVariant {
    int patch, lm, ln;
    double alpha_s, gamma, act_tol, eps_act, relres_tol;
};

static __device__ __forceinline__ double ndtv_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the CTA, valid in every thread; red: 33 doubles of shared memory
static __device__ double ndtv_block_sum(double v, double *red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    v = ndtv_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double r = lane < nwarps ? red[lane] : 0.0;
        r = ndtv_warp_sum(r);
        if (lane == 0) red[32] = r;
    }
    __syncthreads();
    return red[32];
}

// ===========================================================================
// MULT
// ===========================================================================
// per-pixel modes + exclusive scan of the mode counts (one CTA per image)
template <typename Real>
__global__ void __launch_bounds__(1024) ndtv_classify_mult_kernel(NdTvSlots ws, NdTvVariant gv, const Real *u_all,
                                                                 const Real *ubar_all, const Real *alpha_map, int img0)
{
    __shared__ int s_warp[33];
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const Real *u = u_all + (size_t)(img0 + slot) * N;
    const Real *ub = ubar_all + (size_t)(img0 + slot) * N;
    double *pix = ws.pix + ws.pix_stride * slot;
    double *ea = pix, *eb = pix + N, *E = pix + 2 * (size_t)N, *w1 = pix + 3 * (size_t)N, *w2 = pix + 4 * (size_t)N,
           *rc = pix + 6 * (size_t)N;
    int *off = ws.off + ws.off_stride * slot;
    if (threadIdx.x < 4) ws.info[4 * slot + threadIdx.x] = 0;
    // tiles of blockDim.x consecutive pixels (coalesced), a CTA-wide exclusive scan of the mode counts per tile and a
    // running total carried from tile to tile (threads beyond the image take part in the scan with a count of 0)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int base = 0;
    for (int q0 = 0; q0 < N; q0 += blockDim.x) {
        const int q = q0 + (int)threadIdx.x;
        int cnt = 0;
        if (q < N) {
            const int i = q % n, j = q / n;
            const double uq = (double)u[q];
            const double g1 = (i + 1 < n) ? (double)u[q + 1] - uq : 0.0;
            const double g2 = (j + 1 < n) ? (double)u[q + n] - uq : 0.0;
            const double nrm = sqrt(g1 * g1 + g2 * g2);
            const double a = gv.patch ? (double)alpha_map[q] : gv.alpha_s;
            const bool iso = nrm < gv.act_tol;                      // "active" (:109, :231)
            ea[q] = iso ? 1.0 : -g2 / nrm;
            eb[q] = iso ? 0.0 : g1 / nrm;
            E[q] = iso ? gv.eps_act : nrm / a;
            w1[q] = iso ? 0.0 : g1 / nrm;
            w2[q] = iso ? 0.0 : g2 / nrm;
            rc[q] = uq - (double)ub[q];                             // u − ū (:130, :247)
            cnt = iso ? 2 : 1;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = lane < nw ? s_warp[lane] : 0, iv = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, iv, o);
                if (lane >= o) iv += t;
            }
            s_warp[lane] = iv - v;
            if (lane == 31) s_warp[32] = iv;
        }
        __syncthreads();
        if (q < N) off[q] = base + s_warp[warp] + incl - cnt;
        base += s_warp[32];
        __syncthreads();                                            // s_warp is rewritten by the next tile
    }
    if (threadIdx.x == 0) off[N] = base;
}

// node coefficients of a mode of pixel (i,j) with direction (e1,e2): β0 at q, β1 at q+1, β2 at q+n
static __device__ __forceinline__ void ndtv_beta(int i, int j, int n, double e1, double e2, double b[3])
{
    b[1] = (i + 1 < n) ? e1 : 0.0;
    b[2] = (j + 1 < n) ? e2 : 0.0;
    b[0] = -(b[1] + b[2]);
}
static __device__ __forceinline__ void ndtv_mode(const double *ea, const double *eb, int q, int m, bool iso, double &e1, double &e2)
{
    if (iso) { e1 = m == 0 ? 1.0 : 0.0; e2 = m == 0 ? 0.0 : 1.0; }
    else { e1 = ea[q]; e2 = eb[q]; }
}

// diag(E) + B Bᵀ in pixel-stencil form (MB = 2, NH = 5) and b = B r.  grid (slots, chunks)
__global__ void __launch_bounds__(256) ndtv_stencil_mult_kernel(NdTvSlots ws, double *ast_all, size_t ast_stride)
{
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const double *pix = ws.pix + ws.pix_stride * slot;
    const double *ea = pix, *eb = pix + N, *E = pix + 2 * (size_t)N, *rc = pix + 6 * (size_t)N;
    const int *off = ws.off + ws.off_stride * slot;
    double *ast = ast_all + ast_stride * slot;
    double *bvec = ws.vec + ws.vec_stride * slot;
    // forward offsets h = 1..4: (1,0) (−1,1) (0,1) (1,1); shared node = p's slot sp[h], q's slot sq[h] (−1: none)
    const int sp[5] = {0, 1, 2, 2, -1}, sq[5] = {0, 0, 1, 0, -1};
    const int di[5] = {0, 1, -1, 0, 1}, dj[5] = {0, 0, 1, 1, 1};
    for (int p = blockIdx.y * blockDim.x + threadIdx.x; p < N; p += gridDim.y * blockDim.x) {
        const int i = p % n, j = p / n;
        const int a0 = off[p], mp = off[p + 1] - a0;
        double bp[2][3];
        for (int m = 0; m < 2; ++m) {
            double e1 = 0.0, e2 = 0.0;
            if (m < mp) ndtv_mode(ea, eb, p, m, mp == 2, e1, e2);
            ndtv_beta(i, j, n, e1, e2, bp[m]);
        }
        double *blk = ast + (size_t)p * 5 * 4;
        const double r0 = rc[p], r1 = (i + 1 < n) ? rc[p + 1] : 0.0, r2 = (j + 1 < n) ? rc[p + n] : 0.0;
        for (int be = 0; be < 2; ++be)
            for (int al = 0; al < 2; ++al) {
                double v = 0.0;
                if (be < mp && al < mp) {
                    v = bp[be][0] * bp[al][0] + bp[be][1] * bp[al][1] + bp[be][2] * bp[al][2];
                    if (be == al) v += E[p];
                }
                blk[be * 2 + al] = v;
            }
        for (int m = 0; m < mp; ++m) bvec[a0 + m] = bp[m][0] * r0 + bp[m][1] * r1 + bp[m][2] * r2;
        for (int h = 1; h < 5; ++h) {
            double *bh = blk + 4 * h;
            const int ii = i + di[h], jj = j + dj[h];
            const bool in = ii >= 0 && ii < n && jj >= 0 && jj < n && sp[h] >= 0;
            int mq = 0;
            double bq[2][3] = {{0, 0, 0}, {0, 0, 0}};
            if (in) {
                const int q = ii + n * jj;
                mq = off[q + 1] - off[q];
                for (int m = 0; m < mq; ++m) {
                    double e1, e2;
                    ndtv_mode(ea, eb, q, m, mq == 2, e1, e2);
                    ndtv_beta(ii, jj, n, e1, e2, bq[m]);
                }
            }
            for (int be = 0; be < 2; ++be)
                for (int al = 0; al < 2; ++al)
                    bh[be * 2 + al] = (in && be < mq && al < mp) ? bq[be][sq[h]] * bp[al][sp[h]] : 0.0;
        }
    }
}

// p = r − Bᵀζ on the nodes of one image (whole CTA)
static __device__ void ndtv_dual_primal(const NdTvSlots &ws, int slot, const double *zeta, double *p)
{
    const int n = ws.n, N = ws.N;
    const double *pix = ws.pix + ws.pix_stride * slot;
    const double *ea = pix, *eb = pix + N, *rc = pix + 6 * (size_t)N;
    const int *off = ws.off + ws.off_stride * slot;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const int i = k % n, j = k / n;
        double s = 0.0;
        for (int w = 0; w < 3; ++w) {       // pixel k (node k is its β0), k−1 (β1), k−n (β2)
            if ((w == 1 && i == 0) || (w == 2 && j == 0)) continue;
            const int q = w == 0 ? k : (w == 1 ? k - 1 : k - n);
            const int qi = w == 1 ? i - 1 : i, qj = w == 2 ? j - 1 : j;
            const int o0 = off[q], nm = off[q + 1] - o0;
            for (int m = 0; m < nm; ++m) {
                double e1, e2, b[3];
                ndtv_mode(ea, eb, q, m, nm == 2, e1, e2);
                ndtv_beta(qi, qj, n, e1, e2, b);
                s += b[w] * zeta[o0 + m];
            }
        }
        p[k] = rc[k] - s;
    }
}

// p = r − Bᵀζ; res = B p − E ζ (through the stencils) → work; relres = ‖res‖/‖b‖.  One CTA per image.
__global__ void __launch_bounds__(1024) ndtv_residual_mult_kernel(NdTvSlots ws, double *relres_img, int img0)
{
    __shared__ double red[33];
    const int slot = blockIdx.x, tid = threadIdx.x;
    const int n = ws.n, N = ws.N;
    double *pix = ws.pix + ws.pix_stride * slot;
    const double *ea = pix, *eb = pix + N, *E = pix + 2 * (size_t)N;
    double *p = pix + 7 * (size_t)N;
    const int *off = ws.off + ws.off_stride * slot;
    double *vec = ws.vec + ws.vec_stride * slot;
    const double *bvec = vec, *zeta = vec + 2 * (size_t)N;
    double *work = vec + 4 * (size_t)N;
    const int Nd = off[N];
    double bn2 = 0.0;
    for (int a = tid; a < Nd; a += blockDim.x) bn2 = fma(bvec[a], bvec[a], bn2);
    bn2 = ndtv_block_sum(bn2, red);
    ndtv_dual_primal(ws, slot, zeta, p);
    __syncthreads();
    double rn2 = 0.0;
    for (int q = tid; q < N; q += blockDim.x) {
        const int i = q % n, j = q / n;
        const double pq = p[q];
        const double d1 = (i + 1 < n) ? p[q + 1] - pq : 0.0;
        const double d2 = (j + 1 < n) ? p[q + n] - pq : 0.0;
        const int o0 = off[q], nm = off[q + 1] - o0;
        for (int m = 0; m < nm; ++m) {
            double e1, e2;
            ndtv_mode(ea, eb, q, m, nm == 2, e1, e2);
            const double r = e1 * d1 + e2 * d2 - E[q] * zeta[o0 + m];
            work[o0 + m] = r;
            rn2 = fma(r, r, rn2);
        }
    }
    rn2 = ndtv_block_sum(rn2, red);
    if (tid == 0) relres_img[img0 + slot] = bn2 > 0.0 ? sqrt(rn2 / bn2) : 0.0;
}

// y += x over the first cnt[slot] entries (MULT: cnt = off[N]); grid (slots, chunks)
__global__ void __launch_bounds__(256) ndtv_axpy_mult_kernel(NdTvSlots ws, int dst, int src)
{
    const int slot = blockIdx.x;
    const int Nd = ws.off[ws.off_stride * slot + ws.N];
    double *vec = ws.vec + ws.vec_stride * slot;
    double *y = vec + (size_t)dst * 2 * ws.N;
    const double *x = vec + (size_t)src * 2 * ws.N;
    for (int a = blockIdx.y * blockDim.x + threadIdx.x; a < Nd; a += gridDim.y * blockDim.x) y[a] += x[a];
}
__global__ void __launch_bounds__(256) ndtv_copy_mult_kernel(NdTvSlots ws, int dst, int src)
{
    const int slot = blockIdx.x;
    const int Nd = ws.off[ws.off_stride * slot + ws.N];
    double *vec = ws.vec + ws.vec_stride * slot;
    double *y = vec + (size_t)dst * 2 * ws.N;
    const double *x = vec + (size_t)src * 2 * ws.N;
    for (int a = blockIdx.y * blockDim.x + threadIdx.x; a < Nd; a += gridDim.y * blockDim.x) y[a] = x[a];
}

// functional −Σ⟨(Gp)_q, w_q⟩ per pixel (:132, :249) and its patch sums.  grid (slots, patch groups).
// A vanished pivot or a residual above the tolerance poisons the result with NaN (→ BPLTV_ERR_NUMERIC).
__global__ void __launch_bounds__(512) ndtv_finish_mult_kernel(NdTvSlots ws, NdTvVariant gv, const double *relres_img,
                                                               double *out_img, int img0)
{
    __shared__ double red[33];
    const int slot = blockIdx.x, tid = threadIdx.x;
    const int n = ws.n, N = ws.N;
    double *pix = ws.pix + ws.pix_stride * slot;
    const double *w1 = pix + 3 * (size_t)N, *w2 = pix + 4 * (size_t)N, *p = pix + 7 * (size_t)N;
    const int ng = gv.lm * gv.ln;
    const bool poisoned = ws.info[4 * slot + 1] != 0 || !(relres_img[img0 + slot] <= gv.relres_tol);
    if (ng == 1) {
        double acc = 0.0;
        for (int q = tid; q < N; q += blockDim.x) {
            const int i = q % n, j = q / n;
            const double pq = p[q];
            const double d1 = (i + 1 < n) ? p[q + 1] - pq : 0.0;
            const double d2 = (j + 1 < n) ? p[q + n] - pq : 0.0;
            acc -= d1 * w1[q] + d2 * w2[q];
        }
        acc = ndtv_block_sum(acc, red);
        if (tid == 0 && blockIdx.y == 0) out_img[(size_t)(img0 + slot)] = poisoned ? nan("") : acc;
        return;
    }
    // every patch group recomputes the (cheap) per-pixel functional into its own pass over the pixels it needs
    for (int g = blockIdx.y; g < ng; g += gridDim.y) {
        const int pi = g % gv.lm, pj = g / gv.lm;
        const int i0 = (int)(((long long)pi * n + gv.lm - 1) / gv.lm), i1 = (int)(((long long)(pi + 1) * n + gv.lm - 1) / gv.lm);
        const int j0 = (int)(((long long)pj * n + gv.ln - 1) / gv.ln), j1 = (int)(((long long)(pj + 1) * n + gv.ln - 1) / gv.ln);
        const int h = i1 - i0, w = j1 - j0;
        double acc = 0.0;
        for (int k = tid; k < h * w; k += blockDim.x) {
            const int i = i0 + k % h, j = j0 + k / h, q = i + n * j;
            const double pq = p[q];
            const double d1 = (i + 1 < n) ? p[q + 1] - pq : 0.0;
            const double d2 = (j + 1 < n) ? p[q + n] - pq : 0.0;
            const double v = -(d1 * w1[q] + d2 * w2[q]);
            acc += v;
        }
        acc = ndtv_block_sum(acc, red);
        if (tid == 0) out_img[(size_t)(img0 + slot) * ng + g] = poisoned ? nan("") : acc;
    }
}

// ===========================================================================
// NODE
// ===========================================================================
// per pixel: T = (B − C)·[α] and the functional weights w (:144-156, :199-211); per node: diagonal c and right-hand side.
// grid (slots, chunks)
template <typename Real>
__global__ void __launch_bounds__(256) ndtv_classify_node_kernel(NdTvSlots ws, NdTvVariant gv, const Real *u_all,
                                                                 const Real *ubar_all, const Real *alpha_map, int img0)
{
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const Real *u = u_all + (size_t)(img0 + slot) * N;
    const Real *ub = ubar_all + (size_t)(img0 + slot) * N;
    double *pix = ws.pix + ws.pix_stride * slot;
    if (blockIdx.y == 0 && threadIdx.x < 4) ws.info[4 * slot + threadIdx.x] = 0;
    for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < N; q += gridDim.y * blockDim.x) {
        const int i = q % n, j = q / n;
        const double uq = (double)u[q];
        const double g1 = (i + 1 < n) ? (double)u[q + 1] - uq : 0.0;
        const double g2 = (j + 1 < n) ? (double)u[q + n] - uq : 0.0;
        const double nrm = sqrt(g1 * g1 + g2 * g2);
        const bool act = fmax(0.0, nrm - 1.0 / gv.gamma) != 0.0;        // :146-147, :201-202
        const double sc = gv.patch ? 1.0 : gv.alpha_s;                  // scalar: α folded into T
        double t11, t12, t22, v1, v2;
        if (act) {      // −C = Den − prodesc(Gu/den³, Gu) (:150-155); w = Den·Gu
            const double id = 1.0 / nrm, d3 = nrm * nrm * nrm;
            const double a1 = g1 / d3, a2 = g2 / d3;
            t11 = sc * (id - a1 * g1); t12 = sc * -(a1 * g2); t22 = sc * (id - a2 * g2);
            v1 = id * g1; v2 = id * g2;
        } else {        // B = γ·Inact; w = γ·Gu
            t11 = sc * gv.gamma; t12 = 0.0; t22 = sc * gv.gamma;
            v1 = gv.gamma * g1; v2 = gv.gamma * g2;
        }
        pix[q] = t11; pix[(size_t)N + q] = t12; pix[2 * (size_t)N + q] = t22;
        pix[3 * (size_t)N + q] = v1; pix[4 * (size_t)N + q] = v2;
        const double r = (double)ub[q] - uq;                            // ū − u (:157, :212)
        if (gv.patch) {
            const double a = (double)alpha_map[q];
            pix[5 * (size_t)N + q] = 1.0 / a;
            pix[6 * (size_t)N + q] = r / a;
        } else {
            pix[5 * (size_t)N + q] = 1.0;
            pix[6 * (size_t)N + q] = r;
        }
    }
}

// coefficient 2-vector of node slot s (0: q, 1: q+1, 2: q+n) in (Gp) of pixel (i,j)
static __device__ __forceinline__ void ndtv_gcoef(int i, int j, int n, int s, double &c1, double &c2)
{
    const double h1 = (i + 1 < n) ? 1.0 : 0.0, h2 = (j + 1 < n) ? 1.0 : 0.0;
    if (s == 0) { c1 = -h1; c2 = -h2; }
    else if (s == 1) { c1 = h1; c2 = 0.0; }
    else { c1 = 0.0; c2 = h2; }
}

// diag(c) + GᵀTG in pixel-stencil form (MB = 1, NH = 5): ast[v·5 + h] = A[v + d_h, v]; p ← rhs (the solve is in place).
__global__ void __launch_bounds__(256) ndtv_stencil_node_kernel(NdTvSlots ws, double *ast_all, size_t ast_stride)
{
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    double *pix = ws.pix + ws.pix_stride * slot;
    const double *t11 = pix, *t12 = pix + N, *t22 = pix + 2 * (size_t)N, *cd = pix + 5 * (size_t)N;
    double *ast = ast_all + ast_stride * slot;
    for (int v = blockIdx.y * blockDim.x + threadIdx.x; v < N; v += gridDim.y * blockDim.x) {
        const int i = v % n, j = v / n;
        double acc[5] = {cd[v], 0.0, 0.0, 0.0, 0.0};
        // pixels containing node v: q = v (slot 0), v−1 (slot 1), v−n (slot 2); forward partners of v inside each
        for (int w = 0; w < 3; ++w) {
            if ((w == 1 && i == 0) || (w == 2 && j == 0)) continue;
            const int q = w == 0 ? v : (w == 1 ? v - 1 : v - n);
            const int qi = w == 1 ? i - 1 : i, qj = w == 2 ? j - 1 : j;
            double c1, c2;
            ndtv_gcoef(qi, qj, n, w, c1, c2);
            const double s1 = t11[q] * c1 + t12[q] * c2, s2 = t12[q] * c1 + t22[q] * c2;      // T·g_v
            for (int s = 0; s < 3; ++s) {
                double e1, e2;
                ndtv_gcoef(qi, qj, n, s, e1, e2);
                const double val = e1 * s1 + e2 * s2;
                // node of slot s relative to v: slot0 = q, slot1 = q+(1,0), slot2 = q+(0,1); v = q + slot-w offset
                const int di = (s == 1) - (w == 1), dj = (s == 2) - (w == 2);
                int h = -1;
                if (di == 0 && dj == 0) h = 0;
                else if (di == 1 && dj == 0) h = 1;
                else if (di == -1 && dj == 1) h = 2;
                else if (di == 0 && dj == 1) h = 3;
                if (h >= 0) acc[h] += val;
            }
        }
        for (int h = 0; h < 5; ++h) ast[(size_t)v * 5 + h] = acc[h];
        pix[7 * (size_t)N + v] = pix[6 * (size_t)N + v];
    }
}

// (c + GᵀTG) x at node v, matrix-free; mag = (|c| + |G|ᵀ|T||G|)|x| at v, the scale of the rounding errors of the product
static __device__ __forceinline__ double ndtv_node_apply(const double *pix, int n, int N, const double *x, int v, double &mag)
{
    const double *t11 = pix, *t12 = pix + N, *t22 = pix + 2 * (size_t)N, *cd = pix + 5 * (size_t)N;
    const int i = v % n, j = v / n;
    double s = cd[v] * x[v];
    mag = fabs(s);
    for (int w = 0; w < 3; ++w) {
        if ((w == 1 && i == 0) || (w == 2 && j == 0)) continue;
        const int q = w == 0 ? v : (w == 1 ? v - 1 : v - n);
        const int qi = w == 1 ? i - 1 : i, qj = w == 2 ? j - 1 : j;
        const double xq = x[q];
        const bool h1 = qi + 1 < n, h2 = qj + 1 < n;
        const double d1 = h1 ? x[q + 1] - xq : 0.0, d2 = h2 ? x[q + n] - xq : 0.0;
        const double m1 = h1 ? fabs(x[q + 1]) + fabs(xq) : 0.0, m2 = h2 ? fabs(x[q + n]) + fabs(xq) : 0.0;
        const double s1 = t11[q] * d1 + t12[q] * d2, s2 = t12[q] * d1 + t22[q] * d2;          // T (Gx)_q
        const double a1 = fabs(t11[q]) * m1 + fabs(t12[q]) * m2, a2 = fabs(t12[q]) * m1 + fabs(t22[q]) * m2;
        double c1, c2;
        ndtv_gcoef(qi, qj, n, w, c1, c2);
        s += c1 * s1 + c2 * s2;
        mag += fabs(c1) * a1 + fabs(c2) * a2;
    }
    return s;
}

// work = rhs − A p (matrix-free); backward error η = ‖work‖ / (‖rhs‖ + ‖|A||p|‖) → relres_img.  One CTA per image.
// (‖work‖/‖rhs‖ alone cannot fall below eps·‖A‖‖p‖/‖rhs‖ ≈ 1e-9 here: the entries of A reach αγ.)
__global__ void __launch_bounds__(1024) ndtv_residual_node_kernel(NdTvSlots ws, double *relres_img, int img0)
{
    __shared__ double red[33];
    const int slot = blockIdx.x, tid = threadIdx.x;
    const int n = ws.n, N = ws.N;
    double *pix = ws.pix + ws.pix_stride * slot;
    const double *rhs = pix + 6 * (size_t)N, *p = pix + 7 * (size_t)N;
    double *work = pix + 8 * (size_t)N;
    double bn2 = 0.0, rn2 = 0.0, mn2 = 0.0;
    for (int v = tid; v < N; v += blockDim.x) {
        const double b = rhs[v];
        double mag;
        const double r = b - ndtv_node_apply(pix, n, N, p, v, mag);
        work[v] = r;
        bn2 = fma(b, b, bn2);
        rn2 = fma(r, r, rn2);
        mn2 = fma(mag, mag, mn2);
    }
    bn2 = ndtv_block_sum(bn2, red);
    rn2 = ndtv_block_sum(rn2, red);
    mn2 = ndtv_block_sum(mn2, red);
    if (tid == 0) { const double den = sqrt(bn2) + sqrt(mn2); relres_img[img0 + slot] = den > 0.0 ? sqrt(rn2) / den : 0.0; }
}

__global__ void __launch_bounds__(256) ndtv_axpy_node_kernel(NdTvSlots ws, int dst, int src)
{
    const int slot = blockIdx.x, N = ws.N;
    double *pix = ws.pix + ws.pix_stride * slot;
    double *y = pix + (size_t)dst * N;
    const double *x = pix + (size_t)src * N;
    for (int a = blockIdx.y * blockDim.x + threadIdx.x; a < N; a += gridDim.y * blockDim.x) y[a] += x[a];
}

// scalar: Σ_q ⟨(Gp)_q, w_q⟩ (:159); patch: p_v (Gᵀw)_v per node (:213), then the PatchOp-adjoint sums.
__global__ void __launch_bounds__(512) ndtv_finish_node_kernel(NdTvSlots ws, NdTvVariant gv, const double *relres_img,
                                                               double *out_img, int img0)
{
    __shared__ double red[33];
    const int slot = blockIdx.x, tid = threadIdx.x;
    const int n = ws.n, N = ws.N;
    double *pix = ws.pix + ws.pix_stride * slot;
    const double *w1 = pix + 3 * (size_t)N, *w2 = pix + 4 * (size_t)N, *p = pix + 7 * (size_t)N;
    const int ng = gv.lm * gv.ln;
    const bool poisoned = ws.info[4 * slot + 1] != 0 || !(relres_img[img0 + slot] <= gv.relres_tol);
    if (!gv.patch) {
        double acc = 0.0;
        for (int q = tid; q < N; q += blockDim.x) {
            const int i = q % n, j = q / n;
            const double pq = p[q];
            const double d1 = (i + 1 < n) ? p[q + 1] - pq : 0.0;
            const double d2 = (j + 1 < n) ? p[q + n] - pq : 0.0;
            acc += d1 * w1[q] + d2 * w2[q];
        }
        acc = ndtv_block_sum(acc, red);
        if (tid == 0 && blockIdx.y == 0) out_img[(size_t)(img0 + slot)] = poisoned ? nan("") : acc;
        return;
    }
    for (int g = blockIdx.y; g < ng; g += gridDim.y) {
        const int pi = g % gv.lm, pj = g / gv.lm;
        const int i0 = (int)(((long long)pi * n + gv.lm - 1) / gv.lm), i1 = (int)(((long long)(pi + 1) * n + gv.lm - 1) / gv.lm);
        const int j0 = (int)(((long long)pj * n + gv.ln - 1) / gv.ln), j1 = (int)(((long long)(pj + 1) * n + gv.ln - 1) / gv.ln);
        const int h = i1 - i0, w = j1 - j0;
        double acc = 0.0;
        for (int k = tid; k < h * w; k += blockDim.x) {
            const int i = i0 + k % h, j = j0 + k / h, q = i + n * j;
            double s = 0.0;      // (Gᵀw)_q
            if (i > 0) s += w1[q - 1];
            if (i + 1 < n) s -= w1[q];
            if (j > 0) s += w2[q - n];
            if (j + 1 < n) s -= w2[q];
            acc += p[q] * s;
        }
        acc = ndtv_block_sum(acc, red);
        if (tid == 0) out_img[(size_t)(img0 + slot) * ng + g] = poisoned ? nan("") : acc;
    }
}

}  // namespace bpltv
