// gradient_sumregs.cuh — adjoint solve and λ-gradient of the sum-of-regularisers model (fp64).
//
// Replaces sumregs_gradient / sumregs_gradient_reg of /root/reference/src/SumRegsLearningFunction.jl
// (scalar: :264-327 / :112-167; patch, non-regularised: :330-407), which assemble a sparse 7n²
// saddle-point matrix (or an n² matrix) per image and call a sparse LU.
//
// Same compliance-form reformulation as gradient.cuh, with three operators: after eliminating the
// multipliers every variant is (I + Σ_k G_kᵀ D_k G_k) p = r with per-pixel 2×2 tensors D_kq that are
// s·I on flat pixels or s·t tᵀ, t ⟂ ∇_k u; we factor the SPD system in multiplier space
//       (diag(E) + B Bᵀ) ζ = B r ,   p = r − Bᵀ ζ ,   E = 1/s ,
// one unknown ("mode") per sloped (pixel, operator) and two per flat one, numbered pixel-major /
// operator-minor, which makes the matrix banded with half-bandwidth ≤ the modes of 2n+1 pixels.  The
// band is factorised by gradient.cuh's blocked Cholesky (grad_factor_kernel) and solved by its
// band_solve; this file adds the three-operator classification, the assembly and the functional.
// The assembly accumulates the off-diagonal node contributions with atomicAdd; every such entry has at
// most two contributions (addition of two numbers is commutative), the diagonal and the right-hand side
// are summed in a fixed order, so the result is deterministic like the TV path.
//
// The patch variant of sumregs_gradient_reg (:195-262) is row-scaled by a different λ-map per operator,
// cannot be symmetrised and so has no SPD compliance form: it is solved in node space by the band LU of
// lu_band.cuh (run_gradient3_lu below).
#pragma once
#include "env_switches.h"
#include "gradient.cuh"
#include "sumregs_stencils.cuh"
#include "gradient_lu.cuh"

namespace bpltv {

#ifndef BPLTV_SUMREGS_REG_LU_DEFAULT
#define BPLTV_SUMREGS_REG_LU_DEFAULT true
#endif

struct Grad3Variant {
    int regularised, patch, lm, ln;
    double alpha[3], gamma, act_tol, eps_act;
    int refine;
};

// per-slot layout of GradSlots::pix for this path (N = n² doubles each):
//   [5k+0..5k+4] ea, eb, E, w1, w2 of operator k;  [15] r;  [16] p;  [17..19] per-node functional of operator k
// GradSlots::off has 3N+1 entries, index 3q+k.
static __device__ __forceinline__ const double *g3_plane(const double *pix, int N, int idx) { return pix + (size_t)idx * N; }

// β of mode m of (q,k) for stencil coefficients (c1, c2)
static __device__ __forceinline__ double g3_beta(const double *pix, int N, int q, int k, int m, bool iso, double c1, double c2)
{
    if (iso) return m == 0 ? c1 : c2;
    return g3_plane(pix, N, 5 * k)[q] * c1 + g3_plane(pix, N, 5 * k + 1)[q] * c2;
}

// ---------------------------------------------------------------------------
// K1: classification of every (pixel, operator) + exclusive scan of the mode counts
// ---------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(GRAD_THREADS) grad3_classify_kernel(GradSlots ws, Grad3Variant gv, const Real *u_all,
                                                                      const Real *ubar_all, const Real *alpha_maps,
                                                                      int img0)
{
    __shared__ int s_warp[32];
    __shared__ int s_total;
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const Real *u = u_all + (size_t)(img0 + slot) * N;
    const Real *ub = ubar_all + (size_t)(img0 + slot) * N;
    double *pix = slot_ptr(ws.pix, ws.pix_stride, slot);
    int *off = slot_ptr(ws.off, ws.off_stride, slot);
    int *ext = slot_ptr(ws.ext, ws.ext_stride, slot);
    int *info = ws.info + 4 * slot;

    const int per = (N + blockDim.x - 1) / blockDim.x;
    const int q0 = threadIdx.x * per, q1 = min(N, q0 + per);
    int cnt = 0;
    for (int q = q0; q < q1; ++q) {
        const int i = q % n, j = q / n;
        pix[(size_t)15 * N + q] = gv.regularised ? (double)ub[q] - (double)u[q] : (double)u[q] - (double)ub[q];
        for (int k = 0; k < 3; ++k) {
            double g1, g2;
            op_apply<Real>(k, i, j, n, u, q, g1, g2);
            const double nrm = sqrt(g1 * g1 + g2 * g2);
            const double a = gv.patch ? (double)alpha_maps[(size_t)k * N + q] : gv.alpha[k];
            bool iso;
            double e, v1, v2;
            if (gv.regularised) {   // act = max(0,|G_k u|-1/γ) != 0 (:121-123); flat = !act
                iso = !(fmax(0.0, nrm - 1.0 / gv.gamma) != 0.0);
                if (iso) { v1 = gv.gamma * g1; v2 = gv.gamma * g2; e = 1.0 / (a * gv.gamma); }
                else { v1 = g1 / nrm; v2 = g2 / nrm; e = nrm / a; }
            } else {                // act = |G_k u| < 1e-12 (:273)
                iso = nrm < gv.act_tol;
                if (iso) { v1 = 0.0; v2 = 0.0; e = gv.eps_act; }
                else { v1 = g1 / nrm; v2 = g2 / nrm; e = nrm / a; }
            }
            double *pk = pix + (size_t)(5 * k) * N;
            pk[q] = iso ? 1.0 : -g2 / nrm;
            pk[(size_t)N + q] = iso ? 0.0 : g1 / nrm;
            pk[(size_t)2 * N + q] = e;
            pk[(size_t)3 * N + q] = v1;
            pk[(size_t)4 * N + q] = v2;
            off[3 * q + k] = iso ? 2 : 1;
            cnt += iso ? 2 : 1;
        }
    }
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
        s_warp[lane] = iv - v;
        if (lane == 31) s_total = iv;
    }
    __syncthreads();
    int run = s_warp[warp] + incl - cnt;
    for (int e = 3 * q0; e < 3 * q1; ++e) {
        const int c = off[e];
        off[e] = run;
        run += c;
    }
    if (threadIdx.x == 0) { off[3 * N] = s_total; info[0] = s_total; info[2] = 0; }
    __syncthreads();
    // envelope: the modes of pixel q couple with modes of pixels up to q+2n (two centred stencils
    // sharing the node q+n)
    int bwmax = 0;
    for (int q = threadIdx.x; q < N; q += blockDim.x) {
        const int qq = min(q + 2 * n, N - 1);
        const int last = off[3 * qq + 3] - 1;
        for (int a = off[3 * q]; a < off[3 * q + 3]; ++a) {
            ext[a] = last;
            bwmax = max(bwmax, last - a);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bwmax = max(bwmax, __shfl_xor_sync(0xffffffffu, bwmax, o));
    __syncthreads();
    if (lane == 0) s_warp[warp] = bwmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        int m = 0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) m = max(m, s_warp[w]);
        info[1] = min(ws.LD, m + 1 + GRAD_NB);
    }
}

// ---------------------------------------------------------------------------
// K2: band ← diag(E) + B Bᵀ (lower band), b ← B r, accumulated node by node
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GRAD_THREADS) grad3_assemble_kernel(GradSlots ws)
{
    const int slot = blockIdx.x;
    const int n = ws.n, N = ws.N;
    const double *pix = slot_ptr(ws.pix, ws.pix_stride, slot);
    const int *off = slot_ptr(ws.off, ws.off_stride, slot);
    const int *info = ws.info + 4 * slot;
    const int Nd = info[0], LDa = info[1];
    double *ab = slot_ptr(ws.ab, ws.ab_stride, slot);
    double *bvec = slot_ptr(ws.mode, ws.mode_stride, slot);
    const double *rc = g3_plane(pix, N, 15);

    const size_t total = (size_t)Nd * LDa;
    {
        double2 *ab2 = reinterpret_cast<double2 *>(ab);
        const size_t n2 = total >> 1;
        const double2 z2 = make_double2(0.0, 0.0);
        for (size_t k = threadIdx.x; k < n2; k += blockDim.x) ab2[k] = z2;
        if (threadIdx.x == 0 && (total & 1)) ab[total - 1] = 0.0;
    }
    __syncthreads();
    // diagonal E + Σ β² and right-hand side Σ β r of every mode, over the nodes of its own stencil in a
    // fixed order (these sums have up to four terms; the off-diagonal entries below have at most two —
    // two different stencils share at most two nodes — so their atomic accumulation is order-independent
    // and the whole assembly is deterministic)
    for (int e = threadIdx.x; e < 3 * N; e += blockDim.x) {
        const int q = e / 3, k = e - 3 * q, i = q % n, j = q / n;
        const double E = g3_plane(pix, N, 5 * k + 2)[q];
        const int a0 = off[e], nm = off[e + 1] - a0;
        for (int m = 0; m < nm; ++m) {
            double d = E, bsum = 0.0;
            visit_stencil(k, i, j, n, [&](int node, double c1, double c2) {
                const double b = g3_beta(pix, N, q, k, m, nm == 2, c1, c2);
                d += b * b;
                bsum += b * rc[node];
            });
            ab[(size_t)(a0 + m) * LDa] = d;
            bvec[a0 + m] = bsum;
        }
    }
    __syncthreads();
    // node contributions β_a β_a' (C = I)
    for (int v = threadIdx.x; v < N; v += blockDim.x) {
        const int i = v % n, j = v / n;
        int ma[20];
        double mb[20];
        int cnt = 0;
        visit_node(i, j, n, [&](int q, int k, double c1, double c2) {
            const int a0 = off[3 * q + k], nm = off[3 * q + k + 1] - a0;
            const bool iso = nm == 2;
            for (int m = 0; m < nm; ++m) {
                const double b = g3_beta(pix, N, q, k, m, iso, c1, c2);
                if (b != 0.0) { ma[cnt] = a0 + m; mb[cnt] = b; ++cnt; }
            }
        });
        for (int x = 0; x < cnt; ++x)
            for (int y = x + 1; y < cnt; ++y) {
                const int lo = min(ma[x], ma[y]), hi = max(ma[x], ma[y]);
                atomicAdd(&ab[(size_t)lo * LDa + (hi - lo)], mb[x] * mb[y]);
            }
    }
}

// p = r − Bᵀζ on the nodes of one image (whole CTA)
static __device__ void dual_primal3(const GradSlots &ws, int slot, const double *zeta, double *p)
{
    const int n = ws.n, N = ws.N;
    const double *pix = slot_ptr(ws.pix, ws.pix_stride, slot);
    const int *off = slot_ptr(ws.off, ws.off_stride, slot);
    const double *rc = g3_plane(pix, N, 15);
    for (int v = threadIdx.x; v < N; v += blockDim.x) {
        const int i = v % n, j = v / n;
        double s = 0.0;
        visit_node(i, j, n, [&](int q, int k, double c1, double c2) {
            const int a0 = off[3 * q + k], nm = off[3 * q + k + 1] - a0;
            for (int m = 0; m < nm; ++m) s += g3_beta(pix, N, q, k, m, nm == 2, c1, c2) * zeta[a0 + m];
        });
        p[v] = rc[v] - s;
    }
}

// ---------------------------------------------------------------------------
// K4: ζ = A⁻¹b, iterative refinement with the stencil residual, p, functional per operator,
// patch sums.  out_img: 3·lm·ln doubles per image, layout [operator][patch] like the m×n×3 array.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GRAD_THREADS) grad3_solve_kernel(GradSlots ws, Grad3Variant gv, double *out_img,
                                                                   double *relres_img, int img0, int zs_cap)
{
    extern __shared__ __align__(16) unsigned char dyn[];
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
    double *p = pix + (size_t)16 * N;
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
    auto solve = [&](double *vec) {
        if (in_smem) {
            for (int a = tid; a < Nd; a += blockDim.x) zs[a] = vec[a];
            __syncthreads();
            band_solve<26, true>(ab, sinv, s_nr, Nd, LDa, zs, sh);
            for (int a = tid; a < Nd; a += blockDim.x) vec[a] = zs[a];
            __syncthreads();
        } else {
            band_solve<26, true>(ab, sinv, s_nr, Nd, LDa, vec, sh);
        }
    };
    solve(zeta);
    double relres = 0.0;
    for (int it = 0; it <= gv.refine; ++it) {
        dual_primal3(ws, slot, zeta, p);
        __syncthreads();
        double rn2 = 0.0;
        for (int e = tid; e < 3 * N; e += blockDim.x) {
            const int q = e / 3, k = e - 3 * q, i = q % n, j = q / n;
            double d1, d2;
            op_apply<double>(k, i, j, n, p, q, d1, d2);
            const int a0 = off[e], nm = off[e + 1] - a0;
            const double E = g3_plane(pix, N, 5 * k + 2)[q];
            for (int m = 0; m < nm; ++m) {
                const double r = g3_beta(pix, N, q, k, m, nm == 2, d1, d2) - E * zeta[a0 + m];
                work[a0 + m] = r;
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
    // functional per operator: scalar  ±Σ_q ⟨(G_k p)_q, w_kq⟩ (:166, :326);
    // patch (non-regularised)  −p_ν (G_kᵀ w_k)_ν summed over each patch (:395-405)
    const double sign = gv.regularised ? 1.0 : -1.0;
    const int ng = gv.lm * gv.ln;
    for (int k = 0; k < 3; ++k) {
        const double *w1 = g3_plane(pix, N, 5 * k + 3), *w2 = g3_plane(pix, N, 5 * k + 4);
        double *fk = pix + (size_t)(17 + k) * N;
        for (int v = tid; v < N; v += blockDim.x) {
            const int i = v % n, j = v / n;
            double val;
            if (!gv.patch) {
                double d1, d2;
                op_apply<double>(k, i, j, n, p, v, d1, d2);
                val = sign * (d1 * w1[v] + d2 * w2[v]);
            } else {
                double s = 0.0;
                visit_node(i, j, n, [&](int q, int kk, double c1, double c2) {
                    if (kk == k) s += c1 * w1[q] + c2 * w2[q];
                });
                val = -p[v] * s;
            }
            fk[v] = val;
        }
        __syncthreads();
        for (int g = 0; g < ng; ++g) {
            const int pi = g % gv.lm, pj = g / gv.lm;
            double acc = 0.0;
            for (int v = tid; v < N; v += blockDim.x) {
                const int i = v % n, j = v / n;
                const int qi = (int)(((long long)i * gv.lm) / n), qj = (int)(((long long)j * gv.ln) / n);
                if (qi == pi && qj == pj) acc += fk[v];
            }
            acc = block_sum(acc, sh);
            if (tid == 0) out_img[((size_t)(img0 + slot) * 3 + k) * ng + g] = acc;
            __syncthreads();
        }
    }
    if (tid == 0) relres_img[img0 + slot] = relres;
}

template <typename Real>
struct Grad3Problem {
    const Real *u, *ubar;
    int M, N, O;
    double alpha[3];
    const Real *alpha_maps;   // 3 maps of M·N or nullptr
    int lm, ln;
    bool regularised;
    double gamma, act_tol, eps_act;
};

template <typename Real>
static int run_gradient3(GradWork &w, const Grad3Problem<Real> &gp, int sm_count, size_t smem_optin, cudaStream_t st,
                         double *d_grad_out, long long *launches)
{
    const int n = gp.M;
    const int N = gp.M * gp.N;
    const int ng = gp.lm * gp.ln;
    if (gp.M != gp.N) return grad_fail(w, -1, "square images required");
    if (3 * ng > 1024) return grad_fail(w, -1, "lambda grid larger than 341 entries per operator is not supported");
    auto band_lu = [&]() {
        LuProblem<Real> lp;
        lp.u = gp.u; lp.ubar = gp.ubar; lp.M = gp.M; lp.N = gp.N; lp.O = gp.O;
        for (int k = 0; k < 3; ++k) lp.alpha[k] = gp.alpha[k];
        lp.alpha_maps = gp.alpha_maps; lp.lm = gp.lm; lp.ln = gp.ln; lp.gamma = gp.gamma; lp.nops = 3;
        return run_gradient_lu<Real>(w, lp, sm_count, smem_optin, st, d_grad_out, launches);
    };
    if (gp.regularised && gp.alpha_maps) return band_lu();   // row-scaled, non-symmetric (:246): node-space band LU
    {   // scalar sumregs_gradient_reg (:112-167, γ = 1e3): n² unknowns with half-bandwidth 2n in node space
        // instead of ≤ 6n² modes with half-bandwidth ≤ 6(2n+1) in multiplier space.  BPLTV_SUMREGS_REG_LU=0/1.
        const char *lu_env = bpltv::env_get("BPLTV_SUMREGS_REG_LU");
        const bool lu = lu_env && *lu_env ? atoi(lu_env) != 0 : BPLTV_SUMREGS_REG_LU_DEFAULT;
        if (gp.regularised && lu) {
            const int rc = band_lu();
            if (rc != -1) return rc;      // -1: the LU does not take this shape; the compliance form below may
        }
    }
    GradSlots ws;
    ws.N = N; ws.n = n;
    ws.NdMax = 6 * N;
    ws.LD = std::min(ws.NdMax, 6 * (2 * n + 1)) + 1 + GRAD_NB;
    ws.nblkMax = ws.NdMax / GRAD_NB + 1;
    ws.pix_stride = (size_t)20 * N;
    ws.mode_stride = (size_t)3 * ws.NdMax;
    ws.sinv_stride = (size_t)ws.nblkMax * GRAD_NB * GRAD_NB;
    ws.ab_stride = ((size_t)ws.NdMax * ws.LD + 1) & ~(size_t)1;
    ws.off_stride = (size_t)3 * N + 2;
    ws.ext_stride = (size_t)ws.NdMax;
    size_t smem = (size_t)(4 * GRAD_NB * GRAD_NB + GRAD_NB * ((ws.LD + 7) & ~7)) * sizeof(double);
    const size_t stage_bytes = (size_t)32 * (GRAD_THREADS - 32) * sizeof(double);
    const int use_stage = smem + stage_bytes <= smem_optin ? 1 : 0;
    if (use_stage) smem += stage_bytes;
    smem = std::max<size_t>(smem, (size_t)(4 * GRAD_NB * GRAD_NB + 32 * GRAD_NB * GRAD_NB) * sizeof(double));
    if (smem > smem_optin)
        return grad_fail(w, -1, "image too large for the sum-of-regularisers Cholesky panel in shared memory (n <= 136)");

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
    if (w.cap_O < (size_t)gp.O || w.cap_ng < (size_t)(3 * ng)) {
        if (w.out_img) cudaFree(w.out_img);
        if (w.relres) cudaFree(w.relres);
        if (!w.relres_max) cudaMalloc(&w.relres_max, 8);
        cudaError_t e = cudaMalloc(&w.out_img, (size_t)gp.O * 3 * ng * 8);
        if (e == cudaSuccess) e = cudaMalloc(&w.relres, (size_t)gp.O * 8);
        if (e != cudaSuccess) { cudaGetLastError(); w.cap_O = 0; return grad_fail(w, -6, "gradient output allocation failed"); }
        w.cap_O = gp.O; w.cap_ng = 3 * ng;
    }
    ws.pix = (double *)w.pix; ws.mode = (double *)w.mode; ws.sinv = (double *)w.sinv; ws.ab = (double *)w.ab;
    ws.off = (int *)w.off; ws.ext = (int *)w.ext; ws.info = (int *)w.info;

    Grad3Variant gv;
    gv.regularised = gp.regularised ? 1 : 0;
    gv.patch = gp.alpha_maps != nullptr;
    gv.lm = gp.lm; gv.ln = gp.ln;
    for (int k = 0; k < 3; ++k) gv.alpha[k] = gp.alpha[k];
    gv.gamma = gp.gamma; gv.act_tol = gp.act_tol; gv.eps_act = gp.eps_act;
    gv.refine = 1;
    const double guard = 1e-13;

    const size_t nr_bytes = ((size_t)ws.nblkMax * 2 + 15) & ~(size_t)15;
    if (nr_bytes + 8192 + 4096 > smem_optin) return grad_fail(w, -1, "image too large for the solve kernel's block table");
    const size_t zs_only = std::min<size_t>((size_t)ws.NdMax * 8, smem_optin - 8192 - nr_bytes);
    const int zs_cap = (int)(zs_only / 8);
    const size_t zs_bytes = nr_bytes + zs_only;
    cudaFuncSetAttribute(grad3_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zs_bytes);
    for (int img0 = 0; img0 < gp.O; img0 += slots) {
        const int cnt = std::min(slots, gp.O - img0);
        grad3_classify_kernel<Real><<<cnt, GRAD_THREADS, 0, st>>>(ws, gv, gp.u, gp.ubar, gp.alpha_maps, img0);
        grad3_assemble_kernel<<<cnt, GRAD_THREADS, 0, st>>>(ws);
        {
            cudaError_t fe = launch_factor(ws, guard, use_stage, bpltv::env_get("BPLTV_GRAD_DBG") ? atoi(bpltv::env_get("BPLTV_GRAD_DBG")) : 0, cnt, factor_cluster_size(cnt, sm_count, ws.LD), smem, st);
            if (fe != cudaSuccess) { cudaGetLastError(); return grad_fail(w, -2, std::string("factor launch failed: ") + cudaGetErrorString(fe)); }
        }
        grad3_solve_kernel<<<cnt, GRAD_THREADS, zs_bytes, st>>>(ws, gv, (double *)w.out_img, (double *)w.relres, img0, zs_cap);
        *launches += 4;
    }
    const int nout = 3 * ng;
    grad_reduce_kernel<<<1, std::max(32, (nout + 31) / 32 * 32), 0, st>>>((double *)w.out_img, (double *)w.relres, gp.O,
                                                                          nout, d_grad_out, (double *)w.relres_max);
    *launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return grad_fail(w, -2, std::string("gradient kernel launch failed: ") + cudaGetErrorString(e));
    return 0;
}

}  // namespace bpltv
