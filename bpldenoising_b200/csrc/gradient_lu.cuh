// gradient_lu.cuh — host driver of the node-space band LU (kernels: lu_band.cuh).
//
// Solves (I + Σ_k diag(α_k) G_kᵀ(B_k − C_k)G_k) p = ū − u per image and forms g_k = p ⊙ G_kᵀ w_k, i.e.
//   nops = 3  sumregs_gradient_reg, patch (/root/reference/src/SumRegsLearningFunction.jl:195-262, row-scaled,
//             non-symmetric) and scalar (:112-167);
//   nops = 1  gradient_reg of the TV learning function, scalar (/root/reference/src/TVLearningFunctionVec.jl:137-161)
//             and patch (:192-215, `α[:] .* Gᵀ(B−C)G`, row-scaled as written), forward differences only,
//             half-bandwidth n.
#pragma once
#include "env_switches.h"
#include "gradient.cuh"
#include "lu_band.cuh"

namespace bpltv {

// half-bandwidth of the node-space system: n (forward differences alone) or 2n (with the centred ones), rounded up
// to EVEN — the 16-byte alignment of the kernels' vector accesses rests on it (n ≥ 4, so it stays below n² − 1)
static inline int lu_band_halfwidth(int nops, int n) { return ((nops == 1 ? n : 2 * n) + 1) & ~1; }

template <typename Real>
struct LuProblem {
    const Real *u, *ubar;
    int M, N, O;
    double alpha[3];
    const Real *alpha_maps;   // nops maps of M·N or nullptr
    int lm, ln, nops;
    double gamma;
};

// ---------------------------------------------------------------------------
// host driver of the node-space band LU
// ---------------------------------------------------------------------------
// CTAs per image of the LU factorisation: as many (power of two, ≤ 4: measured no gain beyond — the block step is
// then its panel chain) as leave every image of the wave its own cluster; one when the trailing window has too few 32×32 tiles to share.  BPLTV_LU_CLUSTER overrides.
static inline int lu_cluster_ctas(int images_in_wave, int sm_count, int bw)
{
    const char *env = bpltv::env_get("BPLTV_LU_CLUSTER");
    if (env && *env) { const int c = atoi(env); if (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) return c; }
    if (bw < 200) return 1;
    int C = 1;
    while (C < 4 && 2 * C * images_in_wave <= sm_count) C *= 2;
    return C;
}

static inline cudaError_t launch_lu_factor(const LuSlots &ws, int cnt, int C, size_t smem, cudaStream_t st)
{
    while (C > 1) {
        cudaError_t e = cudaFuncSetAttribute(lu_factor_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess && C > 8)
            e = cudaFuncSetAttribute(lu_factor_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(cnt * C));
        cfg.blockDim = dim3(LU_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nclusters = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&nclusters, lu_factor_kernel<true>, &cfg);
        if (e == cudaSuccess && nclusters >= 1 && (C <= 8 || nclusters >= cnt)) {
            e = cudaLaunchKernelEx(&cfg, lu_factor_kernel<true>, ws);
            if (e == cudaSuccess) return e;
        }
        cudaGetLastError();
        C /= 2;      // not schedulable as asked: smaller clusters, finally the single-CTA kernel
    }
    cudaError_t e = cudaFuncSetAttribute(lu_factor_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lu_factor_kernel<false><<<cnt, LU_THREADS, smem, st>>>(ws);
    return cudaGetLastError();
}

template <typename Real>
static int run_gradient_lu(GradWork &w, const LuProblem<Real> &gp, int sm_count, size_t smem_optin, cudaStream_t st,
                            double *d_grad_out, long long *launches)
{
    const int n = gp.M, N = gp.M * gp.N, ng = gp.lm * gp.ln, nops = gp.nops;
    if (gp.M != gp.N) return grad_fail(w, -1, "square images required");
    if (nops * ng > 1024) return grad_fail(w, -1, "lambda grid larger than 1024 entries is not supported");
    if (n < 4) return grad_fail(w, -1, "images smaller than 4x4 are not supported by the band LU");
    LuSlots ws;
    ws.n = n; ws.N = N; ws.nops = nops; ws.bw = lu_band_halfwidth(nops, n); ws.bwx = ws.bw + LU_NB; ws.LD = 2 * ws.bwx + 1;
    ws.ab_stride = ((size_t)N * ws.LD + 3) & ~(size_t)3;
    ws.pix_stride = (size_t)LU_PLANES * N;
    ws.use_pin = (2 * ws.bw <= LU_THREADS && lu_factor_smem(ws.bw, true) <= smem_optin) ? 1 : 0;
    const size_t fsmem = lu_factor_smem(ws.bw, ws.use_pin != 0);
    if (fsmem > smem_optin) return grad_fail(w, -1, "image too large for the band-LU panels in shared memory (n <= 430)");
    const bool vec_in_smem = lu_solve_smem(N, true) + 1024 <= smem_optin;
    const size_t ssmem = lu_solve_smem(N, vec_in_smem);

    const size_t per_slot = (ws.ab_stride + ws.pix_stride) * 8 + 16;
    // workspace: kept across calls; the driver is asked about free memory only when it has to grow
    const size_t key = (size_t)N * 4096 + (size_t)ws.LD;      // image size and band pitch (TV: n+…, sum of regularisers: 2n+…)
    const int want = std::min(gp.O, sm_count);
    int slots;
    if (w.lu_cap_N == key && w.lu_cap_slots >= (size_t)want) {
        slots = want;
    } else {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t have = w.lu_cap_N == key ? w.lu_cap_slots : 0;
        const size_t budget = (free_b + have * per_slot) / 2;
        slots = (int)std::min<size_t>((size_t)want, std::max<size_t>(1, budget / per_slot));
        if (w.lu_cap_N != key || w.lu_cap_slots < (size_t)slots) {
            void **all[] = {&w.lu_ab, &w.lu_pix, &w.lu_info};
            for (void **p : all) { if (*p) cudaFree(*p); *p = nullptr; }
            cudaError_t e = cudaMalloc(&w.lu_ab, ws.ab_stride * 8 * slots);
            if (e == cudaSuccess) e = cudaMalloc(&w.lu_pix, ws.pix_stride * 8 * slots);
            if (e == cudaSuccess) e = cudaMalloc(&w.lu_info, 16 * (size_t)slots);
            if (e != cudaSuccess) {
                cudaGetLastError();
                w.lu_cap_slots = 0; w.lu_cap_N = 0;
                return grad_fail(w, -6, std::string("band-LU workspace allocation failed: ") + cudaGetErrorString(e));
            }
            w.lu_cap_slots = slots; w.lu_cap_N = key;
        } else {
            slots = (int)std::min<size_t>(w.lu_cap_slots, (size_t)want);
        }
    }
    if (w.cap_O < (size_t)gp.O || w.cap_ng < (size_t)(nops * ng)) {
        if (w.out_img) cudaFree(w.out_img);
        if (w.relres) cudaFree(w.relres);
        if (!w.relres_max) cudaMalloc(&w.relres_max, 8);
        cudaError_t e = cudaMalloc(&w.out_img, (size_t)gp.O * nops * ng * 8);
        if (e == cudaSuccess) e = cudaMalloc(&w.relres, (size_t)gp.O * 8);
        if (e != cudaSuccess) { cudaGetLastError(); w.cap_O = 0; return grad_fail(w, -6, "gradient output allocation failed"); }
        w.cap_O = gp.O; w.cap_ng = nops * ng;
    }
    ws.ab = (double *)w.lu_ab; ws.pix = (double *)w.lu_pix; ws.info = (int *)w.lu_info;

    Lu3Params pr;
    for (int k = 0; k < 3; ++k) pr.alpha[k] = gp.alpha[k];
    pr.gamma = gp.gamma; pr.lm = gp.lm; pr.ln = gp.ln;
    pr.refine = 3;
    cudaError_t e = cudaFuncSetAttribute(lu3_solve_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem);
    if (e != cudaSuccess) { cudaGetLastError(); return grad_fail(w, -2, std::string("band-LU kernel attributes: ") + cudaGetErrorString(e)); }
    const int chunks = std::max(1, std::min(64, (N + 255) / 256));
    for (int img0 = 0; img0 < gp.O; img0 += slots) {
        const int cnt = std::min(slots, gp.O - img0);
        cudaMemsetAsync(ws.ab, 0, ws.ab_stride * 8 * cnt, st);
        lu3_classify_kernel<Real><<<dim3(cnt, chunks), 256, 0, st>>>(ws, pr.gamma, gp.u, gp.ubar, img0);
        lu3_assemble_kernel<Real><<<dim3(cnt, chunks), 256, 0, st>>>(ws, pr, gp.alpha_maps);
        e = launch_lu_factor(ws, cnt, lu_cluster_ctas(cnt, sm_count, ws.bw), fsmem, st);
        if (e != cudaSuccess) { cudaGetLastError(); return grad_fail(w, -2, std::string("band-LU factor launch failed: ") + cudaGetErrorString(e)); }
        lu3_solve_kernel<Real><<<cnt, LU_THREADS, ssmem, st>>>(ws, pr, gp.alpha_maps, (double *)w.out_img,
                                                              (double *)w.relres, img0, vec_in_smem ? 1 : 0);
        *launches += 4;
    }
    const int nout = nops * ng;
    grad_reduce_kernel<<<1, std::max(32, (nout + 31) / 32 * 32), 0, st>>>((double *)w.out_img, (double *)w.relres, gp.O,
                                                                          nout, d_grad_out, (double *)w.relres_max);
    *launches += 1;
    e = cudaGetLastError();
    if (e != cudaSuccess) return grad_fail(w, -2, std::string("band-LU kernel launch failed: ") + cudaGetErrorString(e));
    return 0;
}

}  // namespace bpltv
