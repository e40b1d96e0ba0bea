// emu_lu.cpp — TEST INFRASTRUCTURE: runs lu3_classify → lu3_assemble → lu_factor → lu3_solve of
// bpldenoising_b200/csrc/lu_band.cuh for ONE image on the CPU thread emulation (emu_cuda.h).
// Built by tests/test_emu_lu.py with g++ -std=c++20; never shipped.
#include "emu_cuda.h"

#include <cstring>

#include "../../bpldenoising_b200/csrc/lu_band.cuh"

using namespace bpltv;

extern "C" int emu_lu_gradient(int use_pin, int csize, int nops, int n, const double *u, const double *ubar, const double *alpha_maps, const double *alpha3,
                               double gamma, int lm, int ln, int refine, int vec_in_smem, double *out, double *relres,
                               int *pivot_flag, double *band_out /* N·LD or NULL: the assembled (unfactored) band */,
                               int *ld_out)
{
    const int N = n * n;
    LuSlots ws;
    ws.n = n; ws.N = N; ws.nops = nops; ws.use_pin = (use_pin && 2 * (((nops == 1 ? n : 2 * n) + 1) & ~1) <= LU_THREADS) ? 1 : 0; ws.bw = ((nops == 1 ? n : 2 * n) + 1) & ~1; /* = lu_band_halfwidth (gradient_lu.cuh) */ ws.bwx = ws.bw + LU_NB; ws.LD = 2 * ws.bwx + 1;
    ws.ab_stride = ((size_t)N * ws.LD + 3) & ~(size_t)3; ws.pix_stride = (size_t)LU_PLANES * N;
    if (ld_out) *ld_out = ws.LD;
    std::vector<double> ab_store(ws.ab_stride + 2, 0.0), pix(ws.pix_stride, 0.0);
    int info[4] = {0, 0, 0, 0};
    double *abp = ab_store.data(); if (reinterpret_cast<std::uintptr_t>(abp) & 15) ++abp;   // 16-byte aligned like cudaMalloc
    ws.ab = abp; ws.pix = pix.data(); ws.info = info;
    Lu3Params pr;
    for (int k = 0; k < 3; ++k) pr.alpha[k] = alpha3 ? alpha3[k] : 0.0;
    pr.gamma = gamma; pr.lm = lm; pr.ln = ln; pr.refine = refine;
    const int chunks = std::max(1, std::min(64, (N + 255) / 256));
    emu::launch(dim3(1, chunks), 256, [&] { lu3_classify_kernel<double>(ws, gamma, u, ubar, 0); });
    emu::launch(dim3(1, chunks), 256, [&] { lu3_assemble_kernel<double>(ws, pr, alpha_maps); });
    if (band_out) std::memcpy(band_out, abp, ws.ab_stride * sizeof(double));
    if (csize > 1) emu::launch(dim3(csize), LU_THREADS, [&] { lu_factor_kernel<true>(ws); }, lu_factor_smem(ws.bw, ws.use_pin != 0) / 8, csize);
    else emu::launch(dim3(1), LU_THREADS, [&] { lu_factor_kernel<false>(ws); }, lu_factor_smem(ws.bw, ws.use_pin != 0) / 8);
    std::vector<double> out_img(nops * lm * ln, 0.0);
    double rr = -1.0;
    emu::launch(dim3(1), LU_THREADS, [&] { lu3_solve_kernel<double>(ws, pr, alpha_maps, out_img.data(), &rr, 0, vec_in_smem); },
                lu_solve_smem(N, vec_in_smem != 0) / 8);
    std::memcpy(out, out_img.data(), out_img.size() * sizeof(double));
    *relres = rr;
    *pivot_flag = info[0];
    return 0;
}
