// emu_tblock.cpp — TEST INFRASTRUCTURE: runs pdps_tblock_kernel of bpldenoising_b200/csrc/pdps_tblock.cuh (kernel C, the
// headline kernel: T iterations per pass, software-pipelined along the column march in registers, stage 0 fed by a
// shared-memory ring that the TMA engine fills) on the CPU thread emulation.  Bulk copies are memcpys by the issuing
// thread and mbarriers atomic words (see the BPLTV_EMU block of the header), so the ring protocol — issue TB_PF columns
// ahead, wait on the slot's phase, re-read f of earlier columns — runs as written.  Built by tests/test_emu_tblock.py
// with g++ -std=c++20 -ffp-contract=off; never shipped.
#include "emu_cuda.h"

#include "../../bpldenoising_b200/csrc/pdps_tblock.cuh"

using namespace bpltv;

template <typename Real>
static std::vector<StepConsts<Real>> steps(int maxiter, double tau0, double sigma0, double opnorm)
{
    std::vector<StepConsts<Real>> h(std::max(maxiter, 1));
    double sigma = sigma0 / opnorm, tau = tau0 / opnorm;
    for (int k = 0; k < maxiter; ++k) {
        const double omega = 1.0 / std::sqrt(1.0 + 2.0 * tau);
        StepConsts<Real> s;
        s.tau = (Real)tau; s.sigma = (Real)sigma; s.omega = (Real)omega;
        s.one_p_tau = (Real)1 + s.tau;
        s.one_p_omega = (Real)1 + s.omega;
        s.inv_one_p_tau = (Real)(1.0 / (1.0 + tau));
        s.tau_over_one_p_tau = (Real)(tau / (1.0 + tau));
        s.rcp_one_p_tau = (Real)1 / s.one_p_tau;
        h[k] = s;
        tau = tau * omega; sigma = sigma / omega;
    }
    return h;
}

// maxiter must be a multiple of T (the library finishes a remainder with kernel A)
template <typename Real, int VEC, int T, bool RING>
static int run(int M, int N, int O, int grid, int maxiter, int strict, const double *f_in, double alpha_s, const double *amap_in,
               double *u_out)
{
    if (M % VEC || maxiter % T) return -1;
    if (RING && (M * sizeof(Real)) % 16) return -2;
    const int threads = (M / VEC + 31) / 32 * 32;
    const size_t n = (size_t)M * N * O, plane = (size_t)M * N;
    std::vector<Real> store(7 * (n + 4) + plane + 8, (Real)0);
    auto al16 = [](Real *p) { while (reinterpret_cast<std::uintptr_t>(p) & 15) ++p; return p; };
    Real *x[2], *y1[2], *y2[2], *f, *am;
    Real *p = store.data();
    for (int b = 0; b < 2; ++b) { x[b] = al16(p); p = x[b] + n; y1[b] = al16(p); p = y1[b] + n; y2[b] = al16(p); p = y2[b] + n; }
    f = al16(p); p = f + n; am = al16(p);
    for (size_t k = 0; k < n; ++k) f[k] = (Real)f_in[k];
    if (amap_in) for (size_t k = 0; k < plane; ++k) am[k] = (Real)amap_in[k];
    const auto st = steps<Real>(maxiter, 5.0, 0.99 / 5, std::sqrt(8.0));
    const size_t smem_doubles = RING ? (tblock_ring_bytes<Real, T>(M) + 7) / 8 : 0;
    int cur = 0;
    for (int it = 0; it < maxiter; it += T) {
        TBlockArgs<Real, T> a;
        a.x_in = x[cur]; a.y1_in = y1[cur]; a.y2_in = y2[cur]; a.f = f;
        a.x_out = x[cur ^ 1]; a.y1_out = y1[cur ^ 1]; a.y2_out = y2[cur ^ 1];
        a.alpha_map = amap_in ? am : nullptr; a.M = M; a.N = N; a.O = O; a.total_cols = (long long)N * O;
        a.alpha_s = (Real)alpha_s; a.bm = BatchMap<Real>();
        for (int s = 0; s < T; ++s) a.sc[s] = st[it + s];
        emu::launch(dim3((unsigned)grid), threads, [&] {
            if (amap_in) { if (strict) pdps_tblock_kernel<Real, VEC, T, true, true, RING, false, 256, 1>(a); else pdps_tblock_kernel<Real, VEC, T, true, false, RING, false, 256, 1>(a); }
            else { if (strict) pdps_tblock_kernel<Real, VEC, T, false, true, RING, false, 256, 1>(a); else pdps_tblock_kernel<Real, VEC, T, false, false, RING, false, 256, 1>(a); }
        }, smem_doubles);
        cur ^= 1;
    }
    for (size_t k = 0; k < n; ++k) u_out[k] = (double)x[cur][k];
    return 0;
}

template <typename Real, int VEC, bool RING>
static int run_t(int T, int M, int N, int O, int grid, int maxiter, int strict, const double *f, double alpha_s, const double *amap, double *u)
{
    if (T == 2) return run<Real, VEC, 2, RING>(M, N, O, grid, maxiter, strict, f, alpha_s, amap, u);
    if (T == 3) return run<Real, VEC, 3, RING>(M, N, O, grid, maxiter, strict, f, alpha_s, amap, u);
    if (T == 4) return run<Real, VEC, 4, RING>(M, N, O, grid, maxiter, strict, f, alpha_s, amap, u);
    return -3;
}

// vec: rows per thread; the 16-byte widths (fp64: 2, fp32: 4) are the ring kernels, the narrower ones load directly
extern "C" int emu_pdps_tblock(int prec, int vec, int T, int M, int N, int O, int grid, int maxiter, int strict, const double *f,
                               double alpha_s, const double *amap, double *u_out)
{
    if (prec == 32) {
        if (vec == 4) return run_t<float, 4, true>(T, M, N, O, grid, maxiter, strict, f, alpha_s, amap, u_out);
        if (vec == 2) return run_t<float, 2, false>(T, M, N, O, grid, maxiter, strict, f, alpha_s, amap, u_out);
        return run_t<float, 1, false>(T, M, N, O, grid, maxiter, strict, f, alpha_s, amap, u_out);
    }
    if (vec == 2) return run_t<double, 2, true>(T, M, N, O, grid, maxiter, strict, f, alpha_s, amap, u_out);
    return run_t<double, 1, false>(T, M, N, O, grid, maxiter, strict, f, alpha_s, amap, u_out);
}
