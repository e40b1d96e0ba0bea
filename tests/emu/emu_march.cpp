// emu_march.cpp — TEST INFRASTRUCTURE: runs pdps_march_kernel of bpldenoising_b200/csrc/pdps_march.cuh (kernel A:
// one launch = one fused PDPS iteration over the stack; column ranges per CTA, previous column carried in
// registers, row neighbours by warp shuffle and one shared-memory slot per warp boundary, ping-pong state) on the
// CPU thread emulation.  Built by tests/test_emu_march.py with g++ -std=c++20 -ffp-contract=off; never shipped.
#include "emu_cuda.h"

#include "../../bpldenoising_b200/csrc/pdps_march.cuh"

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

template <typename Real, int VEC>
static int run(int M, int N, int O, int grid, int threads, int maxiter, int strict, const double *f_in, double alpha_s,
               const double *amap_in, double *u_out)
{
    if (M % VEC || threads * VEC < M || threads % 32) return -1;
    const size_t n = (size_t)M * N * O, plane = (size_t)M * N;
    // 16-byte aligned state (the vector accesses of VecIO), two copies each
    std::vector<Real> store(7 * (n + 4) + plane + 8, (Real)0);
    auto al16 = [](Real *p) { while (reinterpret_cast<std::uintptr_t>(p) & 15) ++p; return p; };
    Real *x[2], *y1[2], *y2[2], *f, *am;
    Real *p = store.data();
    for (int b = 0; b < 2; ++b) { x[b] = al16(p); p = x[b] + n; y1[b] = al16(p); p = y1[b] + n; y2[b] = al16(p); p = y2[b] + n; }
    f = al16(p); p = f + n; am = al16(p);
    for (size_t k = 0; k < n; ++k) f[k] = (Real)f_in[k];
    if (amap_in) for (size_t k = 0; k < plane; ++k) am[k] = (Real)amap_in[k];
    const auto st = steps<Real>(maxiter, 5.0, 0.99 / 5, std::sqrt(8.0));
    int cur = 0;
    for (int it = 0; it < maxiter; ++it) {
        MarchArgs<Real> a;
        a.x_in = x[cur]; a.y1_in = y1[cur]; a.y2_in = y2[cur]; a.f = f;
        a.x_out = x[cur ^ 1]; a.y1_out = y1[cur ^ 1]; a.y2_out = y2[cur ^ 1];
        a.alpha_map = amap_in ? am : nullptr; a.sc = st[it]; a.M = M; a.N = N; a.O = O;
        a.total_cols = (long long)N * O; a.prefetch_dist = 0; a.alpha_s = (Real)alpha_s; a.rho = (Real)0;
        a.bm = BatchMap<Real>();
        emu::launch(dim3((unsigned)grid), threads, [&] {
            if (amap_in) { if (strict) pdps_march_kernel<Real, VEC, true, true, 256, 2>(a); else pdps_march_kernel<Real, VEC, true, false, 256, 2>(a); }
            else { if (strict) pdps_march_kernel<Real, VEC, false, true, 256, 2>(a); else pdps_march_kernel<Real, VEC, false, false, 256, 2>(a); }
        });
        cur ^= 1;
    }
    for (size_t k = 0; k < n; ++k) u_out[k] = (double)x[cur][k];
    return 0;
}

extern "C" int emu_pdps_march(int prec, int vec, int M, int N, int O, int grid, int threads, int maxiter, int strict,
                              const double *f, double alpha_s, const double *amap, double *u_out)
{
    if (prec == 32) {
        if (vec == 4) return run<float, 4>(M, N, O, grid, threads, maxiter, strict, f, alpha_s, amap, u_out);
        if (vec == 2) return run<float, 2>(M, N, O, grid, threads, maxiter, strict, f, alpha_s, amap, u_out);
        return run<float, 1>(M, N, O, grid, threads, maxiter, strict, f, alpha_s, amap, u_out);
    }
    if (vec == 2) return run<double, 2>(M, N, O, grid, threads, maxiter, strict, f, alpha_s, amap, u_out);
    return run<double, 1>(M, N, O, grid, threads, maxiter, strict, f, alpha_s, amap, u_out);
}
