// emu_sumregs.cpp — TEST INFRASTRUCTURE: runs sumregs_resident_kernel of
// bpldenoising_b200/csrc/pdps_sumregs.cuh (one launch = the whole three-operator PDPS solve, one image per
// thread-block cluster, halo columns pushed through distributed shared memory) on the CPU thread emulation.
// Built by tests/test_emu_sumregs.py with g++ -std=c++20 -ffp-contract=off; never shipped.
#include "emu_cuda.h"

#include <cstring>

#include "../../bpldenoising_b200/csrc/pdps_sumregs.cuh"

using namespace bpltv;

// step-size recursion as upload_steps of bpltv_api.cu (S1, S2, S12)
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

template <typename Real>
static int run(int M, int N, int O, int CS, int threads, int maxiter, int strict, int init_mode, const double *f_in,
               const double *alpha3, const double *amap_in, double *u_out)
{
    const int NC = (N + CS - 1) / CS;
    if ((CS - 1) * NC >= N) return -1;
    if ((NC * M + threads - 1) / threads > 8 || threads % 32) return -2;
    const size_t n = (size_t)M * N * O, plane = (size_t)M * N;
    std::vector<Real> f(n), u(n, (Real)0), amap;
    for (size_t k = 0; k < n; ++k) f[k] = (Real)f_in[k];
    if (amap_in) { amap.resize(3 * plane); for (size_t k = 0; k < 3 * plane; ++k) amap[k] = (Real)amap_in[k]; }
    const auto st = steps<Real>(maxiter, 5.0, 0.99 / 5, std::sqrt(18.0));
    SumRegsResArgs<Real> a;
    a.f = f.data(); a.u_out = u.data(); a.amap = amap_in ? amap.data() : nullptr; a.steps = st.data();
    for (int k = 0; k < 3; ++k) a.alpha[k] = (Real)(alpha3 ? alpha3[k] : 0.0);
    a.maxiter = maxiter; a.M = M; a.N = N; a.O = O; a.init_mode = init_mode; a.NC = NC;
    const size_t smem_doubles = (sumregs_resident_plane_bytes<Real>(NC, M) + 32 + 7) / 8;      // planes + four halo mbarriers
    emu::launch(dim3((unsigned)(O * CS)), threads, [&] {
        if (amap_in) { if (strict) sumregs_resident_kernel<Real, 8, true, true>(a); else sumregs_resident_kernel<Real, 8, true, false>(a); }
        else { if (strict) sumregs_resident_kernel<Real, 8, false, true>(a); else sumregs_resident_kernel<Real, 8, false, false>(a); }
    }, smem_doubles, CS);
    for (size_t k = 0; k < n; ++k) u_out[k] = (double)u[k];
    return 0;
}

extern "C" int emu_sumregs_resident(int prec, int M, int N, int O, int CS, int threads, int maxiter, int strict,
                                    int init_mode, const double *f, const double *alpha3, const double *amap, double *u_out)
{
    return prec == 32 ? run<float>(M, N, O, CS, threads, maxiter, strict, init_mode, f, alpha3, amap, u_out)
                      : run<double>(M, N, O, CS, threads, maxiter, strict, init_mode, f, alpha3, amap, u_out);
}
