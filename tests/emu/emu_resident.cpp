// emu_resident.cpp — TEST INFRASTRUCTURE: runs pdps_resident_kernel of bpldenoising_b200/csrc/pdps_resident.cuh
// (kernel B: the whole TV solve in one launch, one image per thread-block cluster, x and f in registers, the duals
// and x̄ in shared memory, boundary columns pushed through distributed shared memory, row neighbours by warp
// shuffle) on the CPU thread emulation.  Built by tests/test_emu_resident.py with g++ -std=c++20
// -ffp-contract=off; never shipped.
#include "emu_cuda.h"

#include "../../bpldenoising_b200/csrc/pdps_resident.cuh"

using namespace bpltv;

// step-size recursion as upload_steps of bpltv_api.cu (S1, S2)
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

static int g_async = 1;
extern "C" void emu_resident_set_async(int on) { g_async = on; }

template <typename Real>
static int run(int M, int N, int O, int CS, int threads, int maxiter, int strict, const double *f_in, double alpha_s,
               const double *amap_in, double *u_out)
{
    if (M < 2 || (M & 1) || threads % 32 || threads % (M / 2)) return -1;
    const int CG = threads / (M / 2);
    const int NC = (N + CS - 1) / CS;
    if ((CS - 1) * NC >= N) return -2;
    if ((NC + CG - 1) / CG > 4) return -3;
    const size_t n = (size_t)M * N * O, plane = (size_t)M * N;
    std::vector<Real> f(n), u(n, (Real)0), amap;
    for (size_t k = 0; k < n; ++k) f[k] = (Real)f_in[k];
    if (amap_in) { amap.resize(plane); for (size_t k = 0; k < plane; ++k) amap[k] = (Real)amap_in[k]; }
    const auto st = steps<Real>(maxiter, 5.0, 0.99 / 5, std::sqrt(8.0));
    ResidentArgs<Real> a;
    a.f = f.data(); a.u_out = u.data(); a.alpha_map = amap_in ? amap.data() : nullptr; a.steps = st.data();
    a.maxiter = maxiter; a.M = M; a.N = N; a.O = O; a.init_mode = 0; a.NC = NC; a.alpha_s = (Real)alpha_s;
    a.bm = BatchMap<Real>();
    const size_t smem_doubles = (resident_plane_bytes<Real>(NC, M) + 16 + 7) / 8;      // planes + the two halo mbarriers
    if (g_async) {      // halo columns by (emulated) st.async + mbarrier, CTA barriers between the phases
        emu::launch(dim3((unsigned)(O * CS)), threads, [&] {
            if (amap_in) { if (strict) pdps_resident_kernel<Real, 4, true, true, true>(a); else pdps_resident_kernel<Real, 4, true, false, true>(a); }
            else { if (strict) pdps_resident_kernel<Real, 4, false, true, true>(a); else pdps_resident_kernel<Real, 4, false, false, true>(a); }
        }, smem_doubles, CS);
    } else {            // plain DSMEM stores + two cluster barriers per iteration
        emu::launch(dim3((unsigned)(O * CS)), threads, [&] {
            if (amap_in) { if (strict) pdps_resident_kernel<Real, 4, true, true, false>(a); else pdps_resident_kernel<Real, 4, true, false, false>(a); }
            else { if (strict) pdps_resident_kernel<Real, 4, false, true, false>(a); else pdps_resident_kernel<Real, 4, false, false, false>(a); }
        }, smem_doubles, CS);
    }
    for (size_t k = 0; k < n; ++k) u_out[k] = (double)u[k];
    return 0;
}

// the temporally blocked variant (pdps_resident_tb_kernel): threads = tpc · ceil((NC+4)/KC) rounded up to a warp
template <typename Real>
static int run_tb(int M, int N, int O, int CS, int KC, int maxiter, int strict, const double *f_in, double alpha_s,
                  const double *amap_in, double *u_out)
{
    if (M < 2 || (M & 1)) return -1;
    const int NC = (N + CS - 1) / CS, E = NC + 4, tpc = M / 2;
    if (NC < 2 || (CS - 1) * NC >= N) return -2;
    const int threads = (tpc * ((E + KC - 1) / KC) + 31) & ~31;
    const size_t n = (size_t)M * N * O, plane = (size_t)M * N;
    std::vector<Real> f(n), u(n, (Real)0), amap;
    for (size_t k = 0; k < n; ++k) f[k] = (Real)f_in[k];
    if (amap_in) { amap.resize(plane); for (size_t k = 0; k < plane; ++k) amap[k] = (Real)amap_in[k]; }
    const auto st = steps<Real>(maxiter, 5.0, 0.99 / 5, std::sqrt(8.0));
    ResidentArgs<Real> a;
    a.f = f.data(); a.u_out = u.data(); a.alpha_map = amap_in ? amap.data() : nullptr; a.steps = st.data();
    a.maxiter = maxiter; a.M = M; a.N = N; a.O = O; a.init_mode = 0; a.NC = NC; a.alpha_s = (Real)alpha_s;
    a.bm = BatchMap<Real>();
    const size_t smem_doubles = ((size_t)(3 * E + 4) * M * sizeof(Real) + 7) / 8;
    auto body = [&] {
        if (KC == 1) {
            if (amap_in) { if (strict) pdps_resident_tb_kernel<Real, 1, true, true>(a); else pdps_resident_tb_kernel<Real, 1, true, false>(a); }
            else { if (strict) pdps_resident_tb_kernel<Real, 1, false, true>(a); else pdps_resident_tb_kernel<Real, 1, false, false>(a); }
        } else if (KC == 2) {
            if (amap_in) { if (strict) pdps_resident_tb_kernel<Real, 2, true, true>(a); else pdps_resident_tb_kernel<Real, 2, true, false>(a); }
            else { if (strict) pdps_resident_tb_kernel<Real, 2, false, true>(a); else pdps_resident_tb_kernel<Real, 2, false, false>(a); }
        } else {
            if (amap_in) { if (strict) pdps_resident_tb_kernel<Real, 4, true, true>(a); else pdps_resident_tb_kernel<Real, 4, true, false>(a); }
            else { if (strict) pdps_resident_tb_kernel<Real, 4, false, true>(a); else pdps_resident_tb_kernel<Real, 4, false, false>(a); }
        }
    };
    emu::launch(dim3((unsigned)(O * CS)), threads, body, smem_doubles, CS);
    for (size_t k = 0; k < n; ++k) u_out[k] = (double)u[k];
    return 0;
}

extern "C" int emu_pdps_resident_tb(int prec, int M, int N, int O, int CS, int KC, int maxiter, int strict,
                                    const double *f, double alpha_s, const double *amap, double *u_out)
{
    return prec == 32 ? run_tb<float>(M, N, O, CS, KC, maxiter, strict, f, alpha_s, amap, u_out)
                      : run_tb<double>(M, N, O, CS, KC, maxiter, strict, f, alpha_s, amap, u_out);
}

extern "C" int emu_pdps_resident(int prec, int M, int N, int O, int CS, int threads, int maxiter, int strict,
                                 const double *f, double alpha_s, const double *amap, double *u_out)
{
    return prec == 32 ? run<float>(M, N, O, CS, threads, maxiter, strict, f, alpha_s, amap, u_out)
                      : run<double>(M, N, O, CS, threads, maxiter, strict, f, alpha_s, amap, u_out);
}
