// emu_stream.cpp — TEST INFRASTRUCTURE: the one-thread-per-pixel kernels on the CPU thread emulation — kernel G
// (pdps_generic_kernel), the streaming pair of the sum-of-regularisers solve (sumregs_primal_kernel /
// sumregs_dual_kernel) and the deterministic two-stage cost reduction (cost_partial_kernel + sum_partials_kernel),
// launched as bpltv_api.cu launches them.  Built by tests/test_emu_stream.py with g++ -std=c++20
// -ffp-contract=off; never shipped.
#include "emu_cuda.h"

#include "../../bpldenoising_b200/csrc/pdps_generic.cuh"
#include "../../bpldenoising_b200/csrc/pdps_sumregs.cuh"

using namespace bpltv;

template <typename Real>
static std::vector<StepConsts<Real>> steps(int maxiter, double opnorm)
{
    std::vector<StepConsts<Real>> h(std::max(maxiter, 1));
    double sigma = (0.99 / 5) / opnorm, tau = 5.0 / opnorm;
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

// kernel G, launched like launch_generic (bpltv_api.cu): grid (ceil(M/bt), N, O)
template <typename Real>
static int generic(int M, int N, int O, int maxiter, int strict, const double *f_in, double alpha_s, const double *amap_in,
                   double *u_out)
{
    const size_t n = (size_t)M * N * O, plane = (size_t)M * N;
    std::vector<Real> x[2], y1[2], y2[2], f(n), am(plane);
    for (int b = 0; b < 2; ++b) { x[b].assign(n, 0); y1[b].assign(n, 0); y2[b].assign(n, 0); }
    for (size_t k = 0; k < n; ++k) f[k] = (Real)f_in[k];
    if (amap_in) for (size_t k = 0; k < plane; ++k) am[k] = (Real)amap_in[k];
    const auto st = steps<Real>(maxiter, std::sqrt(8.0));
    const int bt = M >= 256 ? 256 : std::max(32, (M + 31) / 32 * 32);
    int cur = 0;
    for (int it = 0; it < maxiter; ++it) {
        GenericArgs<Real> a;
        a.x_in = x[cur].data(); a.y1_in = y1[cur].data(); a.y2_in = y2[cur].data(); a.f = f.data();
        a.x_out = x[cur ^ 1].data(); a.y1_out = y1[cur ^ 1].data(); a.y2_out = y2[cur ^ 1].data();
        a.alpha_map = amap_in ? am.data() : nullptr; a.steps = st.data(); a.it = it; a.M = M; a.N = N; a.O = O;
        a.alpha_s = (Real)alpha_s; a.rho = (Real)0; a.bm = BatchMap<Real>();
        emu::launch(dim3((unsigned)((M + bt - 1) / bt), (unsigned)N, (unsigned)O), bt, [&] {
            if (amap_in) { if (strict) pdps_generic_kernel<Real, true, true>(a); else pdps_generic_kernel<Real, true, false>(a); }
            else { if (strict) pdps_generic_kernel<Real, false, true>(a); else pdps_generic_kernel<Real, false, false>(a); }
        });
        cur ^= 1;
    }
    for (size_t k = 0; k < n; ++k) u_out[k] = (double)x[cur][k];
    return 0;
}

// the streaming pair, launched like run_sumregs_pdps (bpltv_api.cu): two in-place launches per iteration
template <typename Real>
static int sumregs_stream(int M, int N, int O, int maxiter, int strict, const double *f_in, const double *alpha3,
                          const double *amap_in, double *u_out)
{
    const size_t n = (size_t)M * N * O, plane = (size_t)M * N;
    std::vector<Real> x(n, 0), xb(n, 0), y(6 * n, 0), f(n), am(3 * plane);
    for (size_t k = 0; k < n; ++k) f[k] = (Real)f_in[k];
    if (amap_in) for (size_t k = 0; k < 3 * plane; ++k) am[k] = (Real)amap_in[k];
    const auto st = steps<Real>(maxiter, std::sqrt(18.0));
    SumRegsArgs<Real> a;
    a.x = x.data(); a.xb = xb.data(); a.f = f.data(); a.y = y.data(); a.amap = amap_in ? am.data() : nullptr;
    for (int k = 0; k < 3; ++k) a.alpha[k] = (Real)(alpha3 ? alpha3[k] : 0.0);
    a.M = M; a.N = N; a.O = O;
    const unsigned grid = (unsigned)((n + 255) / 256);
    for (int it = 0; it < maxiter; ++it) {
        a.sc = st[it];
        emu::launch(dim3(grid), 256, [&] { if (strict) sumregs_primal_kernel<Real, true>(a); else sumregs_primal_kernel<Real, false>(a); });
        emu::launch(dim3(grid), 256, [&] {
            if (amap_in) { if (strict) sumregs_dual_kernel<Real, true, true>(a); else sumregs_dual_kernel<Real, true, false>(a); }
            else { if (strict) sumregs_dual_kernel<Real, false, true>(a); else sumregs_dual_kernel<Real, false, false>(a); }
        });
    }
    for (size_t k = 0; k < n; ++k) u_out[k] = (double)x[k];
    return 0;
}

extern "C" int emu_pdps_generic(int prec, int M, int N, int O, int maxiter, int strict, const double *f, double alpha_s,
                                const double *amap, double *u_out)
{
    return prec == 32 ? generic<float>(M, N, O, maxiter, strict, f, alpha_s, amap, u_out)
                      : generic<double>(M, N, O, maxiter, strict, f, alpha_s, amap, u_out);
}

extern "C" int emu_sumregs_stream(int prec, int M, int N, int O, int maxiter, int strict, const double *f,
                                  const double *alpha3, const double *amap, double *u_out)
{
    return prec == 32 ? sumregs_stream<float>(M, N, O, maxiter, strict, f, alpha3, amap, u_out)
                      : sumregs_stream<double>(M, N, O, maxiter, strict, f, alpha3, amap, u_out);
}

// cost = 0.5‖u − ū‖² as run_cost (bpltv_api.cu) launches it
extern "C" double emu_cost(const double *u, const double *ubar, long long n)
{
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>(((size_t)n + 256 * 8 - 1) / (256 * 8), 1024));
    std::vector<double> partials(1024, 0.0);
    double out = 0.0;
    emu::launch(dim3((unsigned)blocks), 256, [&] { cost_partial_kernel<double>(u, ubar, (size_t)n, partials.data()); });
    emu::launch(dim3(1), 256, [&] { sum_partials_kernel(partials.data(), blocks, 0.5, &out); });
    return out;
}

// PatchOp up-sampling (S7) and the per-image squared errors of the λ-sweeps, as bpltv_api.cu launches them
extern "C" void emu_patch_upsample(const double *lam, int lm, int ln, double *map, int M, int N)
{
    emu::launch(dim3((unsigned)((M * N + 255) / 256)), 256, [&] { patch_upsample_kernel<double>(lam, lm, ln, map, M, N); });
}

extern "C" void emu_sqerr_images(const double *u, const double *ubar, int plane, int f_mod, int V, double *out)
{
    emu::launch(dim3((unsigned)V), 256, [&] { sqerr_image_kernel<double>(u, ubar, plane, f_mod, out); });
}
