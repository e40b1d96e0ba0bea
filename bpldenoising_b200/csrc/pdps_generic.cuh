// pdps_generic.cuh — any-size fused PDPS iteration (one thread per pixel) plus the
// small helper kernels around the solve (λ-map up-sampling, precision conversion,
// cost reduction).  The generic kernel is the fallback for shapes the column-march
// kernel does not take (odd M, M·sizeof > one CTA) and an independent second
// implementation the tests cross-check against.
#pragma once
#include "common.cuh"

namespace bpltv {

template <typename Real>
struct GenericArgs {
    const Real *x_in, *y1_in, *y2_in, *f;
    Real *x_out, *y1_out, *y2_out;
    const Real *alpha_map;
    const StepConsts<Real> *steps;
    int it;
    int M, N, O;
    Real alpha_s, rho;
    BatchMap<Real> bm;
};

// x̄ at pixel (i,j) of image base pointers; also returns x_new through xn.
template <typename Real, bool STRICT>
static __device__ __forceinline__ Real xbar_at(const Real *x, const Real *f, const Real *y1, const Real *y2,
                                               int i, int j, int M, const StepConsts<Real> &sc, Real &xn)
{
    const size_t k = (size_t)j * M + i;
    const Real y1up = (i > 0) ? __ldg(y1 + k - 1) : (Real)0;
    const Real y2lf = (j > 0) ? __ldg(y2 + k - M) : (Real)0;
    Real xb;
    xn = primal_update<Real, STRICT>(__ldg(x + k), __ldg(f + k), y1up, __ldg(y1 + k), y2lf, __ldg(y2 + k), sc, xb);
    return xb;
}

template <typename Real, bool MAP, bool STRICT>
__global__ void __launch_bounds__(256) pdps_generic_kernel(const GenericArgs<Real> a)
{
    const int M = a.M, N = a.N;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int o = blockIdx.z;
    if (i >= M) return;
    const size_t img = (size_t)o * M * N;
    const Real *x = a.x_in + img, *f = a.f + (size_t)a.bm.f_image(o) * M * N, *y1 = a.y1_in + img, *y2 = a.y2_in + img;
    const StepConsts<Real> sc = a.steps[a.it];
    const size_t k = (size_t)j * M + i;

    Real xn, tmp;
    const Real xb = xbar_at<Real, STRICT>(x, f, y1, y2, i, j, M, sc, xn);
    Real d1 = 0, d2 = 0;
    if (i + 1 < M) {
        const Real xb1 = xbar_at<Real, STRICT>(x, f, y1, y2, i + 1, j, M, sc, tmp);
        d1 = STRICT ? StrictOps<Real>::sub(xb1, xb) : xb1 - xb;
    }
    if (j + 1 < N) {
        const Real xb2 = xbar_at<Real, STRICT>(x, f, y1, y2, i, j + 1, M, sc, tmp);
        d2 = STRICT ? StrictOps<Real>::sub(xb2, xb) : xb2 - xb;
    }
    Real v1 = __ldg(y1 + k), v2 = __ldg(y2 + k);
    const Real al = MAP ? __ldg(a.alpha_map + (size_t)a.bm.lam_set(o) * a.bm.map_stride + k) : a.bm.scalar(o, a.alpha_s);
    if (a.rho != (Real)0) dual_update_rho<Real, STRICT>(v1, v2, d1, d2, al, a.rho, sc);
    else dual_update<Real, STRICT, false>(v1, v2, d1, d2, al, a.rho, sc);
    a.x_out[img + k] = xn;
    a.y1_out[img + k] = v1;
    a.y2_out[img + k] = v2;
}

// PatchOp up-sampling (S7): pixel i belongs to patch floor(i*lm/M); block-constant.
// Replaces `inplace!(x̄, p, x)` (/root/reference/src/TVLearningFunctionVec.jl:58-60).
template <typename Real>
__global__ void patch_upsample_kernel(const double *lam, int lm, int ln, Real *map, int M, int N)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M * N) return;
    const int i = k % M, j = k / M;
    const int pi = (int)(((long long)i * lm) / M), pj = (int)(((long long)j * ln) / N);
    map[k] = (Real)lam[pi + (size_t)lm * pj];
}

// L parameter sets at once: set l → map + l·M·N (λ-sweeps over patch grids)
template <typename Real>
__global__ void patch_upsample_sets_kernel(const double *lam, int L, int lm, int ln, Real *map, int M, int N)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t plane = (size_t)M * N;
    if (k >= plane * L) return;
    const int l = (int)(k / plane);
    const int q = (int)(k - (size_t)l * plane);
    const int i = q % M, j = q / M;
    const int pi = (int)(((long long)i * lm) / M), pj = (int)(((long long)j * ln) / N);
    map[k] = (Real)lam[(size_t)l * lm * ln + pi + (size_t)lm * pj];
}

// Σ (u_v - ū_{v % f_mod})² of virtual image v = blockIdx.x (deterministic per image)
template <typename Real>
__global__ void __launch_bounds__(256) sqerr_image_kernel(const Real *u, const Real *ubar, int plane, int f_mod,
                                                         double *out)
{
    __shared__ double smem[32];
    const int v = blockIdx.x;
    const Real *uv = u + (size_t)v * plane;
    const Real *ub = ubar + (size_t)(f_mod ? v % f_mod : v) * plane;
    double acc = 0.0;
    for (int k = threadIdx.x; k < plane; k += blockDim.x) {
        const double d = (double)uv[k] - (double)ub[k];
        acc = fma(d, d, acc);
    }
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) out[v] = s;
}

template <typename Dst, typename Src>
__global__ void convert_kernel(const Src *src, Dst *dst, size_t n)
{
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) dst[k] = (Dst)src[k];
}

template <typename Real>
__global__ void fill_kernel(Real *dst, Real v, size_t n)
{
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) dst[k] = v;
}

// cost = 0.5‖u-ū‖² (/root/reference/src/TVLearningFunctionVec.jl:20), deterministic
// two-stage reduction: per-block partial sums in a fixed layout, then one block adds
// them in a fixed order.  Accumulates in double in both precisions.
template <typename Real>
__global__ void __launch_bounds__(256) cost_partial_kernel(const Real *u, const Real *ubar, size_t n,
                                                          double *partials)
{
    __shared__ double smem[32];
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const double d = (double)u[k] - (double)ubar[k];
        acc = fma(d, d, acc);
    }
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// out[0] = scale * Σ partials[0..n)
__global__ void __launch_bounds__(256) sum_partials_kernel(const double *partials, int n, double scale,
                                                          double *out)
{
    __shared__ double smem[32];
    double acc = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) acc += partials[k];
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) out[0] = scale * s;
}

}  // namespace bpltv
