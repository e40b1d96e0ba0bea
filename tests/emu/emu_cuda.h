// emu_cuda.h — TEST INFRASTRUCTURE: just enough of the CUDA execution model on OS threads to run the
// device code of bpldenoising_b200/csrc/lu_band.cuh on a CPU (one std::thread per CUDA thread, CTA and
// warp barriers, warp shuffles through a per-warp exchange buffer).  It checks index arithmetic and
// barrier placement where no GPU exists; it is never part of the product (nothing under
// bpldenoising_b200/ includes it) and says nothing about performance.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct double2 { double x, y; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
inline float2 make_float2(float a, float b) { return float2{a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline double2 make_double2(double a, double b) { return double2{a, b}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
// static shared arrays: one copy per kernel instantiation, shared by the OS threads of the CTA that is running
// (non-cluster launches run their CTAs one after the other); dynamic shared memory goes through emu::dyn_smem()
#define __shared__ static
#define __align__(x)
#define __noinline__

#include <cstdint>
#include <cstdio>
#include <cstdlib>
namespace emu {
inline void check_aligned16(const void *p)
{
    if (reinterpret_cast<std::uintptr_t>(p) & 15) { std::fprintf(stderr, "emu: misaligned 16-byte access %p\n", p); std::abort(); }
}
struct Warp {
    std::barrier<> bar;
    double xch[32];
    Warp() : bar(32) {}
};
struct Cta;
struct Cluster {
    std::barrier<> bar;
    int size;
    std::vector<Cta *> ctas;       // by rank
    Cluster(int ctas_, int threads) : bar((std::ptrdiff_t)ctas_ * threads), size(ctas_) {}
};
struct Cta {
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    std::vector<double> smem;      // dynamic shared memory of this CTA
    Cluster *cluster = nullptr;
    int rank = 0;
    Cta(int threads, size_t smem_doubles) : bar(threads), smem(smem_doubles + 2, 0.0)
    {
        for (int w = 0; w < (threads + 31) / 32; ++w) warps.emplace_back(new Warp());
    }
};
inline thread_local Cta *cta = nullptr;
inline thread_local Warp *warp = nullptr;
inline double *dyn_smem()
{
    double *p = cta->smem.data();
    return (reinterpret_cast<std::uintptr_t>(p) & 15) ? p + 1 : p;     // 16-byte aligned like the GPU's
}
inline double *dyn_smem_of(Cta *c)
{
    double *p = c->smem.data();
    return (reinterpret_cast<std::uintptr_t>(p) & 15) ? p + 1 : p;
}
inline int cluster_rank() { return cta->rank; }
inline int cluster_size() { return cta->cluster ? cta->cluster->size : 1; }
inline void cluster_sync() { if (cta->cluster) cta->cluster->bar.arrive_and_wait(); else cta->bar.arrive_and_wait(); }
}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

inline void __syncthreads() { emu::cta->bar.arrive_and_wait(); }
inline void __syncwarp() { emu::warp->bar.arrive_and_wait(); }
// warp shuffles / votes through the per-warp exchange buffer (all 32 lanes take part, as on the GPU)
template <typename T>
inline T emu_shfl(T v, int src)
{
    static_assert(sizeof(T) <= sizeof(double), "exchange slot");
    emu::Warp &w = *emu::warp;
    std::memcpy(&w.xch[threadIdx.x & 31], &v, sizeof(T));
    w.bar.arrive_and_wait();
    T r;
    std::memcpy(&r, &w.xch[src & 31], sizeof(T));
    w.bar.arrive_and_wait();
    return r;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, src); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int o) { return emu_shfl(v, (int)(threadIdx.x & 31) ^ o); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
    const int lane = (int)(threadIdx.x & 31);
    return emu_shfl(v, lane >= (int)d ? lane - (int)d : lane);
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d)
{
    const int lane = (int)(threadIdx.x & 31);
    return emu_shfl(v, lane + (int)d < 32 ? lane + (int)d : lane);
}
inline int __any_sync(unsigned, int pred)
{
    int any = 0;
    for (int l = 0; l < 32; ++l) any |= emu_shfl(pred ? 1 : 0, l);      // 32 exchanges: simple, and only used at set-up
    return any;
}

using std::max;
using std::min;

// one correctly rounded IEEE operation each (compile the harness with -ffp-contract=off), like the intrinsics
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fmaf_rn(float a, float b, float c) { return std::fma(a, b, c); }
inline int __double2hiint(double a) { unsigned long long u; std::memcpy(&u, &a, 8); return (int)(u >> 32); }
inline int __double2loint(double a) { unsigned long long u; std::memcpy(&u, &a, 8); return (int)(u & 0xffffffffu); }
inline unsigned __float_as_uint(float a) { unsigned u; std::memcpy(&u, &a, 4); return u; }
// stand-ins for the hardware's reciprocal-square-root seeds, so the emulated kernels run BallScale's chain for real.
// fp64 (MUFU.RSQ64H writes the upper word only): the exact value cut down to its upper word, 20 significand bits —
// the chain's own Newton step absorbs any seed of that accuracy.  fp32 (MUFU.RSQ, ≈ 1 ulp): the correctly rounded
// value — the √ part of the fp32 chain is the vendor's fast path, which relies on the accuracy of its own seed (with 1-2
// bits cut off it misrounds a few √ in 10⁸), so its seed cannot be modelled pessimistically; the fp32 chain with the
// real seed is checked on the GPU over every operand `a` (tests/test_gpu_pdps.py).
inline double emu_rsqrt_seed(double a)
{
    double y = 1.0 / std::sqrt(a);
    unsigned long long u; std::memcpy(&u, &y, 8); u &= 0xffffffff00000000ull; std::memcpy(&y, &u, 8);
    return y;
}
inline float emu_rsqrt_seed(float a)
{
    float y = (float)(1.0 / std::sqrt((double)a));
    return y;
}
inline double rsqrt(double a) { return 1.0 / std::sqrt(a); }
inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
template <typename T> inline T __ldg(const T *p) { return *p; }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMin(int *p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline int atomicMax(int *p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long l) { double r; std::memcpy(&r, &l, 8); return r; }

// cooperative_groups::this_cluster() for kernels that use the library interface directly
namespace cooperative_groups {
struct cluster_group {
    unsigned num_blocks() const { return (unsigned)emu::cluster_size(); }
    unsigned block_rank() const { return (unsigned)emu::cluster_rank(); }
    void sync() const { emu::cluster_sync(); }
    // the address of `p` (inside this CTA's dynamic shared memory) in the CTA of rank `r`
    template <typename T> T *map_shared_rank(T *p, unsigned r) const
    {
        emu::Cta *me = emu::cta, *other = me->cluster ? me->cluster->ctas[r] : me;
        const std::ptrdiff_t off = reinterpret_cast<unsigned char *>(p) - reinterpret_cast<unsigned char *>(emu::dyn_smem_of(me));
        return reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(emu::dyn_smem_of(other)) + off);
    }
};
inline cluster_group this_cluster() { return cluster_group(); }
}  // namespace cooperative_groups

namespace emu {
// kernel<<<grid, threads>>>(args...): blocks run one after the other, the threads of a block concurrently
// (threads must be a multiple of 32: every warp barrier expects 32 arrivals)
// `csize` consecutive blocks (in x) form a cluster and run concurrently; clusters run one after the other.
template <typename F>
void launch(dim3 grid, int threads, F &&body, size_t smem_doubles = 0, int csize = 1)
{
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx0 = 0; bx0 < grid.x; bx0 += (unsigned)csize) {
            Cluster cl(csize, threads);
            std::vector<std::unique_ptr<Cta>> ctas;
            for (int r = 0; r < csize; ++r) {
                ctas.emplace_back(new Cta(threads, smem_doubles));
                ctas.back()->rank = r;
                ctas.back()->cluster = csize > 1 ? &cl : nullptr;
                cl.ctas.push_back(ctas.back().get());
            }
            std::vector<std::thread> ts;
            for (int r = 0; r < csize; ++r)
                for (int t = 0; t < threads; ++t)
                    ts.emplace_back([&, r, t] {
                        threadIdx = dim3((unsigned)t); blockIdx = dim3(bx0 + (unsigned)r, by, bz);
                        blockDim = dim3((unsigned)threads); gridDim = grid;
                        cta = ctas[r].get(); warp = cta->warps[t / 32].get();
                        body();
                    });
            for (auto &th : ts) th.join();
        }
}
}  // namespace emu
