// halo_async.cuh — halo exchange between the CTAs of a thread-block cluster without a cluster barrier (device code;
// used by the cluster-resident solves pdps_resident.cuh and pdps_sumregs.cuh; emulation-compatible, -DBPLTV_EMU).
#pragma once
#ifdef BPLTV_EMU
#include <thread>
#else
#include <cuda_runtime.h>
#endif

namespace bpltv {

// ---------------------------------------------------------------------------
// Halo exchange WITHOUT a cluster barrier.  cooperative_groups' cluster.sync() is
// barrier.cluster.arrive.release + wait.acquire, which ptxas renders as MEMBAR.ALL.GPU; ERRBAR; CGAERRBAR; UCGABAR_ARV;
// UCGABAR_WAIT; CCTL.IVALL — a GPU-scope memory fence and an L1 invalidation by every thread, twice per iteration
// (cuobjdump of round 1's kernel B).  A producer–consumer pair does not need it: the owner of a boundary column sends
// it with `st.async` — a remote shared-memory store that counts its bytes on an mbarrier in the RECEIVER's shared memory
// (complete_tx) — and the receiver waits on that mbarrier for the whole column.  The data is visible when the phase
// completes; no fence, no cluster-wide rendez-vous, and a CTA only waits for the neighbour it actually needs.
// Write-after-read safety of the single halo slots: the threads that read a halo column are exactly the threads that
// send the column the neighbour needs before it can overwrite that slot (x̄-halo readers send y2, y2-halo readers send
// x̄), each send depends on the value read, and a phase completes only when ALL of them have sent.
// Emulation (tests/emu): plain remote stores + an atomic byte counter, the wait compares it with (phase+1)·bytes.
// ---------------------------------------------------------------------------
#ifdef BPLTV_EMU
static inline void halo_bar_init(unsigned long long *bar) { __atomic_store_n(bar, 0ULL, __ATOMIC_RELEASE); }
static inline void halo_bar_arm(unsigned long long *, unsigned) {}
template <typename Real>
static inline void halo_push2(Real *remote, Real v0, Real v1, unsigned long long *remote_bar)
{
    remote[0] = v0; remote[1] = v1;
    __atomic_fetch_add(remote_bar, (unsigned long long)(2 * sizeof(Real)), __ATOMIC_RELEASE);
}
template <typename Real>
static inline void halo_push1(Real *remote, Real v, unsigned long long *remote_bar)
{
    remote[0] = v;
    __atomic_fetch_add(remote_bar, (unsigned long long)sizeof(Real), __ATOMIC_RELEASE);
}
static inline void halo_wait(unsigned long long *bar, int phase, unsigned colbytes)
{
    while (__atomic_load_n(bar, __ATOMIC_ACQUIRE) < (unsigned long long)(phase + 1) * colbytes) std::this_thread::yield();
}
#else
static __device__ __forceinline__ unsigned res_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void halo_bar_init(unsigned long long *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(res_smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one arrival + the bytes the phase expects (the neighbour's sends may have come first: the count only has to balance)
static __device__ __forceinline__ void halo_bar_arm(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(res_smem_u32(bar)), "r"(bytes) : "memory");
}
// `remote` / `remote_bar`: generic addresses in the neighbour's shared memory (cluster.map_shared_rank)
static __device__ __forceinline__ void halo_push2(double *remote, double v0, double v1, unsigned long long *remote_bar)
{
    unsigned ra, rb;
    asm volatile("cvt.u32.u64 %0, %1;" : "=r"(ra) : "l"(__cvta_generic_to_shared(remote)));
    asm volatile("cvt.u32.u64 %0, %1;" : "=r"(rb) : "l"(__cvta_generic_to_shared(remote_bar)));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(ra), "d"(v0),
                 "d"(v1), "r"(rb)
                 : "memory");
}
static __device__ __forceinline__ void halo_push2(float *remote, float v0, float v1, unsigned long long *remote_bar)
{
    unsigned ra, rb;
    asm volatile("cvt.u32.u64 %0, %1;" : "=r"(ra) : "l"(__cvta_generic_to_shared(remote)));
    asm volatile("cvt.u32.u64 %0, %1;" : "=r"(rb) : "l"(__cvta_generic_to_shared(remote_bar)));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(ra), "f"(v0),
                 "f"(v1), "r"(rb)
                 : "memory");
}
static __device__ __forceinline__ void halo_push1(double *remote, double v, unsigned long long *remote_bar)
{
    unsigned ra, rb;
    asm volatile("cvt.u32.u64 %0, %1;" : "=r"(ra) : "l"(__cvta_generic_to_shared(remote)));
    asm volatile("cvt.u32.u64 %0, %1;" : "=r"(rb) : "l"(__cvta_generic_to_shared(remote_bar)));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(ra), "d"(v), "r"(rb) : "memory");
}
static __device__ __forceinline__ void halo_push1(float *remote, float v, unsigned long long *remote_bar)
{
    unsigned ra, rb;
    asm volatile("cvt.u32.u64 %0, %1;" : "=r"(ra) : "l"(__cvta_generic_to_shared(remote)));
    asm volatile("cvt.u32.u64 %0, %1;" : "=r"(rb) : "l"(__cvta_generic_to_shared(remote_bar)));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(ra), "f"(v), "r"(rb) : "memory");
}
static __device__ __forceinline__ void halo_wait(unsigned long long *bar, int phase, unsigned)
{
    const unsigned addr = res_smem_u32(bar), parity = (unsigned)phase & 1u;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RES_WAIT:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RES_DONE;\n"
        "bra RES_WAIT;\n"
        "RES_DONE:\n"
        "}\n" ::"r"(addr), "r"(parity) : "memory");
}
#endif


}  // namespace bpltv
