// pdps_resident.cuh — kernel B: whole image resident on chip for all iterations.
// (interface; the kernel body lands after the streaming path is validated)
#pragma once
#include "common.cuh"

namespace bpltv {

template <typename Real>
struct ResidentArgs {
    const Real *f;
    Real *u_out;
    const Real *alpha_map;
    const StepConsts<Real> *steps;
    int maxiter, M, N, O, init_mode;
    Real alpha_s;
};

template <typename Real>
static inline bool resident_eligible(size_t /*smem_optin*/, int /*M*/, int /*N*/) { return false; }

template <typename Real>
static inline int resident_cluster_size(int /*M*/, int /*N*/) { return 1; }

template <typename Real>
static inline int launch_resident(const ResidentArgs<Real> &, bool /*map*/, bool /*strict*/, cudaStream_t) { return -1; }

}  // namespace bpltv
