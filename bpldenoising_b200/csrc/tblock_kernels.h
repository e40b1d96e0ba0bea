// tblock_kernels.h — host-side handles of the temporally blocked PDPS kernels (pdps_tblock.cuh).
// The ~120 instantiations live in their own translation units (tblock_f64_t2.cu, …), one per
// precision and depth, so that `make -j` compiles them in parallel.
#pragma once
#include "pdps_tblock.cuh"

namespace bpltv {

template <typename Real, int T>
using TBlockFn = void (*)(const TBlockArgs<Real, T>);

// Kernel for vector width `vec`, λ-map / scalar, strict / fast arithmetic, plain / λ-sweep stack;
// nullptr when that combination is not built (vec not offered for the precision; sweeps at T = 3).
template <typename Real, int T>
TBlockFn<Real, T> tblock_kernel(int vec, bool map, bool strict, bool batch);

}  // namespace bpltv
