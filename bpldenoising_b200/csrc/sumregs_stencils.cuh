// sumregs_stencils.cuh — the three difference operators of the sum-of-regularisers model as stencils
// (forward S4, backward S10, centred S11 of docs/SEMANTICS.md; `matrix(op, n)` of the un-vendored
// FwdGradientOp / BwdGradientOp / CenteredGradientOp, call sites
// /root/reference/src/SumRegsLearningFunction.jl:119, :135, :151).  Shared by the compliance-form
// gradient (gradient_sumregs.cuh) and the node-space LU gradient (gradient_sumregs_lu.cuh); depends on
// nothing, so the CPU thread emulation of tests/emu/ can compile it.
#pragma once

namespace bpltv {

// (G_k p)(q): forward (k=0, S4), backward (k=1, S10), centred (k=2, S11) differences
template <typename T>
static __device__ __forceinline__ void op_apply(int k, int i, int j, int n, const T *p, int q, double &d1, double &d2)
{
    d1 = 0.0; d2 = 0.0;
    if (k == 0) {
        if (i + 1 < n) d1 = (double)p[q + 1] - (double)p[q];
        if (j + 1 < n) d2 = (double)p[q + n] - (double)p[q];
    } else if (k == 1) {
        if (i >= 1) d1 = (double)p[q] - (double)p[q - 1];
        if (j >= 1) d2 = (double)p[q] - (double)p[q - n];
    } else {
        if (i >= 1 && i <= n - 2) d1 = 0.5 * ((double)p[q + 1] - (double)p[q - 1]);
        if (j >= 1 && j <= n - 2) d2 = 0.5 * ((double)p[q + n] - (double)p[q - n]);
    }
}

// Every (pixel q, operator k) whose stencil touches node (i,j), with the coefficients c1, c2 that
// components 1 and 2 of (G_k ·)(q) put on that node: fn(q, k, c1, c2).
template <typename F>
static __device__ __forceinline__ void visit_node(int i, int j, int n, F &&fn)
{
    const int v = j * n + i;
    {   // forward differences
        const double c1 = (i + 1 < n) ? -1.0 : 0.0, c2 = (j + 1 < n) ? -1.0 : 0.0;
        if (c1 != 0.0 || c2 != 0.0) fn(v, 0, c1, c2);
        if (i > 0) fn(v - 1, 0, 1.0, 0.0);
        if (j > 0) fn(v - n, 0, 0.0, 1.0);
    }
    {   // backward differences
        const double c1 = (i >= 1) ? 1.0 : 0.0, c2 = (j >= 1) ? 1.0 : 0.0;
        if (c1 != 0.0 || c2 != 0.0) fn(v, 1, c1, c2);
        if (i + 1 < n) fn(v + 1, 1, -1.0, 0.0);
        if (j + 1 < n) fn(v + n, 1, 0.0, -1.0);
    }
    // centred differences (rows / columns 1..n-2 only)
    if (i - 1 >= 1) fn(v - 1, 2, 0.5, 0.0);           // pixel row i-1 ≤ n-2 always
    if (i + 1 <= n - 2) fn(v + 1, 2, -0.5, 0.0);      // pixel row i+1 ≥ 1 always
    if (j - 1 >= 1) fn(v - n, 2, 0.0, 0.5);
    if (j + 1 <= n - 2) fn(v + n, 2, 0.0, -0.5);
}

// The nodes of the stencil of (pixel (i,j), operator k), with the coefficients of components 1 and 2:
// fn(node, c1, c2), in a fixed order.
template <typename F>
static __device__ __forceinline__ void visit_stencil(int k, int i, int j, int n, F &&fn)
{
    const int q = j * n + i;
    if (k == 0) {
        const double c1 = (i + 1 < n) ? -1.0 : 0.0, c2 = (j + 1 < n) ? -1.0 : 0.0;
        if (c1 != 0.0 || c2 != 0.0) fn(q, c1, c2);
        if (i + 1 < n) fn(q + 1, 1.0, 0.0);
        if (j + 1 < n) fn(q + n, 0.0, 1.0);
    } else if (k == 1) {
        const double c1 = (i >= 1) ? 1.0 : 0.0, c2 = (j >= 1) ? 1.0 : 0.0;
        if (c1 != 0.0 || c2 != 0.0) fn(q, c1, c2);
        if (i >= 1) fn(q - 1, -1.0, 0.0);
        if (j >= 1) fn(q - n, 0.0, -1.0);
    } else {
        if (i >= 1 && i <= n - 2) { fn(q - 1, -0.5, 0.0); fn(q + 1, 0.5, 0.0); }
        if (j >= 1 && j <= n - 2) { fn(q - n, 0.0, -0.5); fn(q + n, 0.0, 0.5); }
    }
}

}  // namespace bpltv
