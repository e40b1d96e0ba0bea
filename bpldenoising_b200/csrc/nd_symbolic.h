// nd_symbolic.h — host side of the nested-dissection solver (nd_solver.cuh): the elimination tree of an
// n×n pixel grid and the index tables of its frontal matrices.  Pure C++ (no CUDA), data independent:
// built once per (n, W) and shared by every image of every evaluation.
//
// Why: the adjoint systems of gradient / gradient_reg (/root/reference/src/TVLearningFunctionVec.jl:98-161,
// :192-254) couple an unknown only to unknowns of pixels at most W apart (W = 1 for forward differences).
// The reference hands them to a general sparse LU; round 1 factorised them as BANDS — O(n⁴) operations and
// an O(n²)-long dependent chain per image, on one SM.  A nested-dissection ordering of the same matrix
// costs O(n³) operations and O(n² log n) memory, and its elimination tree is log₂-deep with independent
// subtrees: thousands of small dense fronts at the bottom, a handful of n-sized ones at the top.
//
// Tree: a rectangle of pixels is cut by a separator LINE of width W across its longer side until both sides
// are ≤ LEAF; a front = its own pixels (separator line, or all pixels of a leaf: the "pivot" pixels) followed
// by the "ring" — every in-image pixel within distance W of the front's whole region, all of which belong
// to ancestors.  Fronts are sorted by processing step = deepest level first; the children of a front at
// depth d are exactly at depth d+1, i.e. in the previous step.
#pragma once
#include <algorithm>
#include <vector>

namespace bpltv {

struct NdFront {
    int npiv, nring;    // pivot / ring PIXELS (unknowns per pixel are data: 1 in node space, 1-2 per TV mode pixel)
    int pix0;           // pixlist[pix0 .. pix0+npiv+nring): pixel ids q = i + n·j, pivots first
    int nbr0;           // nbr[nbr0 + kp·NNB + e]: local index (in this front) of neighbour e of pivot pixel kp, or -1
    int cmap0;          // cmap[cmap0 + k]: local index in the PARENT's front of this front's ring pixel k
    int parent;         // front id (sorted order) or -1
    int child0, child1; // front ids or -1
    int depth;
};

// neighbour offsets of a pixel for coupling radius W.  Index h = 0 is the pixel itself; h = 1..NH-1 are the
// "forward" offsets (dj > 0, or dj == 0 and di > 0); neighbour e = 2(h-1) is +offset h, e = 2(h-1)+1 its negative.
static inline int nd_nh(int W) { return 1 + ((2 * W + 1) * (2 * W + 1) - 1) / 2; }
static inline int nd_nnb(int W) { return (2 * W + 1) * (2 * W + 1) - 1; }
static inline void nd_fwd_offset(int W, int h, int &di, int &dj)
{
    const int idx = h - 1;
    if (idx < W) { di = idx + 1; dj = 0; return; }
    const int r = idx - W;
    dj = 1 + r / (2 * W + 1);
    di = r % (2 * W + 1) - W;
}

struct NdSymbolic {
    int n = 0, W = 0, leaf = 0, nnb = 0;
    std::vector<NdFront> fronts;       // sorted by step
    std::vector<int> step_start;       // nsteps + 1
    std::vector<int> pixlist, nbr, cmap;
    std::vector<int> step_max_front_pix, step_max_ring_pix, step_max_piv_pix;
    int max_front_pix = 0, max_ring_pix = 0;
    int nsteps() const { return (int)step_start.size() - 1; }

    void build(int n_, int W_, int leaf_)
    {
        n = n_; W = W_; leaf = leaf_; nnb = nd_nnb(W);
        struct Raw { int i0, i1, j0, j1, s0, s1, dir, depth, parent, c0, c1; };   // dir: 0 leaf, 1 row cut, 2 column cut
        std::vector<Raw> raw;
        // iterative construction (explicit stack) so that huge grids cannot overflow the call stack
        struct Job { int i0, i1, j0, j1, depth, parent, slot; };
        std::vector<Job> todo;
        todo.push_back({0, n, 0, n, 0, -1, 0});
        while (!todo.empty()) {
            const Job jb = todo.back(); todo.pop_back();
            const int h = jb.i1 - jb.i0, w = jb.j1 - jb.j0;
            Raw r{jb.i0, jb.i1, jb.j0, jb.j1, 0, 0, 0, jb.depth, jb.parent, -1, -1};
            const int id = (int)raw.size();
            if (jb.parent >= 0) (jb.slot == 0 ? raw[jb.parent].c0 : raw[jb.parent].c1) = id;
            // a cut needs a pixel on both sides of the separator
            if (std::max(h, w) > leaf && std::max(h, w) >= W + 2) {
                if (h >= w) {
                    r.dir = 1; r.s0 = jb.i0 + (h - W) / 2; r.s1 = r.s0 + W;
                    raw.push_back(r);
                    todo.push_back({jb.i0, r.s0, jb.j0, jb.j1, jb.depth + 1, id, 0});
                    todo.push_back({r.s1, jb.i1, jb.j0, jb.j1, jb.depth + 1, id, 1});
                } else {
                    r.dir = 2; r.s0 = jb.j0 + (w - W) / 2; r.s1 = r.s0 + W;
                    raw.push_back(r);
                    todo.push_back({jb.i0, jb.i1, jb.j0, r.s0, jb.depth + 1, id, 0});
                    todo.push_back({jb.i0, jb.i1, r.s1, jb.j1, jb.depth + 1, id, 1});
                }
            } else {
                raw.push_back(r);
            }
        }
        int maxdepth = 0;
        for (const Raw &r : raw) maxdepth = std::max(maxdepth, r.depth);
        // sorted order: deepest first, creation order within a depth
        std::vector<int> order(raw.size()), rank(raw.size());
        for (size_t k = 0; k < raw.size(); ++k) order[k] = (int)k;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return raw[a].depth > raw[b].depth; });
        for (size_t k = 0; k < order.size(); ++k) rank[order[k]] = (int)k;
        const int nsteps_ = maxdepth + 1;
        step_start.assign(nsteps_ + 1, 0);
        for (const Raw &r : raw) ++step_start[(maxdepth - r.depth) + 1];
        for (int s = 0; s < nsteps_; ++s) step_start[s + 1] += step_start[s];

        fronts.assign(raw.size(), NdFront());
        pixlist.clear(); nbr.clear(); cmap.clear();
        std::vector<int> local((size_t)n * n, -1);
        auto fill_pixels = [&](const Raw &r, std::vector<int> &out, int &npiv) {
            out.clear();
            if (r.dir == 0) {
                for (int j = r.j0; j < r.j1; ++j) for (int i = r.i0; i < r.i1; ++i) out.push_back(i + n * j);
            } else if (r.dir == 1) {
                for (int j = r.j0; j < r.j1; ++j) for (int i = r.s0; i < r.s1; ++i) out.push_back(i + n * j);
            } else {
                for (int j = r.s0; j < r.s1; ++j) for (int i = r.i0; i < r.i1; ++i) out.push_back(i + n * j);
            }
            npiv = (int)out.size();
            const int a0 = std::max(0, r.i0 - W), a1 = std::min(n, r.i1 + W), b0 = std::max(0, r.j0 - W), b1 = std::min(n, r.j1 + W);
            for (int j = b0; j < b1; ++j)
                for (int i = a0; i < a1; ++i)
                    if (!(i >= r.i0 && i < r.i1 && j >= r.j0 && j < r.j1)) out.push_back(i + n * j);
        };
        std::vector<int> px, ppx;
        // pass 1: pixel lists and neighbour tables
        for (size_t k = 0; k < order.size(); ++k) {
            const Raw &r = raw[order[k]];
            NdFront &f = fronts[k];
            int npiv = 0;
            fill_pixels(r, px, npiv);
            f.npiv = npiv; f.nring = (int)px.size() - npiv;
            f.pix0 = (int)pixlist.size();
            f.depth = r.depth;
            f.parent = r.parent >= 0 ? rank[r.parent] : -1;
            f.child0 = r.c0 >= 0 ? rank[r.c0] : -1;
            f.child1 = r.c1 >= 0 ? rank[r.c1] : -1;
            pixlist.insert(pixlist.end(), px.begin(), px.end());
            for (size_t l = 0; l < px.size(); ++l) local[px[l]] = (int)l;
            f.nbr0 = (int)nbr.size();
            for (int kp = 0; kp < npiv; ++kp) {
                const int q = px[kp], i = q % n, j = q / n;
                for (int e = 0; e < nnb; ++e) {
                    int di, dj;
                    nd_fwd_offset(W, 1 + e / 2, di, dj);
                    if (e & 1) { di = -di; dj = -dj; }
                    const int ii = i + di, jj = j + dj;
                    nbr.push_back((ii >= 0 && ii < n && jj >= 0 && jj < n) ? local[ii + n * jj] : -1);
                }
            }
            for (size_t l = 0; l < px.size(); ++l) local[px[l]] = -1;
        }
        // pass 2: ring pixel -> local index in the parent's front
        for (size_t k = 0; k < order.size(); ++k) {
            NdFront &f = fronts[k];
            f.cmap0 = (int)cmap.size();
            if (f.parent < 0) continue;
            const NdFront &p = fronts[f.parent];
            for (int l = 0; l < p.npiv + p.nring; ++l) local[pixlist[p.pix0 + l]] = l;
            for (int l = 0; l < f.nring; ++l) cmap.push_back(local[pixlist[f.pix0 + f.npiv + l]]);   // never -1 (see header)
            for (int l = 0; l < p.npiv + p.nring; ++l) local[pixlist[p.pix0 + l]] = -1;
        }
        if (cmap.empty()) cmap.push_back(0);
        if (nbr.empty()) nbr.push_back(-1);
        step_max_front_pix.assign(nsteps_, 0); step_max_ring_pix.assign(nsteps_, 0); step_max_piv_pix.assign(nsteps_, 0);
        max_front_pix = max_ring_pix = 0;
        for (int s = 0; s < nsteps_; ++s)
            for (int t = step_start[s]; t < step_start[s + 1]; ++t) {
                const NdFront &f = fronts[t];
                step_max_front_pix[s] = std::max(step_max_front_pix[s], f.npiv + f.nring);
                step_max_ring_pix[s] = std::max(step_max_ring_pix[s], f.nring);
                step_max_piv_pix[s] = std::max(step_max_piv_pix[s], f.npiv);
                max_front_pix = std::max(max_front_pix, f.npiv + f.nring);
                max_ring_pix = std::max(max_ring_pix, f.nring);
            }
    }
};

}  // namespace bpltv
