// emu_nd.cpp — TEST INFRASTRUCTURE: runs the nested-dissection adjoint solver (nd_symbolic.h, nd_solver.cuh,
// nd_tv.cuh of bpldenoising_b200/csrc) for ONE image on the CPU thread emulation (emu_cuda.h), in the launch order
// of gradient_nd.cuh.  Built by tests/test_emu_nd.py with g++ -std=c++20 -DBPLTV_EMU; never shipped.
#include "emu_cuda.h"

#include <cstring>

#include "../../bpldenoising_b200/csrc/nd_tv.cuh"

using namespace bpltv;

// regularised != 0: NODE form (gradient_reg), else MULT form (gradient).  alpha_map: n·n doubles or NULL.
// stats_out (8 doubles or NULL): relres, guarded pivots, vanished-pivot flag, L doubles, max U doubles, fronts, levels, unknowns
extern "C" int emu_nd_gradient(int regularised, int n, const double *u, const double *ubar, const double *alpha_map,
                               double alpha_s, double gamma, double act_tol, double eps_act, int lm, int ln, int refine,
                               int leaf, double *out, double *stats_out, double *p_out, double *ast_out, int *off_out, int use_small)
{
    const int N = n * n, ng = lm * ln;
    const bool node = regularised != 0;
    const int mb = node ? 1 : 2;
    NdSymbolic sym;
    sym.build(n, 1, leaf);
    const int nf = (int)sym.fronts.size(), nsteps = sym.nsteps();
    const size_t posg_len = sym.pixlist.size() + nf;

    std::vector<double> pix((size_t)NDTV_PLANES * N, 0.0), vec((size_t)6 * N, 0.0), ast((size_t)5 * mb * mb * N, 0.0);
    std::vector<int> off(N + 2, 0), posg(posg_len, 0);
    std::vector<long long> foff((size_t)4 * nf, 0), totals(4, 0);
    int info[4] = {0, 0, 0, 0};
    NdTvSlots ws;
    ws.n = n; ws.N = N; ws.pix = pix.data(); ws.pix_stride = pix.size(); ws.off = off.data(); ws.off_stride = off.size();
    ws.vec = vec.data(); ws.vec_stride = vec.size(); ws.info = info;
    NdTvVariant gv;
    gv.patch = alpha_map != nullptr; gv.lm = lm; gv.ln = ln; gv.alpha_s = alpha_s; gv.gamma = gamma; gv.act_tol = act_tol;
    gv.eps_act = eps_act; gv.relres_tol = 1e300;

    NdDev nd;
    nd.n = n; nd.N = N; nd.W = 1; nd.nnb = sym.nnb; nd.nh = nd_nh(1); nd.mb = mb; nd.nfronts = nf; nd.nsteps = nsteps;
    nd.fronts = sym.fronts.data(); nd.pixlist = sym.pixlist.data(); nd.nbr = sym.nbr.data(); nd.cmap = sym.cmap.data();
    nd.step_start = sym.step_start.data();
    nd.off = node ? nullptr : off.data(); nd.off_stride = 0;
    nd.posg = posg.data(); nd.posg_stride = 0; nd.foff = foff.data(); nd.foff_stride = 0; nd.totals = totals.data();
    nd.ast = ast.data(); nd.ast_stride = 0; nd.info = info;

    const int chunks = std::max(1, std::min(4, (N + 255) / 256));
    if (node) {
        emu::launch(dim3(1, chunks), 256, [&] { ndtv_classify_node_kernel<double>(ws, gv, u, ubar, alpha_map, 0); });
        emu::launch(dim3(1, chunks), 256, [&] { ndtv_stencil_node_kernel(ws, ast.data(), 0); });
    } else {
        emu::launch(dim3(1), 512, [&] { ndtv_classify_mult_kernel<double>(ws, gv, u, ubar, alpha_map, 0); });
        emu::launch(dim3(1, chunks), 256, [&] { ndtv_stencil_mult_kernel(ws, ast.data(), 0); });
    }
    if (ast_out) std::memcpy(ast_out, ast.data(), ast.size() * sizeof(double));
    if (off_out) std::memcpy(off_out, off.data(), (N + 1) * sizeof(int));
    // the device computes the sizes in both forms here (the product uses host tables in the NODE form; the test
    // compares them with nd_front_sizes through the totals)
    emu::launch(dim3((nf + 7) / 8, 1), 256, [&] { nd_dims_kernel(nd); });
    emu::launch(dim3(1), 256, [&] { nd_scan_kernel(nd); });
    std::vector<double> Lp((size_t)totals[0] + 2, 0.0), U0((size_t)totals[1] + 2, 0.0), U1((size_t)totals[1] + 2, 0.0),
        UV0((size_t)totals[2] + 2, 0.0), UV1((size_t)totals[2] + 2, 0.0);
    auto al16 = [](std::vector<double> &v) { double *p = v.data(); return (reinterpret_cast<std::uintptr_t>(p) & 15) ? p + 1 : p; };
    nd.L = al16(Lp); nd.U[0] = al16(U0); nd.U[1] = al16(U1); nd.UV[0] = al16(UV0); nd.UV[1] = al16(UV1);
    nd.L_stride = nd.U_stride = nd.UV_stride = 0;

    const double guard = node ? 0.0 : 1e-13;
    std::vector<NdLevelPlan> plan(nsteps);
    for (int s = 0; s < nsteps; ++s) {
        plan[s] = nd_level_plan(sym, s, mb, node ? 1.0 : 1.25);
        if (!use_small && plan[s].small) {
            plan[s].small = false;
            plan[s].smem_f = nd_factor_smem(plan[s].nFw, s > 0 ? mb * sym.step_max_ring_pix[s - 1] : 0);
            plan[s].smem_s = nd_solve_smem(plan[s].nFw);
        }
        if (use_small == 2 && plan[s].small) {      // test hook: a starved arena forces several rounds per CTA
            plan[s].arena_f = nd_small_arena(plan[s].nFw, mb * sym.step_max_piv_pix[s], 1, true);
            plan[s].arena_s = nd_small_arena(plan[s].nFw, mb * sym.step_max_piv_pix[s], 1, false);
            plan[s].smem_f = nd_small_smem(plan[s].arena_f); plan[s].smem_s = nd_small_smem(plan[s].arena_s);
        }
    }
    for (int s = 0; s < nsteps; ++s) {
        const NdLevelPlan &lp = plan[s];
        if (lp.small)
            emu::launch(dim3((lp.nfr + ND_SMALL_WARPS - 1) / ND_SMALL_WARPS, 1), 32 * ND_SMALL_WARPS,
                        [&] { nd_factor_small_kernel(nd, lp.t0, lp.nfr, s & 1, guard, lp.arena_f); }, lp.smem_f / 8 + 2);
        else
            emu::launch(dim3(lp.nfr, 1), lp.threads_f, [&] { nd_factor_kernel(nd, lp.t0, s & 1, guard, lp.nFw); }, lp.smem_f / 8 + 2);
    }
    auto solve = [&](double *v) {
        for (int s = 0; s < nsteps; ++s) {
            const NdLevelPlan &lp = plan[s];
            if (lp.small)
                emu::launch(dim3((lp.nfr + ND_SMALL_WARPS - 1) / ND_SMALL_WARPS, 1), 32 * ND_SMALL_WARPS,
                            [&] { nd_fwd_small_kernel(nd, lp.t0, lp.nfr, s & 1, v, 0, lp.arena_s); }, lp.smem_s / 8 + 2);
            else
                emu::launch(dim3(lp.nfr, 1), lp.threads_s, [&] { nd_fwd_kernel(nd, lp.t0, s & 1, v, 0); }, lp.smem_s / 8 + 2);
        }
        for (int s = nsteps - 1; s >= 0; --s) {
            const NdLevelPlan &lp = plan[s];
            if (lp.small)
                emu::launch(dim3((lp.nfr + ND_SMALL_WARPS - 1) / ND_SMALL_WARPS, 1), 32 * ND_SMALL_WARPS,
                            [&] { nd_bwd_small_kernel(nd, lp.t0, lp.nfr, v, 0, lp.arena_s); }, lp.smem_s / 8 + 2);
            else
                emu::launch(dim3(lp.nfr, 1), lp.threads_s, [&] { nd_bwd_kernel(nd, lp.t0, v, 0); }, lp.smem_s / 8 + 2);
        }
    };
    double relres = -1.0;
    std::vector<double> out_img(ng, 0.0);
    if (node) {
        double *p = pix.data() + 7 * (size_t)N, *work = pix.data() + 8 * (size_t)N;
        solve(p);
        emu::launch(dim3(1), 512, [&] { ndtv_residual_node_kernel(ws, &relres, 0); });
        for (int it = 0; it < refine; ++it) {
            solve(work);
            emu::launch(dim3(1, chunks), 256, [&] { ndtv_axpy_node_kernel(ws, 7, 8); });
            emu::launch(dim3(1), 512, [&] { ndtv_residual_node_kernel(ws, &relres, 0); });
        }
        emu::launch(dim3(1, std::min(ng, 3)), 512, [&] { ndtv_finish_node_kernel(ws, gv, &relres, out_img.data(), 0); });
    } else {
        double *zeta = vec.data() + 2 * (size_t)N, *work = vec.data() + 4 * (size_t)N;
        emu::launch(dim3(1, chunks), 256, [&] { ndtv_copy_mult_kernel(ws, 1, 0); });
        solve(zeta);
        emu::launch(dim3(1), 512, [&] { ndtv_residual_mult_kernel(ws, &relres, 0); });
        for (int it = 0; it < refine; ++it) {
            solve(work);
            emu::launch(dim3(1, chunks), 256, [&] { ndtv_axpy_mult_kernel(ws, 1, 2); });
            emu::launch(dim3(1), 512, [&] { ndtv_residual_mult_kernel(ws, &relres, 0); });
        }
        emu::launch(dim3(1, std::min(ng, 3)), 512, [&] { ndtv_finish_mult_kernel(ws, gv, &relres, out_img.data(), 0); });
    }
    std::memcpy(out, out_img.data(), ng * sizeof(double));
    if (p_out) std::memcpy(p_out, pix.data() + 7 * (size_t)N, N * sizeof(double));
    if (stats_out) {
        stats_out[0] = relres; stats_out[1] = info[0]; stats_out[2] = info[1]; stats_out[3] = (double)totals[0];
        stats_out[4] = (double)totals[1]; stats_out[5] = nf; stats_out[6] = nsteps; stats_out[7] = node ? N : off[N];
    }
    return 0;
}

// the symbolic structure alone: checks of the tree (every pixel a pivot exactly once, rings inside ancestors)
extern "C" int emu_nd_symbolic_check(int n, int W, int leaf, int *nfronts, int *nsteps, int *max_front_pix)
{
    NdSymbolic sym;
    sym.build(n, W, leaf);
    const int nf = (int)sym.fronts.size();
    *nfronts = nf; *nsteps = sym.nsteps(); *max_front_pix = sym.max_front_pix;
    std::vector<int> owner((size_t)n * n, -1);
    for (int t = 0; t < nf; ++t) {
        const NdFront &f = sym.fronts[t];
        for (int k = 0; k < f.npiv; ++k) {
            const int q = sym.pixlist[f.pix0 + k];
            if (owner[q] != -1) return 1;               // a pixel eliminated twice
            owner[q] = t;
        }
    }
    for (int q = 0; q < n * n; ++q) if (owner[q] < 0) return 2;
    for (int t = 0; t < nf; ++t) {
        const NdFront &f = sym.fronts[t];
        if (f.parent >= 0 && !(f.parent > t)) return 3;                       // parents come later
        if (f.parent >= 0 && sym.fronts[f.parent].depth != f.depth - 1) return 4;
        for (int k = 0; k < f.nring; ++k) {                                   // ring pixels belong to proper ancestors
            const int q = sym.pixlist[f.pix0 + f.npiv + k];
            int a = f.parent;
            while (a >= 0 && a != owner[q]) a = sym.fronts[a].parent;
            if (a < 0) return 5;
            if (f.parent >= 0) {
                const NdFront &p = sym.fronts[f.parent];
                const int l = sym.cmap[f.cmap0 + k];
                if (l < 0 || l >= p.npiv + p.nring || sym.pixlist[p.pix0 + l] != q) return 6;
            }
        }
        // every in-image neighbour (distance ≤ W) of a pivot pixel is in the front or in a descendant
        for (int k = 0; k < f.npiv; ++k) {
            const int q = sym.pixlist[f.pix0 + k], i = q % n, j = q / n;
            for (int e = 0; e < sym.nnb; ++e) {
                int di, dj;
                nd_fwd_offset(W, 1 + e / 2, di, dj);
                if (e & 1) { di = -di; dj = -dj; }
                const int ii = i + di, jj = j + dj;
                const int l = sym.nbr[f.nbr0 + k * sym.nnb + e];
                if (ii < 0 || ii >= n || jj < 0 || jj >= n) { if (l != -1) return 7; continue; }
                const int qq = ii + n * jj;
                if (l >= 0) { if (sym.pixlist[f.pix0 + l] != qq) return 8; continue; }
                int a = owner[qq];                      // must be a descendant of t
                while (a >= 0 && a != t) a = sym.fronts[a].parent;
                if (a < 0) return 9;
            }
        }
    }
    return 0;
}

// pixel elimination order of the tree (pivot pixels front by front, in processing order)
extern "C" int emu_nd_order(int n, int W, int leaf, int *pix_order)
{
    NdSymbolic sym;
    sym.build(n, W, leaf);
    int k = 0;
    for (const NdFront &f : sym.fronts)
        for (int l = 0; l < f.npiv; ++l) pix_order[k++] = sym.pixlist[f.pix0 + l];
    return k;
}
