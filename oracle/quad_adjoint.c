/* quad_adjoint.c — TEST INFRASTRUCTURE (never linked into the product): the adjoint systems of
 * /root/reference/src/TVLearningFunctionVec.jl solved in IEEE binary128 (__float128, 113-bit mantissa), to
 * decide what the reference's gradient IS at the 1e-10 level the parity bar asks for.
 *
 * gradient (:98-135 scalar, :219-254 patch) assembles
 *     Adj = [ I  −Gᵀ ; Act·G + Inact·α(Den − prodKuKu)·G   Inact + eps·Act ],   Adj \ [u − ū; 0]
 * with entries from 1 to 4.5e15 and hands it to a sparse LU; in double precision the answer moves by 1e-6…3e-5
 * with the ordering.  Here the same system is solved by a band LU with partial pivoting in
 * binary128 (unknowns interleaved per pixel: p, μ₁, μ₂ → half-bandwidth 3n+2), in two flavours:
 *   assemble_quad = 0   the ENTRIES are formed in double with the reference's operation order (what Julia's sparse
 *                       algebra stores), only the solve is exact  → "the reference's matrix, solved exactly";
 *   assemble_quad = 1   entries formed in binary128 from the same double u  → "the system the formulas mean".
 * The difference between the two is the rounding of `Den − prodKuKu` (a rank-one projector formed by cancellation,
 * then scaled by α/|∇u| up to 1e11), i.e. noise of the reference's own assembly.
 * quad_gradient_compliance solves the multiplier-space form the CUDA path factorises, (diag(E) + B Bᵀ) ζ = B r,
 * p = r − Bᵀζ, in binary128; quad_gradient_reg the node-space system of gradient_reg (:137-161, :192-215).
 * All return p (double-rounded) and the per-pixel functional; the caller sums (scalar) or patch-sums it.
 */
#include <math.h>
#include <quadmath.h>
#include <stdlib.h>
#include <string.h>

typedef __float128 q;

/* band LU with partial pivoting, rows stored with their window: row i holds columns i−kl … i+ku+kl at
 * ab[i·ld + (j − i + kl)], ld = 2kl+ku+1 (the last kl are fill).  Solves in place; returns 0, or k+1 for a zero pivot. */
static int band_solve(int N, int kl, int ku, q *ab, q *b)
{
    const int ld = 2 * kl + ku + 1;
    for (int k = 0; k < N; ++k) {
        int piv = k;
        q best = fabsq(ab[(size_t)k * ld + kl]);
        const int rmax = k + kl < N - 1 ? k + kl : N - 1;
        for (int r = k + 1; r <= rmax; ++r) {
            const q v = fabsq(ab[(size_t)r * ld + (k - r + kl)]);
            if (v > best) { best = v; piv = r; }
        }
        if (best == 0) return k + 1;
        const int cmax = k + ku + kl < N - 1 ? k + ku + kl : N - 1;
        if (piv != k) {
            for (int c = k; c <= cmax; ++c) {
                q *a = &ab[(size_t)k * ld + (c - k + kl)], *bq = &ab[(size_t)piv * ld + (c - piv + kl)];
                const q t = *a; *a = *bq; *bq = t;
            }
            const q t = b[k]; b[k] = b[piv]; b[piv] = t;
        }
        const q d = ab[(size_t)k * ld + kl];
        for (int r = k + 1; r <= rmax; ++r) {
            q *row = &ab[(size_t)r * ld + (k - r + kl)];       /* entry (r, k) */
            if (row[0] == 0) continue;
            const q m = row[0] / d;
            row[0] = 0;
            const q *prow = &ab[(size_t)k * ld + kl];
            for (int c = 1; c <= cmax - k; ++c) row[c] -= m * prow[c];
            b[r] -= m * b[k];
        }
    }
    for (int k = N - 1; k >= 0; --k) {
        const int cmax = k + ku + kl < N - 1 ? k + ku + kl : N - 1;
        q s = b[k];
        const q *prow = &ab[(size_t)k * ld + kl];
        for (int c = 1; c <= cmax - k; ++c) s -= prow[c] * b[k + c];
        b[k] = s / prow[0];
    }
    return 0;
}

#define AT(i, j) ab[(size_t)(i) * ld + ((j) - (i) + kl)]

/* gradient (non-regularised).  alpha_map NULL → scalar alpha.  Outputs: p (n·n), fpix (n·n): −⟨(Gp)_q, Inact·Den·Gu⟩. */
int quad_gradient_literal(int n, const double *u, const double *ubar, double alpha, const double *alpha_map,
                          double act_tol, double eps_act, int assemble_quad, double *p_out, double *fpix_out)
{
    const int N = n * n, NN = 3 * N, kl = 3 * n + 2, ku = 3 * n + 2, ld = 2 * kl + ku + 1;
    q *ab = calloc((size_t)NN * ld, sizeof(q)), *b = calloc(NN, sizeof(q));
    q *w1 = calloc(N, sizeof(q)), *w2 = calloc(N, sizeof(q));
    if (!ab || !b || !w1 || !w2) return -1;
    for (int v = 0; v < N; ++v) {
        const int i = v % n, j = v / n;
        const int h1 = i + 1 < n, h2 = j + 1 < n;
        const double a = alpha_map ? alpha_map[v] : alpha;
        /* row p_v:  p_v − (Gᵀμ)_v = u_v − ū_v;  (Gᵀμ)_v = μ₁(v−1)[i>0] − μ₁(v)[h1] + μ₂(v−n)[j>0] − μ₂(v)[h2] */
        AT(3 * v, 3 * v) = 1;
        if (i > 0) AT(3 * v, 3 * (v - 1) + 1) = -1;
        if (h1) AT(3 * v, 3 * v + 1) = 1;
        if (j > 0) AT(3 * v, 3 * (v - n) + 2) = -1;
        if (h2) AT(3 * v, 3 * v + 2) = 1;
        b[3 * v] = (q)u[v] - (q)ubar[v];
        /* rows μ_v */
        const double g1d = h1 ? u[v + 1] - u[v] : 0.0, g2d = h2 ? u[v + n] - u[v] : 0.0;      /* G*u[:] in double (:107) */
        const double nrmd = sqrt(g1d * g1d + g2d * g2d);                                      /* xi(Gu) (:108) */
        const int act = nrmd < act_tol;                                                       /* (:109) */
        q t11, t12, t21, t22;
        if (act) {
            /* Act·G + eps·Act: (Gp)_v + eps μ_v = 0 */
            t11 = 1; t12 = 0; t21 = 0; t22 = 1;
            AT(3 * v + 1, 3 * v + 1) = eps_act;
            AT(3 * v + 2, 3 * v + 2) = eps_act;
            w1[v] = 0; w2[v] = 0;
        } else {
            if (assemble_quad) {
                const q g1 = h1 ? (q)u[v + 1] - (q)u[v] : 0, g2 = h2 ? (q)u[v + n] - (q)u[v] : 0;
                const q nrm = sqrtq(g1 * g1 + g2 * g2), id = 1 / nrm, d3 = nrm * nrm * nrm;
                t11 = (q)a * (id - g1 / d3 * g1); t12 = (q)a * (-(g1 / d3 * g2));
                t21 = (q)a * (-(g2 / d3 * g1)); t22 = (q)a * (id - g2 / d3 * g2);
                w1[v] = id * g1; w2[v] = id * g2;
            } else {
                /* Den = 1/den; prodKuKu = prodesc(Gu ./ den.^3, Gu); α·(Den − prodKuKu): one rounding per operation */
                const double den = nrmd, id = 1.0 / den, d3 = den * den * den;
                const double a1 = g1d / d3, a2 = g2d / d3;
                const double e11 = a * (id - a1 * g1d), e12 = a * (0.0 - a1 * g2d), e21 = a * (0.0 - a2 * g1d), e22 = a * (id - a2 * g2d);
                t11 = e11; t12 = e12; t21 = e21; t22 = e22;
                const double v1 = id * g1d, v2 = id * g2d;                                    /* Den·Gu (:132) */
                w1[v] = v1; w2[v] = v2;
            }
            AT(3 * v + 1, 3 * v + 1) = 1;      /* Inact */
            AT(3 * v + 2, 3 * v + 2) = 1;
        }
        /* T·G restricted to pixel v: (Gp)₁ = p(v+1) − p(v) [h1], (Gp)₂ = p(v+n) − p(v) [h2].  The sparse product sums the
         * two contributions to the coefficient of p_v in one addition (double in the reference's flavour). */
        q c0a, c0b;
        if (assemble_quad || act) { c0a = -(h1 ? t11 : 0) - (h2 ? t12 : 0); c0b = -(h1 ? t21 : 0) - (h2 ? t22 : 0); }
        else {
            const double s1 = -(h1 ? (double)t11 : 0.0) - (h2 ? (double)t12 : 0.0), s2 = -(h1 ? (double)t21 : 0.0) - (h2 ? (double)t22 : 0.0);
            c0a = s1; c0b = s2;
        }
        AT(3 * v + 1, 3 * v) = c0a;
        AT(3 * v + 2, 3 * v) = c0b;
        if (h1) { AT(3 * v + 1, 3 * (v + 1)) = t11; AT(3 * v + 2, 3 * (v + 1)) = t21; }
        if (h2) { AT(3 * v + 1, 3 * (v + n)) = t12; AT(3 * v + 2, 3 * (v + n)) = t22; }
    }
    const int rc = band_solve(NN, kl, ku, ab, b);
    if (rc == 0)
        for (int v = 0; v < N; ++v) {
            const int i = v % n, j = v / n;
            const q pv = b[3 * v];
            const q d1 = i + 1 < n ? b[3 * (v + 1)] - pv : 0, d2 = j + 1 < n ? b[3 * (v + n)] - pv : 0;
            p_out[v] = (double)pv;
            fpix_out[v] = (double)(-(d1 * w1[v] + d2 * w2[v]));
        }
    free(ab); free(b); free(w1); free(w2);
    return rc;
}

/* the multiplier-space (compliance) form in binary128: modes numbered in pixel order, half-bandwidth ≤ 2n+2 */
int quad_gradient_compliance(int n, const double *u, const double *ubar, double alpha, const double *alpha_map,
                             double act_tol, double eps_act, double *p_out, double *fpix_out)
{
    const int N = n * n;
    int *off = malloc((N + 1) * sizeof(int));
    q *e1 = calloc(2 * (size_t)N, sizeof(q)), *e2 = calloc(2 * (size_t)N, sizeof(q)), *E = calloc(2 * (size_t)N, sizeof(q));
    q *w1 = calloc(N, sizeof(q)), *w2 = calloc(N, sizeof(q));
    int *pixof = malloc(2 * (size_t)N * sizeof(int));
    if (!off || !e1 || !e2 || !E || !w1 || !w2 || !pixof) return -1;
    int Nd = 0;
    for (int v = 0; v < N; ++v) {
        const int i = v % n, j = v / n;
        const int h1 = i + 1 < n, h2 = j + 1 < n;
        const double a = alpha_map ? alpha_map[v] : alpha;
        const double g1d = h1 ? u[v + 1] - u[v] : 0.0, g2d = h2 ? u[v + n] - u[v] : 0.0;
        const int act = sqrt(g1d * g1d + g2d * g2d) < act_tol;      /* the classification is the reference's (double) */
        off[v] = Nd;
        if (act) {
            e1[Nd] = 1; e2[Nd] = 0; E[Nd] = eps_act; pixof[Nd] = v; ++Nd;
            e1[Nd] = 0; e2[Nd] = 1; E[Nd] = eps_act; pixof[Nd] = v; ++Nd;
        } else {
            const q g1 = h1 ? (q)u[v + 1] - (q)u[v] : 0, g2 = h2 ? (q)u[v + n] - (q)u[v] : 0;
            const q nrm = sqrtq(g1 * g1 + g2 * g2);
            e1[Nd] = -g2 / nrm; e2[Nd] = g1 / nrm; E[Nd] = nrm / (q)a; pixof[Nd] = v; ++Nd;
            w1[v] = g1 / nrm; w2[v] = g2 / nrm;
        }
    }
    off[N] = Nd;
    const int kl = 2 * n + 3, ku = kl, ld = 2 * kl + ku + 1;
    q *ab = calloc((size_t)Nd * ld, sizeof(q)), *b = calloc(Nd, sizeof(q));
    if (!ab || !b) return -1;
    /* node coefficients of mode m: β0 at v, β1 at v+1, β2 at v+n */
#define BETA(m, be) do { const int v_ = pixof[m], i_ = v_ % n, j_ = v_ / n; \
        be[1] = i_ + 1 < n ? e1[m] : 0; be[2] = j_ + 1 < n ? e2[m] : 0; be[0] = -(be[1] + be[2]); } while (0)
    for (int m = 0; m < Nd; ++m) {
        const int v = pixof[m], i = v % n, j = v / n;
        q bm[3]; BETA(m, bm);
        const int nodes[3] = {v, v + 1, v + n};
        const int have[3] = {1, i + 1 < n, j + 1 < n};
        q rhs = 0;
        for (int s = 0; s < 3; ++s) if (have[s]) rhs += bm[s] * ((q)u[nodes[s]] - (q)ubar[nodes[s]]);
        b[m] = rhs;
        /* partner modes: pixels within one step whose stencils share a node */
        for (int dj = -1; dj <= 1; ++dj)
            for (int di = -1; di <= 1; ++di) {
                const int ii = i + di, jj = j + dj;
                if (ii < 0 || ii >= n || jj < 0 || jj >= n) continue;
                const int vq = ii + n * jj;
                for (int mq = off[vq]; mq < off[vq + 1]; ++mq) {
                    q bq[3]; BETA(mq, bq);
                    const int nq[3] = {vq, vq + 1, vq + n};
                    const int hq[3] = {1, ii + 1 < n, jj + 1 < n};
                    q acc = 0; int any = 0;
                    for (int s = 0; s < 3; ++s)
                        for (int s2 = 0; s2 < 3; ++s2)
                            if (have[s] && hq[s2] && nodes[s] == nq[s2]) { acc += bm[s] * bq[s2]; any = 1; }
                    if (mq == m) { acc += E[m]; any = 1; }
                    if (any) AT(m, mq) = acc;
                }
            }
    }
    const int rc = band_solve(Nd, kl, ku, ab, b);
    if (rc == 0) {
        q *p = calloc(N, sizeof(q));
        for (int v = 0; v < N; ++v) p[v] = (q)u[v] - (q)ubar[v];
        for (int m = 0; m < Nd; ++m) {
            const int v = pixof[m], i = v % n, j = v / n;
            q bm[3]; BETA(m, bm);
            p[v] -= bm[0] * b[m];
            if (i + 1 < n) p[v + 1] -= bm[1] * b[m];
            if (j + 1 < n) p[v + n] -= bm[2] * b[m];
        }
        for (int v = 0; v < N; ++v) {
            const int i = v % n, j = v / n;
            const q d1 = i + 1 < n ? p[v + 1] - p[v] : 0, d2 = j + 1 < n ? p[v + n] - p[v] : 0;
            p_out[v] = (double)p[v];
            fpix_out[v] = (double)(-(d1 * w1[v] + d2 * w2[v]));
        }
        free(p);
    }
    free(ab); free(b); free(off); free(e1); free(e2); free(E); free(w1); free(w2); free(pixof);
    return rc;
}

/* gradient_reg: (I + diag(a)·Gᵀ(B − C)G) p = ū − u, a = α (scalar) or the node's α (patch, `α[:] .*`, :212).
 * fpix: scalar — ⟨(Gp)_q, w_q⟩ per pixel (:159); patch — p_v (Gᵀw)_v per node (:213). */
int quad_gradient_reg(int n, const double *u, const double *ubar, double alpha, const double *alpha_map, double gamma,
                      int assemble_quad, double *p_out, double *fpix_out)
{
    const int N = n * n, kl = n + 1, ku = n + 1, ld = 2 * kl + ku + 1;
    q *ab = calloc((size_t)N * ld, sizeof(q)), *b = calloc(N, sizeof(q));
    q *T = calloc(4 * (size_t)N, sizeof(q)), *w1 = calloc(N, sizeof(q)), *w2 = calloc(N, sizeof(q));
    if (!ab || !b || !T || !w1 || !w2) return -1;
    for (int v = 0; v < N; ++v) {
        const int i = v % n, j = v / n;
        const int h1 = i + 1 < n, h2 = j + 1 < n;
        const double g1d = h1 ? u[v + 1] - u[v] : 0.0, g2d = h2 ? u[v + n] - u[v] : 0.0;
        const double nrmd = sqrt(g1d * g1d + g2d * g2d);
        const int act = fmax(0.0, nrmd - 1.0 / gamma) != 0.0;           /* :146-147 */
        if (!act) { T[4 * v] = gamma; T[4 * v + 3] = gamma; w1[v] = (q)gamma * g1d; w2[v] = (q)gamma * g2d; if (!assemble_quad) { w1[v] = gamma * g1d; w2[v] = gamma * g2d; } }
        else if (assemble_quad) {
            const q g1 = h1 ? (q)u[v + 1] - (q)u[v] : 0, g2 = h2 ? (q)u[v + n] - (q)u[v] : 0;
            const q nrm = sqrtq(g1 * g1 + g2 * g2), id = 1 / nrm, d3 = nrm * nrm * nrm;
            T[4 * v] = id - g1 / d3 * g1; T[4 * v + 1] = -(g1 / d3 * g2); T[4 * v + 2] = -(g2 / d3 * g1); T[4 * v + 3] = id - g2 / d3 * g2;
            w1[v] = id * g1; w2[v] = id * g2;
        } else {
            const double den = nrmd, id = 1.0 / den, d3 = den * den * den, a1 = g1d / d3, a2 = g2d / d3;
            /* C = Act(prodGuGu − Den); B − C = Den − prodGuGu on act */
            T[4 * v] = -(a1 * g1d - id); T[4 * v + 1] = -(a1 * g2d); T[4 * v + 2] = -(a2 * g1d); T[4 * v + 3] = -(a2 * g2d - id);
            w1[v] = id * g1d; w2[v] = id * g2d;
        }
    }
    for (int v = 0; v < N; ++v) {
        const int i = v % n, j = v / n;
        const q a = alpha_map ? (q)alpha_map[v] : (q)alpha;
        AT(v, v) += 1;
        b[v] = (q)ubar[v] - (q)u[v];
        /* (Gᵀ T G)[v, ·]: pixels containing node v */
        for (int w = 0; w < 3; ++w) {
            if ((w == 1 && i == 0) || (w == 2 && j == 0)) continue;
            const int qv = w == 0 ? v : (w == 1 ? v - 1 : v - n), qi = qv % n, qj = qv / n;
            const int h1 = qi + 1 < n, h2 = qj + 1 < n;
            q cv[2];
            if (w == 0) { cv[0] = -h1; cv[1] = -h2; } else if (w == 1) { cv[0] = h1; cv[1] = 0; } else { cv[0] = 0; cv[1] = h2; }
            const q s1 = cv[0] * T[4 * qv] + cv[1] * T[4 * qv + 2], s2 = cv[0] * T[4 * qv + 1] + cv[1] * T[4 * qv + 3];   /* g_vᵀ T */
            const int nodes[3] = {qv, qv + 1, qv + n};
            const q c1[3] = {-h1, h1, 0}, c2[3] = {-h2, 0, h2};
            for (int s = 0; s < 3; ++s) {
                if ((s == 1 && !h1) || (s == 2 && !h2)) continue;
                AT(v, nodes[s]) += a * (s1 * c1[s] + s2 * c2[s]);
            }
        }
    }
    const int rc = band_solve(N, kl, ku, ab, b);
    if (rc == 0)
        for (int v = 0; v < N; ++v) {
            const int i = v % n, j = v / n;
            p_out[v] = (double)b[v];
            if (!alpha_map) {
                const q d1 = i + 1 < n ? b[v + 1] - b[v] : 0, d2 = j + 1 < n ? b[v + n] - b[v] : 0;
                fpix_out[v] = (double)(d1 * w1[v] + d2 * w2[v]);
            } else {
                q s = 0;
                if (i > 0) s += w1[v - 1];
                if (i + 1 < n) s -= w1[v];
                if (j > 0) s += w2[v - n];
                if (j + 1 < n) s -= w2[v];
                fpix_out[v] = (double)(b[v] * s);
            }
        }
    free(ab); free(b); free(T); free(w1); free(w2);
    return rc;
}
