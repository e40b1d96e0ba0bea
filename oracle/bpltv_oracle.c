/* bpltv_oracle.c — CPU restatement of BPLDenoising's inner TV solve.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libbpltv.so) never links, imports or calls anything under oracle/.
 *
 * PARITY UNPINNED: the reference (dvillacis/BPLDenoising) delegates the solver
 * arithmetic to three un-vendored, un-pinned Julia packages (VariationalImaging,
 * AlgTools, ImageTools: /root/reference/Project.toml:7,14,25; README.md:6-20)
 * and ships no golden vectors (/root/reference/test/runtests.jl:4-6 is an empty
 * testset).  Julia is not installed here, so this restatement cannot be checked
 * against the reference's own output.  What it follows:
 *   - solver parameters and call protocol: src/TVLearningFunctionVec.jl:33-70
 *   - the published accelerated PDPS recursion that `op_denoise_pdps`
 *     generalises (ImageTools `denoise_pdps`, restated in SURVEY.md §8a row a3)
 *   - the named assumptions S1–S9 of docs/SEMANTICS.md.
 * It is pinned instead by the property tests in tests/test_oracle_*.py
 * (adjointness, fixed point / duality gap, λ→0 limit, an independent numpy
 * restatement, finite differences of the cost).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdlib.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Step-size recursion (S1, S2): σ=σ₀/R_K, τ=τ₀/R_K, γ=1; per iteration
 * ω = accel ? 1/√(1+2γτ) : 1, used for this iteration's over-relaxation,
 * then τ←τω, σ←σ/ω.  out[3*k+{0,1,2}] = (τ_k, σ_k, ω_k).  All in fp64, exactly
 * the operation order of the Julia expressions.                              */
void oracle_step_sizes(double tau0, double sigma0, double opnorm, int accel,
                       int maxiter, double *out)
{
    double sigma = sigma0 / opnorm;
    double tau = tau0 / opnorm;
    const double gamma = 1.0;
    for (int k = 0; k < maxiter; ++k) {
        double omega = 1.0;
        if (accel) {
            double t = 2.0 * gamma; t = t * tau; t = 1.0 + t;
            omega = 1.0 / sqrt(t);
        }
        out[3 * k] = tau; out[3 * k + 1] = sigma; out[3 * k + 2] = omega;
        if (accel) { tau = tau * omega; sigma = sigma / omega; }
    }
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* In-place lower banded Cholesky with a pivot floor — the factorisation of oracle.gradient_dual (the CPU checker of the
 * multiplier-space adjoint systems, /root/reference/src/TVLearningFunctionVec.jl:98-161 after eliminating p).
 * ab is LAPACK's lower band storage as numpy holds it, row-major (bw+1) × Nd: ab[r*Nd + j] = A[j+r, j].  The same
 * right-looking column sweep, operation for operation, as the numpy loop it replaces (oracle.py: _chol_band_guard_py),
 * on a column-contiguous copy; the columns a pivot column updates are independent and shared among the threads.
 * Returns the number of pivots raised to `guard`. */
long long oracle_chol_band_guard(double *ab, int bw, long long Nd, double guard)
{
    const long long ld = (long long)bw + 1;
    double *c = (double *)malloc((size_t)(ld * Nd) * sizeof(double));
    if (!c) return -1;
    for (long long j = 0; j < Nd; ++j)
        for (long long r = 0; r < ld; ++r) c[j * ld + r] = ab[r * Nd + j];
    long long guarded = 0;
    for (long long j = 0; j < Nd; ++j) {
        double *l = c + j * ld;
        double d = l[0];
        if (!(d > guard)) { d = guard; ++guarded; }
        d = sqrt(d);
        l[0] = d;
        const long long m = (long long)bw < Nd - 1 - j ? (long long)bw : Nd - 1 - j;
        for (long long r = 1; r <= m; ++r) l[r] = l[r] / d;
#pragma omp parallel for schedule(static) if (m >= 96)
        for (long long b = 1; b <= m; ++b) {
            const double lb = l[b];
            if (lb == 0.0) continue;
            double *t = c + (j + b) * ld;
            for (long long r = 0; r <= m - b; ++r) t[r] -= l[b + r] * lb;
        }
    }
    for (long long j = 0; j < Nd; ++j)
        for (long long r = 0; r < ld; ++r) ab[r * Nd + j] = c[j * ld + r];
    free(c);
    return guarded;
}

#define REAL double
#define SUF f64
#define SQRT sqrt
#include "pdps_body.inc"
#undef REAL
#undef SUF
#undef SQRT

#define REAL float
#define SUF f32
#define SQRT sqrtf
#include "pdps_body.inc"
#undef REAL
#undef SUF
#undef SQRT
