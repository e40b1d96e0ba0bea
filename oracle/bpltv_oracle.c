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
