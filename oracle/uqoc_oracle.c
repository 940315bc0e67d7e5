/*
 * uqoc_oracle.c -- plain-C restatement of the disorder-sampled SU(2) propagation + fidelity path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may build, load or call this file, and there
 * only as the checker or as the timed CPU port.  The product (universal_quantum_optimal_control_b200)
 * never links it and has no CPU fallback.
 *
 * Third, independent formulation (the CUDA kernels use quaternions + a 3-vector adjoint, oracle/uqoc_oracle.py
 * uses numpy complex matrices with prefix AND suffix products): here plain complex 2x2 arithmetic, stored
 * prefixes and a running suffix.  Reference lines followed (paths relative to the upstream repository):
 *   SCORE.py   = train/unitary_single_qubit_gate/universal_single_qubit_SCORE.py
 *   trainer.py = model/universal_model_trainer.py
 *
 *   H_i   = 1/2 (1+eps) (cos(phi_i) X + sin(phi_i) Y + delta Z)                      SCORE.py:117-124
 *   U_i   = exp(-i H_i tau_i)  (closed form of torch.linalg.matrix_exp for a traceless Hermitian 2x2)  SCORE.py:127
 *   U_out = U_L ... U_1                                                              SCORE.py:131-145
 *   F     = (|Tr(U_out^dagger U_target)|^2 + 2) / 6                                  SCORE.py:168-183
 *   sample s = b*M + j uses pulses[b], target[b], error[:, s]                        trainer.py:80-82
 * Pinned against tests/golden/ (outputs of the unmodified reference) by tests/test_oracle_c.py.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/liboracle_c.so oracle/uqoc_oracle.c -lm
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

typedef struct {
    cplx a, b, c, d; /* [[a, b], [c, d]] */
} m2;

static inline m2 m2_mul(m2 x, m2 y) {
    m2 r = {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
    return r;
}

/* U = cos(h) I - i sin(h) (nx X + ny Y + nz Z),  h = tau (1+eps) w / 2,  n = (cos phi, sin phi, delta) / w */
static inline m2 pulse_unitary(double phi, double tau, double delta, double eps, double *h_out, double *w_out) {
    const double w = sqrt(1.0 + delta * delta);
    const double h = 0.5 * tau * (1.0 + eps) * w;
    const double c = cos(h), s = sin(h);
    const double nx = cos(phi) / w, ny = sin(phi) / w, nz = delta / w;
    m2 U = {c - I * s * nz, -I * s * (nx - I * ny), -I * s * (nx + I * ny), c + I * s * nz};
    *h_out = h;
    *w_out = w;
    return U;
}

/* Tr(A^dagger B) */
static inline cplx tr_adag_b(m2 A, m2 B) { return conj(A.a) * B.a + conj(A.b) * B.b + conj(A.c) * B.c + conj(A.d) * B.d; }

/*
 * pulses (B, L, 2) [phi, tau]; target (B, 2, 2) complex interleaved (re, im); err (2, B*M) [delta; eps].
 * Outputs (any may be NULL): F (B*M) per-sample fidelity, Fsum (B) = sum_j F, grad (B, L, 2) = sum_j dF/d[phi, tau],
 * U_out (B*M, 2, 2) complex interleaved.  Returns 0, or -1 on allocation failure.
 */
int uqoc_c_su2(const double *pulses, const double *target, const double *err, int64_t B, int64_t L, int64_t M, double *F,
               double *Fsum, double *grad, double *U_out) {
    const int64_t Bm = B * M;
    int fail = 0;
    if (Fsum) memset(Fsum, 0, (size_t)B * sizeof(double));
    if (grad) memset(grad, 0, (size_t)(B * L * 2) * sizeof(double));
#pragma omp parallel
    {
        m2 *prefix = (m2 *)malloc((size_t)(L + 1) * sizeof(m2));
        m2 *Us = (m2 *)malloc((size_t)L * sizeof(m2));
        double *hs = (double *)malloc((size_t)L * sizeof(double));
        double *gl = grad ? (double *)calloc((size_t)(L * 2), sizeof(double)) : NULL;
        if (!prefix || !Us || !hs || (grad && !gl)) {
#pragma omp atomic write
            fail = 1;
        } else {
            /* work item = (target b, chunk of <= 256 samples): parallel over both axes so that one-target
             * workloads (BASELINE config 3) use every core; partial sums are added atomically (the order of the
             * floating-point additions varies at the 1e-16 level, irrelevant for a checker). */
            const int64_t CH = 256, nch = (M + CH - 1) / CH;
#pragma omp for schedule(dynamic, 1)
            for (int64_t item = 0; item < B * nch; ++item) {
                const int64_t b = item / nch, j_lo = (item % nch) * CH, j_hi = (j_lo + CH < M) ? j_lo + CH : M;
                const double *pb = pulses + b * L * 2;
                const double *tb = target + b * 8;
                const m2 T = {tb[0] + I * tb[1], tb[2] + I * tb[3], tb[4] + I * tb[5], tb[6] + I * tb[7]};
                double fs = 0.0;
                if (gl) memset(gl, 0, (size_t)(L * 2) * sizeof(double));
                for (int64_t j = j_lo; j < j_hi; ++j) {
                    const int64_t s_idx = b * M + j;
                    const double delta = err[s_idx], eps = err[Bm + s_idx];
                    double w = 1.0;
                    m2 P = {1.0, 0.0, 0.0, 1.0};
                    prefix[0] = P;
                    for (int64_t i = 0; i < L; ++i) {
                        Us[i] = pulse_unitary(pb[2 * i], pb[2 * i + 1], delta, eps, &hs[i], &w);
                        P = m2_mul(Us[i], P); /* later pulses on the left */
                        prefix[i + 1] = P;
                    }
                    const cplx tr = tr_adag_b(P, T);
                    const double Fv = (creal(tr) * creal(tr) + cimag(tr) * cimag(tr) + 2.0) / 6.0;
                    fs += Fv;
                    if (F) F[s_idx] = Fv;
                    if (U_out) {
                        double *u = U_out + s_idx * 8;
                        u[0] = creal(P.a); u[1] = cimag(P.a); u[2] = creal(P.b); u[3] = cimag(P.b);
                        u[4] = creal(P.c); u[5] = cimag(P.c); u[6] = creal(P.d); u[7] = cimag(P.d);
                    }
                    if (gl) {
                        /* dF = (1/3) Re(conj(tr) dtr),  dtr = Tr((S dU_i P_{i-1})^dagger T),  S = U_L ... U_{i+1} */
                        m2 S = {1.0, 0.0, 0.0, 1.0};
                        const double a_half = 0.5 * (1.0 + eps) * w;
                        for (int64_t i = L - 1; i >= 0; --i) {
                            const double phi = pb[2 * i];
                            const double nx = cos(phi) / w, ny = sin(phi) / w, nz = delta / w;
                            /* dU/dtau = -i a (n.sigma) U_i */
                            const m2 ns = {nz, nx - I * ny, nx + I * ny, -nz};
                            m2 dUt = m2_mul(ns, Us[i]);
                            dUt.a *= -I * a_half; dUt.b *= -I * a_half; dUt.c *= -I * a_half; dUt.d *= -I * a_half;
                            /* dU/dphi = -i sin(h) (dn/dphi . sigma),  dn/dphi = (-sin phi, cos phi, 0) / w */
                            const double sh = sin(hs[i]);
                            const double dx = -sin(phi) / w, dy = cos(phi) / w;
                            const m2 dUp = {0.0, -I * sh * (dx - I * dy), -I * sh * (dx + I * dy), 0.0};
                            const m2 At = m2_mul(m2_mul(S, dUt), prefix[i]);
                            const m2 Ap = m2_mul(m2_mul(S, dUp), prefix[i]);
                            const cplx dtr_t = tr_adag_b(At, T), dtr_p = tr_adag_b(Ap, T);
                            gl[2 * i] += (creal(tr) * creal(dtr_p) + cimag(tr) * cimag(dtr_p)) / 3.0;
                            gl[2 * i + 1] += (creal(tr) * creal(dtr_t) + cimag(tr) * cimag(dtr_t)) / 3.0;
                            S = m2_mul(S, Us[i]);
                        }
                    }
                }
                if (Fsum) {
#pragma omp atomic
                    Fsum[b] += fs;
                }
                if (grad)
                    for (int64_t e = 0; e < L * 2; ++e) {
#pragma omp atomic
                        grad[b * L * 2 + e] += gl[e];
                    }
            }
        }
        free(prefix);
        free(Us);
        free(hs);
        free(gl);
    }
    return fail ? -1 : 0;
}

int uqoc_c_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
