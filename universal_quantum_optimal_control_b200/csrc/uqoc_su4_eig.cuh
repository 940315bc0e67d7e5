// Two-qubit SU(4) path, "eigenframe" kernel (sm_100a) -- the default SU(4) kernel.
//
// NOT IN THE REFERENCE: builder-defined contract, see uqoc_su4.cu.
//
// Static disorder makes every pulse of one error sample the SAME Hamiltonian up to a local z-frame:
//   H_k = R_k H' R_k^dagger,   R_k = exp(-i phi1_k ZI/2) exp(-i phi2_k IZ/2)   (diagonal phases),
//   2H' = XI + IX + d1 ZI + d2 IZ + J ZZ                                        (REAL symmetric, pulse-independent).
// So 2H' = V diag(mu) V^T is diagonalised ONCE per sample (cyclic Jacobi in double, ~4 pulses' worth of
// work) and each pulse is exact and cheap:
//   U_k = R_k V D_k V^T R_k^dagger,   D_k = diag exp(-i mu_m tau_k (1+eps) / 2).
// With the state kept in the pulse's own frame, Q_k = R_k^dagger P_k,
//   Q_k = V D_k V^T G_k Q_{k-1},   G_k = R_k^dagger R_{k-1} = diag exp(+-i (dphi1 +- dphi2)/2)   (pulse-only),
// i.e. two real-times-complex 4x4 products and two diagonal phase multiplications: 384 FMA-pipe operations
// + 4 sin/cos per (pulse, sample) instead of ~1700 for the scaling-and-squaring exponential.
//
// Backward: the Pade kernel's B_k = P_k W_k^dagger enters every gradient only through Im Tr(B_k X) with X
// Hermitian, so only the Hermitian matrix A_k = (B_k - B_k^dagger)/(2i) is carried (16 reals).  In frame k
//   dF/dtau_k  = (1+eps)/2 sum_m mu_m (V^T A_k V)_mm,
//   dF/dphi1_k = 1/2 [Tr(A_k ZI) - Tr(X_k ZI)],   X_k = V D_k^dagger (V^T A_k V) D_k V^T,
//   A_{k-1} = G_k^dagger X_k G_k.
// Conjugation by the REAL V splits into a symmetric and an antisymmetric real problem (176 operations
// instead of 512); the diagonal conjugations touch the 6 off-diagonal pairs only.
// FP32: V is rounded from the double Jacobi result, the eigenvalues are carried as hi + lo floats (the
// rounding of mu would otherwise be a systematic phase error multiplied by L), and Q_L gets one
// Newton-Schulz step (removes, to first order, the norm drift caused by V_f^T V_f = I + O(6e-8)).
#pragma once
#include "uqoc_su2_kernels.cuh"
#include "uqoc_su2_x2.cuh"   // F2 / fma2 / mul2 / f2b packed-FP32 helpers

namespace uqoc {

// ------------------------------------------------------------------ Jacobi eigensolver, 4x4 real symmetric
// A = 2H' (d1, d2, J);  on exit V[i][m] = component i of eigenvector m, mu[m] its eigenvalue.
// Fully unrolled (all indices compile-time); the sweep loop is warp-uniform (callers are converged).
__device__ __forceinline__ void su4_jacobi(double d1, double d2, double J, double (&V)[4][4], double (&mu)[4]) {
    double A[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            A[i][j] = 0.0;
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    A[0][0] = d1 + d2 + J; A[1][1] = d1 - d2 - J; A[2][2] = -d1 + d2 - J; A[3][3] = -d1 - d2 + J;
    A[0][1] = A[1][0] = 1.0; A[2][3] = A[3][2] = 1.0;      // IX
    A[0][2] = A[2][0] = 1.0; A[1][3] = A[3][1] = 1.0;      // XI
    const double scale = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2] + A[3][3] * A[3][3] + 8.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[0][3] * A[0][3] + A[1][2] * A[1][2] +
                           A[1][3] * A[1][3] + A[2][3] * A[2][3];
        if (__all_sync(0xffffffffu, off <= 1e-32 * scale)) break;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const double apq = A[p][q];
                if (fabs(apq) > 1e-290) {
                    // rotation by the SMALL angle (|theta| <= pi/4) that zeroes A[p][q], division-free:
                    // cos 2theta = |d|/r, c^2 = (1 + cos 2theta)/2, s = sign(d) apq / (r c), t = s/c
                    const double d = A[q][q] - A[p][p];
                    const double inv_r = ::rsqrt(::fma(d, d, 4.0 * apq * apq));
                    const double c2 = ::fma(0.5 * fabs(d), inv_r, 0.5);
                    const double inv_c = ::rsqrt(c2);
                    const double c = c2 * inv_c;
                    const double s = copysign(apq, apq * d) * inv_r * inv_c;
                    const double t = s * inv_c;
                    A[p][p] -= t * apq;
                    A[q][q] += t * apq;
                    A[p][q] = A[q][p] = 0.0;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (r != p && r != q) {
                            const double arp = A[r][p], arq = A[r][q];
                            A[r][p] = A[p][r] = c * arp - s * arq;
                            A[r][q] = A[q][r] = s * arp + c * arq;
                        }
                        const double vrp = V[r][p], vrq = V[r][q];
                        V[r][p] = c * vrp - s * vrq;
                        V[r][q] = s * vrp + c * vrq;
                    }
                }
            }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) mu[m] = A[m][m];
}

// per-sample data of the eigenframe kernel
template <typename T>
struct Su4Frame {
    T V[4][4];       // V[i][m]
    T lam[4];        // mu/2 (hi part in FP32)
    T lam_lo[4];     // FP32: mu/2 - float(mu/2);  FP64: 0
    T dl[3];         // lam[m] - lam[m+1]  (the backward sweep only needs phase DIFFERENCES)
    T dl_lo[3];
    T te;            // 1 + eps
    // isoclinic factors of V in SO(4) = (SU(2) x SU(2))/Z2:  V K V^T for antisymmetric K = sum alpha_a L_a +
    // sum beta_b R_b rotates alpha by Rl and beta by Rr; symmetric S = t/4 I + sum N_ab L_a R_b goes to Rl N Rr^T
    // (L_a / R_b = left / right multiplication by the quaternion units).  hRr = Rr / 2.
    T Rl[3][3], hRr[3][3];
};

template <typename T>
__device__ __forceinline__ void su4_make_frame(Su4Frame<T>& f, T d1, T d2, T eps, T J) {
    double V[4][4], mu[4];
    su4_jacobi((double)d1, (double)d2, (double)J, V, mu);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) f.V[i][m] = (T)V[i][m];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const double l = 0.5 * mu[m];
        f.lam[m] = (T)l;
        f.lam_lo[m] = (T)(l - (double)f.lam[m]);
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const double l = 0.5 * (mu[m] - mu[m + 1]);
        f.dl[m] = (T)l;
        f.dl_lo[m] = (T)(l - (double)f.dl[m]);
    }
    f.te = (T)1 + eps;
    // Rot[c][a] = 1/4 <U_c, V U_a V^T> with U = L or R; both are antisymmetric with two upper non-zeros (i, j, sign)
    constexpr int UL[3][2][3] = {{{0, 1, -1}, {2, 3, -1}}, {{0, 2, -1}, {1, 3, 1}}, {{0, 3, -1}, {1, 2, -1}}};
    constexpr int UR[3][2][3] = {{{0, 1, -1}, {2, 3, 1}}, {{0, 2, -1}, {1, 3, -1}}, {{0, 3, -1}, {1, 2, 1}}};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double rl = 0.0, rr = 0.0;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                {
                    const int i = UL[c][e][0], j = UL[c][e][1];
                    double m = 0.0;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const int pp = UL[a][g][0], qq = UL[a][g][1];
                        m += UL[a][g][2] * (V[i][pp] * V[j][qq] - V[i][qq] * V[j][pp]);
                    }
                    rl += UL[c][e][2] * m;
                }
                {
                    const int i = UR[c][e][0], j = UR[c][e][1];
                    double m = 0.0;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const int pp = UR[a][g][0], qq = UR[a][g][1];
                        m += UR[a][g][2] * (V[i][pp] * V[j][qq] - V[i][qq] * V[j][pp]);
                    }
                    rr += UR[c][e][2] * m;
                }
            }
            f.Rl[c][a] = (T)(0.5 * rl);
            f.hRr[c][a] = (T)(0.25 * rr);
        }
}

// full-sign sin/cos of the eigenphases from the shared-memory table {sin[1024] | cos[1024]} of (k pi/1024).
// FP32: first-order residual like the packed SU(2) kernel (angle error <= 1.2e-9; the radial error <= 1.2e-6 has
// zero mean and is Hermitian to first order, so the Newton-Schulz step on Q_L removes it); 6 FMA-pipe
// instructions + 2 LDS, the sign (-1)^(k div 1024) restored by an integer XOR.  FP64: SinCos<double, SC_TABLE>.
__device__ __forceinline__ void su4_sincos(float h, float& s, float& c, const float* __restrict__ tab) {
    const float MAGIC = 12582912.0f;
    float kf = fmaf(h, 325.94931f, MAGIC);
    const int k = __float_as_int(kf);
    kf -= MAGIC;
    float r = fmaf(kf, -0.003067961661145091f, h);
    r = fmaf(kf, 8.537380524753502e-11f, r);
    const float st = tab[k & 1023], ct = tab[1024 + (k & 1023)];
    const int sg = (k << 21) & 0x80000000;
    s = __int_as_float(__float_as_int(fmaf(r, ct, st)) ^ sg);
    c = __int_as_float(__float_as_int(fmaf(-r, st, ct)) ^ sg);
}
__device__ __forceinline__ void su4_sincos(double h, double& s, double& c, const double* __restrict__ tab) {
    int kb;
    SinCos<double, SC_TABLE>::eval(h, s, c, kb, tab);
    if (kb & 1) {
        s = -s;
        c = -c;
    }
}

template <typename T>
__device__ __forceinline__ void su4_phases(const Su4Frame<T>& f, T tau, T (&c)[4], T (&s)[4], const T* __restrict__ tab) {
    const T t = tau * f.te;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        T h;
        if constexpr (sizeof(T) == 4) h = fmaf(f.lam[m], t, f.lam_lo[m] * t);
        else h = f.lam[m] * t;
        su4_sincos(h, s[m], c[m], tab);
    }
}
// w[pair(m,n)] = e^{i(h_m - h_n)}, m < n: three table look-ups (adjacent differences) and three complex products
template <typename T>
__device__ __forceinline__ void su4_phase_diffs(const Su4Frame<T>& f, T tau, T (&wr)[6], T (&wi)[6], const T* __restrict__ tab) {
    const T t = tau * f.te;
    T c[3], s[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        T h;
        if constexpr (sizeof(T) == 4) h = fmaf(f.dl[m], t, f.dl_lo[m] * t);
        else h = f.dl[m] * t;
        su4_sincos(h, s[m], c[m], tab);
    }
    wr[0] = c[0]; wi[0] = s[0];                                   // (0,1)
    wr[3] = c[1]; wi[3] = s[1];                                   // (1,2)
    wr[5] = c[2]; wi[5] = s[2];                                   // (2,3)
    wr[1] = c[0] * c[1] - s[0] * s[1]; wi[1] = c[0] * s[1] + s[0] * c[1];       // (0,2) = (0,1)(1,2)
    wr[4] = c[1] * c[2] - s[1] * s[2]; wi[4] = c[1] * s[2] + s[1] * c[2];       // (1,3) = (1,2)(2,3)
    wr[2] = wr[1] * c[2] - wi[1] * s[2]; wi[2] = wr[1] * s[2] + wi[1] * c[2];   // (0,3) = (0,2)(2,3)
}

// ------------------------------------------------------------------ forward step  Q <- V D V^T G Q
// g = {cos a, sin a, cos b, sin b}: G = diag(e^{ia}, e^{ib}, e^{-ib}, e^{-ia}), a = (dphi1+dphi2)/2, b = (dphi1-dphi2)/2
template <typename T>
__device__ __forceinline__ void su4e_fwd_step(T (&qr)[4][4], T (&qi)[4][4], const Su4Frame<T>& f, const T (&c)[4],
                                              const T (&s)[4], T ca, T sa, T cb, T sb) {
    T xr[4][4], xi[4][4];
    // G Q (row phases)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        xr[0][j] = ca * qr[0][j] - sa * qi[0][j]; xi[0][j] = ca * qi[0][j] + sa * qr[0][j];
        xr[1][j] = cb * qr[1][j] - sb * qi[1][j]; xi[1][j] = cb * qi[1][j] + sb * qr[1][j];
        xr[2][j] = cb * qr[2][j] + sb * qi[2][j]; xi[2][j] = cb * qi[2][j] - sb * qr[2][j];
        xr[3][j] = ca * qr[3][j] + sa * qi[3][j]; xi[3][j] = ca * qi[3][j] - sa * qr[3][j];
    }
    // Y = V^T X, then D Y (row m times c_m - i s_m)
    T yr[4][4], yi[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const T ar = f.V[0][m] * xr[0][j] + f.V[1][m] * xr[1][j] + f.V[2][m] * xr[2][j] + f.V[3][m] * xr[3][j];
            const T ai = f.V[0][m] * xi[0][j] + f.V[1][m] * xi[1][j] + f.V[2][m] * xi[2][j] + f.V[3][m] * xi[3][j];
            yr[m][j] = c[m] * ar + s[m] * ai;
            yi[m][j] = c[m] * ai - s[m] * ar;
        }
    // Q = V Y
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            qr[i][j] = f.V[i][0] * yr[0][j] + f.V[i][1] * yr[1][j] + f.V[i][2] * yr[2][j] + f.V[i][3] * yr[3][j];
            qi[i][j] = f.V[i][0] * yi[0][j] + f.V[i][1] * yi[1][j] + f.V[i][2] * yi[2][j] + f.V[i][3] * yi[3][j];
        }
}

// FP32 forward step with two COLUMNS of Q per 64-bit register pair (FFMA2): every product has a per-sample or
// per-pulse scalar on one side, which rides in the 32-bit broadcast operand form, so the 384 FMA-pipe lane
// operations of a step issue as 192 instructions (the scalar kernel is issue-bound: ncu 80 % issue-active).
// qr[i][jp], qi[i][jp]: row i, column pair jp = (2jp, 2jp+1).
__device__ __forceinline__ void su4e_fwd_step_x2(F2 (&qr)[4][2], F2 (&qi)[4][2], const Su4Frame<float>& f, const float (&c)[4],
                                                 const float (&s)[4], float ca, float sa, float cb, float sb) {
    // Operand order matters: an FFMA2 that reads two fresh register pairs sustains 76 % of the pipe, one that
    // reads one fresh pair (the other from the operand-reuse cache, the scalar in the 32-bit broadcast slot)
    // 95 % (tools/ubench/fma_ubench.cu modes 5 / 6).  So every product below is written "outer-product" style:
    // four consecutive instructions share the SAME packed multiplicand and update four different accumulators.
    const F2 Ca = f2b(ca), Sa = f2b(sa), Cb = f2b(cb), Sb = f2b(sb);
    F2 xr[4][2], xi[4][2];
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const F2 Cg = (i == 0 || i == 3) ? Ca : Cb;
            const F2 Sg = (i == 0) ? Sa : (i == 1) ? Sb : (i == 2) ? neg2(Sb) : neg2(Sa);   // e^{+i..} rows 0,1; e^{-i..} rows 2,3
            const F2 t1 = mul2(Cg, qr[i][jp]), t2 = mul2(Sg, qr[i][jp]);
            xr[i][jp] = fma2(neg2(Sg), qi[i][jp], t1);
            xi[i][jp] = fma2(Cg, qi[i][jp], t2);
        }
    }
    // Y = V^T X
    F2 ar[4][2], ai[4][2];
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {
#pragma unroll
        for (int m = 0; m < 4; ++m) ar[m][jp] = mul2(f2b(f.V[0][m]), xr[0][jp]);
#pragma unroll
        for (int m = 0; m < 4; ++m) ai[m][jp] = mul2(f2b(f.V[0][m]), xi[0][jp]);
#pragma unroll
        for (int i = 1; i < 4; ++i) {
#pragma unroll
            for (int m = 0; m < 4; ++m) ar[m][jp] = fma2(f2b(f.V[i][m]), xr[i][jp], ar[m][jp]);
#pragma unroll
            for (int m = 0; m < 4; ++m) ai[m][jp] = fma2(f2b(f.V[i][m]), xi[i][jp], ai[m][jp]);
        }
    }
    // D Y (row m times c_m - i s_m)
    F2 yr[4][2], yi[4][2];
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const F2 cm = f2b(c[m]), sm = f2b(s[m]);
            const F2 t1 = mul2(cm, ar[m][jp]), t2 = mul2(neg2(sm), ar[m][jp]);
            yr[m][jp] = fma2(sm, ai[m][jp], t1);
            yi[m][jp] = fma2(cm, ai[m][jp], t2);
        }
    }
    // Q = V Y
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {
#pragma unroll
        for (int i = 0; i < 4; ++i) qr[i][jp] = mul2(f2b(f.V[i][0]), yr[0][jp]);
#pragma unroll
        for (int i = 0; i < 4; ++i) qi[i][jp] = mul2(f2b(f.V[i][0]), yi[0][jp]);
#pragma unroll
        for (int m = 1; m < 4; ++m) {
#pragma unroll
            for (int i = 0; i < 4; ++i) qr[i][jp] = fma2(f2b(f.V[i][m]), yr[m][jp], qr[i][jp]);
#pragma unroll
            for (int i = 0; i < 4; ++i) qi[i][jp] = fma2(f2b(f.V[i][m]), yi[m][jp], qi[i][jp]);
        }
    }
}

// ------------------------------------------------------------------ Hermitian 4x4 in packed form
// A = S + iK, S symmetric (dg + re), K antisymmetric (im, upper triangle).  Pair index of (i<j):
// (0,1)=0 (0,2)=1 (0,3)=2 (1,2)=3 (1,3)=4 (2,3)=5.
template <typename T>
struct Herm4 {
    T dg[4], re[6], im[6];
};
__host__ __device__ constexpr int su4_pair(int i, int j) { return i == 0 ? j - 1 : (i == 1 ? j + 1 : 5); }

// sym[i][j] / asym[i][j] accessors with compile-time indices
template <typename T>
__device__ __forceinline__ T herm_s(const Herm4<T>& A, int i, int j) {
    return i == j ? A.dg[i] : (i < j ? A.re[su4_pair(i, j)] : A.re[su4_pair(j, i)]);
}
template <typename T>
__device__ __forceinline__ T herm_k(const Herm4<T>& A, int i, int j) {
    return i == j ? (T)0 : (i < j ? A.im[su4_pair(i, j)] : -A.im[su4_pair(j, i)]);
}

// E = W^T A W with W[i][m] = TR ? V[m][i] : V[i][m]   (TR = false: into the eigenbasis; true: back out)
template <typename T, bool TR>
__device__ __forceinline__ void herm_conj_real(Herm4<T>& E, const Herm4<T>& A, const Su4Frame<T>& f) {
    T ys[4][4], yk[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            T a = (T)0, b = (T)0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const T w = TR ? f.V[m][j] : f.V[j][m];
                a += herm_s(A, i, j) * w;
                if (j != i) b += herm_k(A, i, j) * w;
            }
            ys[i][m] = a;
            yk[i][m] = b;
        }
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = m; n < 4; ++n) {
            T a = (T)0, b = (T)0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const T w = TR ? f.V[m][i] : f.V[i][m];
                a += w * ys[i][n];
                if (n != m) b += w * yk[i][n];
            }
            if (n == m) E.dg[m] = a;
            else {
                E.re[su4_pair(m, n)] = a;
                E.im[su4_pair(m, n)] = b;
            }
        }
}

// E = W^T A W (TR = false, W = V: into the eigenbasis) or W A W^T (TR = true: back out) through the isoclinic
// coordinates: elements -> (alpha, beta, N) by add/sub butterflies, three 3x3 rotations (72 FMA), butterflies back.
// ~117 FMA-pipe operations instead of the 176 of herm_conj_real.  tq = Tr(A)/4 (invariant of the whole sweep).
template <typename T, bool TR>
__device__ __forceinline__ void herm_conj_iso(Herm4<T>& E, const Herm4<T>& A, const Su4Frame<T>& f, T tq) {
    // pair index: (0,1)=0 (0,2)=1 (0,3)=2 (1,2)=3 (1,3)=4 (2,3)=5;  re = S off-diagonal, im = K
    const T k01 = A.im[0], k02 = A.im[1], k03 = A.im[2], k12 = A.im[3], k13 = A.im[4], k23 = A.im[5];
    const T s01 = A.re[0], s02 = A.re[1], s03 = A.re[2], s12 = A.re[3], s13 = A.re[4], s23 = A.re[5];
    const T al[3] = {(T)-0.5 * (k01 + k23), (T)0.5 * (k13 - k02), (T)-0.5 * (k03 + k12)};
    const T b2[3] = {k23 - k01, -(k02 + k13), k12 - k03};                                  // 2 beta
    const T u0 = A.dg[0] + A.dg[1], u1 = A.dg[0] - A.dg[1], u2 = A.dg[2] + A.dg[3], u3 = A.dg[2] - A.dg[3];
    const T N2[3][3] = {{(T)0.5 * (u2 - u0), s03 - s12, -(s02 + s13)},
                        {-(s03 + s12), (T)-0.5 * (u1 + u3), s01 - s23},
                        {s02 - s13, -(s01 + s23), (T)0.5 * (u3 - u1)}};                     // 2 N
    T ao[3], bo[3], Y[3][3], N[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T a = (T)0, b = (T)0;
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            a += (TR ? f.Rl[c][e] : f.Rl[e][c]) * al[e];
            b += (TR ? f.hRr[c][e] : f.hRr[e][c]) * b2[e];
        }
        ao[c] = a;
        bo[c] = b;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            T y = (T)0;
#pragma unroll
            for (int e = 0; e < 3; ++e) y += N2[a][e] * (TR ? f.hRr[b][e] : f.hRr[e][b]);
            Y[a][b] = y;
        }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            T n = (T)0;
#pragma unroll
            for (int e = 0; e < 3; ++e) n += (TR ? f.Rl[a][e] : f.Rl[e][a]) * Y[e][b];
            N[a][b] = n;
        }
    E.im[0] = -ao[0] - bo[0]; E.im[1] = -ao[1] - bo[1]; E.im[2] = -ao[2] - bo[2];
    E.im[3] = bo[2] - ao[2];  E.im[4] = ao[1] - bo[1];  E.im[5] = bo[0] - ao[0];
    const T pp = N[1][1] + N[2][2], mm = N[1][1] - N[2][2], cm = tq - N[0][0], cp = tq + N[0][0];
    E.dg[0] = cm - pp; E.dg[1] = cm + pp; E.dg[2] = cp - mm; E.dg[3] = cp + mm;
    E.re[0] = N[1][2] - N[2][1]; E.re[1] = N[2][0] - N[0][2]; E.re[2] = N[0][1] - N[1][0];
    E.re[3] = -N[0][1] - N[1][0]; E.re[4] = -N[0][2] - N[2][0]; E.re[5] = -N[1][2] - N[2][1];
}

// ------------------------------------------------------------------ kernel
constexpr int kSu4eThreads = 64;
constexpr int kSu4eWarps = kSu4eThreads / 32;

// shared memory: sin/cos table (2 x 1024) + per pulse {fw[4], tau, b1[4], b2[4]} + target' (32) + last-frame
// phases (8) + scratch + acc
template <typename T>
__host__ __device__ inline size_t su4e_smem_bytes(int L, bool bwd) {
    size_t n = 2048 + (size_t)L * (bwd ? 13 : 5) + 32 + 8 + kSu4eWarps;
    if (bwd) n += (size_t)kSu4eWarps * L * 3;
    return n * sizeof(T) + 16;
}

// FP32: 6 resident blocks per SM (<= 170 registers) measured best; FP64 needs the full register file
#ifndef UQOC_SU4E_MINB
#define UQOC_SU4E_MINB 6
#endif
// COT (backward only): the adjoint is seeded from a per-sample cotangent dLoss/dU (p.cot) instead of the fidelity
// against the target -- the backward of the strict generator signature (uqoc_su4_generator_backward).
template <typename T, bool BWD, bool COT>
__global__ void __launch_bounds__(kSu4eThreads, sizeof(T) == 4 ? UQOC_SU4E_MINB : 1) su4e_kernel(const Su4Params<T> p) {
    static_assert(BWD || !COT, "the cotangent seed belongs to the backward kernel");
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int L = p.L;
    T* tab = reinterpret_cast<T*>(smem_raw);                // {sin[1024] | cos[1024]} of k pi/1024
    T* fw = tab + 2048;                                     // [L][4]  cos a, sin a, cos b, sin b
    T* b1 = fw + (size_t)L * 4;                             // [L][4]  cos dphi1, sin dphi1, cos dphi2, sin dphi2
    T* b2 = b1 + (BWD ? (size_t)L * 4 : 0);                 // [L][4]  cos(d1+d2), sin(d1+d2), cos(d1-d2), sin(d1-d2)
    T* tauv = b2 + (BWD ? (size_t)L * 4 : 0);               // [L]
    T* tgt = tauv + L;                                      // [32]    T' = R_L^dagger T
    T* rl = tgt + 32;                                       // [8]     last-frame phases e^{i gamma_i}
    T* scratch = rl + 8;                                    // [warps]
    T* acc = scratch + kSu4eWarps;                          // [warps][L][3]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int split = blockIdx.x % p.splits, b = blockIdx.x / p.splits;
    {
        // table first: 16 independent loads per thread in flight while the pulse trigonometry runs
        T tv[2][16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if constexpr (sizeof(T) == 4) {
                tv[0][u] = g_sin_table[tid + u * kSu4eThreads];
                tv[1][u] = g_cos_table[tid + u * kSu4eThreads];
            } else {
                tv[0][u] = g_sin_table_f64[tid + u * kSu4eThreads];
                tv[1][u] = g_cos_table_f64[tid + u * kSu4eThreads];
            }
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            tab[tid + u * kSu4eThreads] = tv[0][u];
            tab[1024 + tid + u * kSu4eThreads] = tv[1][u];
        }
        const T* pb = p.pulses + (size_t)b * L * 3;
        for (int i = tid; i < L; i += kSu4eThreads) {
            const double p1 = (double)pb[3 * i], p2 = (double)pb[3 * i + 1];
            const double q1 = i > 0 ? (double)pb[3 * i - 3] : 0.0, q2 = i > 0 ? (double)pb[3 * i - 2] : 0.0;
            const double e1 = p1 - q1, e2 = p2 - q2;
            double sa, ca, sb, cb;
            ::sincos(0.5 * (e1 + e2), &sa, &ca);
            ::sincos(0.5 * (e1 - e2), &sb, &cb);
            fw[4 * i] = (T)ca; fw[4 * i + 1] = (T)sa; fw[4 * i + 2] = (T)cb; fw[4 * i + 3] = (T)sb;
            tauv[i] = pb[3 * i + 2];
            if (BWD) {
                // e^{i dphi1} = e^{ia} e^{ib}, e^{i dphi2} = e^{ia} e^{-ib}, e^{i(d1+d2)} = e^{2ia}, e^{i(d1-d2)} = e^{2ib}
                b1[4 * i] = (T)(ca * cb - sa * sb); b1[4 * i + 1] = (T)(sa * cb + ca * sb);
                b1[4 * i + 2] = (T)(ca * cb + sa * sb); b1[4 * i + 3] = (T)(sa * cb - ca * sb);
                b2[4 * i] = (T)(ca * ca - sa * sa); b2[4 * i + 1] = (T)(2.0 * sa * ca);
                b2[4 * i + 2] = (T)(cb * cb - sb * sb); b2[4 * i + 3] = (T)(2.0 * sb * cb);
            }
        }
        if (tid < 16) {
            // T'[i][j] = e^{+i gamma_i} T[i][j], gamma = ((p1+p2)/2, (p1-p2)/2, -(p1-p2)/2, -(p1+p2)/2) of the LAST pulse
            const int i = tid >> 2;
            const double p1 = (double)pb[3 * (L - 1)], p2 = (double)pb[3 * (L - 1) + 1];
            const double gam = (i == 0) ? 0.5 * (p1 + p2) : (i == 1) ? 0.5 * (p1 - p2) : (i == 2) ? -0.5 * (p1 - p2) : -0.5 * (p1 + p2);
            double sg, cg;
            ::sincos(gam, &sg, &cg);
            if constexpr (!COT) {
                const double tr_ = (double)p.target[(size_t)b * 32 + 2 * tid], ti_ = (double)p.target[(size_t)b * 32 + 2 * tid + 1];
                tgt[2 * tid] = (T)(cg * tr_ - sg * ti_);
                tgt[2 * tid + 1] = (T)(cg * ti_ + sg * tr_);
            } else {
                tgt[2 * tid] = tgt[2 * tid + 1] = (T)0;
            }
            if ((tid & 3) == 0) {
                rl[2 * i] = (T)cg;
                rl[2 * i + 1] = (T)sg;
            }
        }
        if (BWD)
            for (int i = tid; i < kSu4eWarps * L * 3; i += kSu4eThreads) acc[i] = (T)0;
    }
    __syncthreads();
    const size_t Bm = (size_t)p.B * p.M;
    T fsum = (T)0;
    for (int tile = split; tile < p.n_tiles; tile += p.splits) {
        const long long j = (long long)tile * kSu4eThreads + tid;
        const bool valid = j < p.M;
        const size_t sidx = (size_t)b * p.M + (size_t)(valid ? j : 0);
        Su4Frame<T> f;
        {
            T d1 = (T)0, d2 = (T)0, eps = (T)0;
            if (valid) {
                if (p.err != nullptr) {
                    d1 = p.err[sidx]; d2 = p.err[Bm + sidx]; eps = p.err[2 * Bm + sidx];
                } else {
                    unsigned long long seed; unsigned offset;
                    su4_rng_state(p, seed, offset);
                    philox_su4<T>((uint64_t)(p.j0 + j), (uint32_t)b, seed, offset, p.sig_d, p.sig_e, d1, d2, eps);
                }
                if (p.err_out != nullptr) {
                    p.err_out[sidx] = d1; p.err_out[Bm + sidx] = d2; p.err_out[2 * Bm + sidx] = eps;
                }
            }
            su4_make_frame<T>(f, d1, d2, eps, p.J);
        }
        // ---------------- forward ----------------
        T qr[4][4], qi[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                qr[i][jj] = (i == jj) ? (T)1 : (T)0;
                qi[i][jj] = (T)0;
            }
        if constexpr (sizeof(T) == 4) {
            F2 pr[4][2], pi[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jp = 0; jp < 2; ++jp) {
                    pr[i][jp] = f2(qr[i][2 * jp], qr[i][2 * jp + 1]);
                    pi[i][jp] = f2b(0.0f);
                }
            for (int k = 0; k < L; ++k) {
                float c[4], s[4];
                su4_phases<float>(f, tauv[k], c, s, tab);
                su4e_fwd_step_x2(pr, pi, f, c, s, fw[4 * k], fw[4 * k + 1], fw[4 * k + 2], fw[4 * k + 3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jp = 0; jp < 2; ++jp) {
                    qr[i][2 * jp] = f2lo(pr[i][jp]); qr[i][2 * jp + 1] = f2hi(pr[i][jp]);
                    qi[i][2 * jp] = f2lo(pi[i][jp]); qi[i][2 * jp + 1] = f2hi(pi[i][jp]);
                }
        } else {
            for (int k = 0; k < L; ++k) {
                T c[4], s[4];
                su4_phases<T>(f, tauv[k], c, s, tab);
                su4e_fwd_step<T>(qr, qi, f, c, s, fw[4 * k], fw[4 * k + 1], fw[4 * k + 2], fw[4 * k + 3]);
            }
        }
        if constexpr (sizeof(T) == 4) {
            // one Newton-Schulz step  Q <- Q (3I - Q^dagger Q)/2
            T nr[4][4], ni[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    T ar = (T)0, ai = (T)0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ar += qr[k][i] * qr[k][jj] + qi[k][i] * qi[k][jj];
                        ai += qr[k][i] * qi[k][jj] - qi[k][i] * qr[k][jj];
                    }
                    nr[i][jj] = (T)0.5 * (((i == jj) ? (T)1 : (T)0) - ar);
                    ni[i][jj] = (T)-0.5 * ai;
                }
            T ur[4][4], ui[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    T ar = qr[i][jj], ai = qi[i][jj];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ar += qr[i][k] * nr[k][jj] - qi[i][k] * ni[k][jj];
                        ai += qr[i][k] * ni[k][jj] + qi[i][k] * nr[k][jj];
                    }
                    ur[i][jj] = ar;
                    ui[i][jj] = ai;
                }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    qr[i][jj] = ur[i][jj];
                    qi[i][jj] = ui[i][jj];
                }
        }
        T trr = (T)0, tri = (T)0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const T t_r = tgt[2 * (4 * i + jj)], t_i = tgt[2 * (4 * i + jj) + 1];
                trr += qr[i][jj] * t_r + qi[i][jj] * t_i;
                tri += qr[i][jj] * t_i - qi[i][jj] * t_r;
            }
        const T F = (trr * trr + tri * tri + (T)4) * (T)0.05;
        if (valid) {
            fsum += F;
            if (p.F_out != nullptr) p.F_out[sidx] = F;
            if (!BWD && p.U_out != nullptr) {
                // P_L = R_L Q_L: row i times e^{-i gamma_i}
                T* U = p.U_out + sidx * 32;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const T cg = rl[2 * i], sg = rl[2 * i + 1];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        U[2 * (4 * i + jj)] = cg * qr[i][jj] + sg * qi[i][jj];
                        U[2 * (4 * i + jj) + 1] = cg * qi[i][jj] - sg * qr[i][jj];
                    }
                }
            }
        }
        if constexpr (BWD) {
            // C = (tr/10) w Q T'^dagger;  A = (C - C^dagger)/(2i):  S = (Ci + Ci^T)/2,  K = -(Cr - Cr^T)/2
            T wgt = (T)0;
            if (valid) wgt = p.weight != nullptr ? p.weight[sidx] : (T)1;
            Herm4<T> A;
            {
                T cr_[4][4], ci_[4][4];
                if constexpr (COT) {
                    // C = Q gU'^dagger with gU' = R_L^dagger gU (row i times e^{+i gamma_i}); an invalid slot contributes 0
                    const T* gu = p.cot + sidx * 32;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const T cg = rl[2 * jj], sg = rl[2 * jj + 1];
                        T t_r[4], t_i[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const T g_r = valid ? gu[2 * (4 * jj + k)] : (T)0, g_i = valid ? gu[2 * (4 * jj + k) + 1] : (T)0;
                            t_r[k] = cg * g_r - sg * g_i;
                            t_i[k] = cg * g_i + sg * g_r;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            T mr = (T)0, mi = (T)0;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                mr += qr[i][k] * t_r[k] + qi[i][k] * t_i[k];
                                mi += qi[i][k] * t_r[k] - qr[i][k] * t_i[k];
                            }
                            cr_[i][jj] = mr;
                            ci_[i][jj] = mi;
                        }
                    }
                } else {
                    const T fr = wgt * trr * (T)0.1, fi = wgt * tri * (T)0.1;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            T mr = (T)0, mi = (T)0;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const T t_r = tgt[2 * (4 * jj + k)], t_i = tgt[2 * (4 * jj + k) + 1];
                                mr += qr[i][k] * t_r + qi[i][k] * t_i;
                                mi += qi[i][k] * t_r - qr[i][k] * t_i;
                            }
                            cr_[i][jj] = fr * mr - fi * mi;
                            ci_[i][jj] = fr * mi + fi * mr;
                        }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    A.dg[i] = ci_[i][i];
#pragma unroll
                    for (int jj = i + 1; jj < 4; ++jj) {
                        A.re[su4_pair(i, jj)] = (T)0.5 * (ci_[i][jj] + ci_[jj][i]);
                        A.im[su4_pair(i, jj)] = (T)-0.5 * (cr_[i][jj] - cr_[jj][i]);
                    }
                }
            }
            const T tq = (T)0.25 * (A.dg[0] + A.dg[1] + A.dg[2] + A.dg[3]);       // Tr A / 4: invariant of the sweep
            for (int k = L - 1; k >= 0; --k) {
                T wr[6], wi[6];
                su4_phase_diffs<T>(f, tauv[k], wr, wi, tab);
                const T z1a = A.dg[0] + A.dg[1] - A.dg[2] - A.dg[3];
                const T z2a = A.dg[0] - A.dg[1] + A.dg[2] - A.dg[3];
                Herm4<T> E;
                herm_conj_iso<T, false>(E, A, f, tq);
                T g_tau = f.te * (E.dg[0] * f.lam[0] + E.dg[1] * f.lam[1] + E.dg[2] * f.lam[2] + E.dg[3] * f.lam[3]);
                // E'_mn = E_mn e^{i(h_m - h_n)}
#pragma unroll
                for (int pq = 0; pq < 6; ++pq) {
                    const T er = E.re[pq], ei = E.im[pq];
                    E.re[pq] = er * wr[pq] - ei * wi[pq];
                    E.im[pq] = er * wi[pq] + ei * wr[pq];
                }
                Herm4<T> X;
                herm_conj_iso<T, true>(X, E, f, tq);
                const T z1b = X.dg[0] + X.dg[1] - X.dg[2] - X.dg[3];
                const T z2b = X.dg[0] - X.dg[1] + X.dg[2] - X.dg[3];
                T g_p1 = (T)0.5 * (z1a - z1b), g_p2 = (T)0.5 * (z2a - z2b);
                // A_{k-1} = G^dagger X G:  X_ij e^{-i(gamma_i - gamma_j)}
                {
                    const T c1 = b1[4 * k], s1 = b1[4 * k + 1], c2 = b1[4 * k + 2], s2 = b1[4 * k + 3];
                    const T cp = b2[4 * k], sp = b2[4 * k + 1], cm = b2[4 * k + 2], sm = b2[4 * k + 3];
                    const T pc[6] = {c2, c1, cp, cm, c1, c2};
                    const T ps[6] = {s2, s1, sp, sm, s1, s2};
#pragma unroll
                    for (int e = 0; e < 4; ++e) A.dg[e] = X.dg[e];
#pragma unroll
                    for (int e = 0; e < 6; ++e) {
                        A.re[e] = X.re[e] * pc[e] + X.im[e] * ps[e];
                        A.im[e] = X.im[e] * pc[e] - X.re[e] * ps[e];
                    }
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    g_p1 += __shfl_xor_sync(0xffffffffu, g_p1, d);
                    g_p2 += __shfl_xor_sync(0xffffffffu, g_p2, d);
                    g_tau += __shfl_xor_sync(0xffffffffu, g_tau, d);
                }
                if (lane == 0) {
                    T* dst = acc + ((size_t)warp * L + k) * 3;
                    dst[0] += g_p1; dst[1] += g_p2; dst[2] += g_tau;
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
    if (lane == 0) scratch[warp] = fsum;
    __syncthreads();
    if (tid == 0 && p.Fsum_part != nullptr) {
        T tot = (T)0;
#pragma unroll
        for (int w = 0; w < kSu4eWarps; ++w) tot += scratch[w];
        p.Fsum_part[(size_t)split * p.B + b] = tot;
    }
    if constexpr (BWD) {
        T* gout = p.G_part + ((size_t)split * p.B + b) * L * 3;
        for (int i = tid; i < 3 * L; i += kSu4eThreads) {
            T tot = (T)0;
#pragma unroll
            for (int w = 0; w < kSu4eWarps; ++w) tot += acc[(size_t)w * L * 3 + i];
            gout[i] = tot;
        }
    }
}

}  // namespace uqoc
