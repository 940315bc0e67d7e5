// Two-qubit SU(4) eigenframe kernel with the pulse train SPLIT OVER THE FOUR WARPS of a block (few samples, e.g. one
// target x 32768 samples = BASELINE config 4: the one-sample-per-thread kernel of uqoc_su4_eig.cuh then fills 512 of 888
// resident 64-thread block slots, 1.7 warps per scheduler).  NOT IN THE REFERENCE.
//
// The four warps of a 128-thread block share 32 samples; warp w owns pulses [w C, (w+1) C).  With the state in the
// pulse's own frame the chunk products X_w = prod_{k in chunk} (V D_k V^T G_k) multiply up to Q_L = X_3 X_2 X_1 X_0
// exactly (G_k of a chunk's first pulse is the frame change from the previous chunk's last pulse).  They are exchanged
// through shared memory; every warp forms the prefix P_w = X_w .. X_0 and Q_L, and seeds ITS backward sweep with
//   A_(end of chunk w) = P_w (Q_L^dagger A_L Q_L) P_w^dagger
// (B_k = P_k W_k^dagger = P_k (P_L^dagger B_L P_L) P_k^dagger for unitary prefixes; the Hermitian part conjugates the
// same way) -- the prefix alone, no suffix products, as in the SU(2) kernels.  Each warp writes a disjoint range of pulse
// gradients.  Same per-step arithmetic as su4e_kernel (su4e_fwd_step[_x2], herm_conj_iso, su4_phase_diffs).
#pragma once
#include "uqoc_su4_eig.cuh"

namespace uqoc {

constexpr int kSu4sWarps = 4;
constexpr int kSu4sThreads = 32 * kSu4sWarps;

template <typename T>
__host__ __device__ inline size_t su4s_smem_bytes(int L) {
    // table + per pulse {fw[4], b1[4], b2[4], tau} + target' (32) + last-frame phases (8) + scratch + acc (L x 3) + chunk products
    const size_t n = 2048 + (size_t)L * 13 + 32 + 8 + kSu4sWarps + (size_t)L * 3 + (size_t)kSu4sWarps * 32 * 32;
    return n * sizeof(T) + 16;
}

// C = A B for complex 4x4 in split real / imaginary arrays
template <typename T>
__device__ __forceinline__ void c4_mul(T (&cr)[4][4], T (&ci)[4][4], const T (&ar)[4][4], const T (&ai)[4][4], const T (&br)[4][4],
                                       const T (&bi)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            T xr = (T)0, xi = (T)0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                xr += ar[i][k] * br[k][j] - ai[i][k] * bi[k][j];
                xi += ar[i][k] * bi[k][j] + ai[i][k] * br[k][j];
            }
            cr[i][j] = xr;
            ci[i][j] = xi;
        }
}
// full complex matrix of a packed Hermitian one, and back (the anti-Hermitian rounding residue is dropped)
template <typename T>
__device__ __forceinline__ void herm_unpack(const Herm4<T>& A, T (&hr)[4][4], T (&hi)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hr[i][j] = herm_s(A, i, j);
            hi[i][j] = herm_k(A, i, j);
        }
}
template <typename T>
__device__ __forceinline__ void herm_pack(Herm4<T>& A, const T (&hr)[4][4], const T (&hi)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        A.dg[i] = hr[i][i];
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            A.re[su4_pair(i, j)] = (T)0.5 * (hr[i][j] + hr[j][i]);
            A.im[su4_pair(i, j)] = (T)0.5 * (hi[i][j] - hi[j][i]);
        }
    }
}
// E = Q A Q^dagger (DAG = false) or Q^dagger A Q (DAG = true), A Hermitian
template <typename T, bool DAG>
__device__ __forceinline__ void herm_conj_unitary(Herm4<T>& E, const Herm4<T>& A, const T (&qr)[4][4], const T (&qi)[4][4]) {
    T hr[4][4], hi[4][4], ur[4][4], ui[4][4], tr_[4][4], ti_[4][4];
    herm_unpack(A, hr, hi);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {                       // U = DAG ? Q^dagger : Q
            ur[i][j] = DAG ? qr[j][i] : qr[i][j];
            ui[i][j] = DAG ? -qi[j][i] : qi[i][j];
        }
    c4_mul(tr_, ti_, ur, ui, hr, hi);                       // T = U A
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {                       // E = T U^dagger
            T xr = (T)0, xi = (T)0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                xr += tr_[i][k] * ur[j][k] + ti_[i][k] * ui[j][k];
                xi += ti_[i][k] * ur[j][k] - tr_[i][k] * ui[j][k];
            }
            hr[i][j] = xr;
            hi[i][j] = xi;
        }
    herm_pack(E, hr, hi);
}

#ifndef UQOC_SU4S_MINB
#define UQOC_SU4S_MINB 3
#endif
template <typename T, bool BWD>
__global__ void __launch_bounds__(kSu4sThreads, sizeof(T) == 4 ? UQOC_SU4S_MINB : 1) su4e_split_kernel(const Su4Params<T> p) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int L = p.L;
    const int C = (L + kSu4sWarps - 1) / kSu4sWarps;        // pulses per warp
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* fw = tab + 2048;
    T* b1 = fw + (size_t)L * 4;
    T* b2 = b1 + (size_t)L * 4;
    T* tauv = b2 + (size_t)L * 4;
    T* tgt = tauv + L;
    T* rl = tgt + 32;
    T* scratch = rl + 8;
    T* acc = scratch + kSu4sWarps;                          // [L][3]: the warps write disjoint pulse ranges
    T* xq = acc + (size_t)L * 3;                            // [warps][32 reals][32 lanes]
    UQOC_ASSERT((size_t)(reinterpret_cast<unsigned char*>(xq + (size_t)kSu4sWarps * 32 * 32) - smem_raw) <= dyn_smem_bytes());
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int split = blockIdx.x % p.splits, b = blockIdx.x / p.splits;
    {
        for (int i = tid; i < 1024; i += kSu4sThreads) {
            if constexpr (sizeof(T) == 4) {
                tab[i] = g_sin_table[i];
                tab[1024 + i] = g_cos_table[i];
            } else {
                tab[i] = g_sin_table_f64[i];
                tab[1024 + i] = g_cos_table_f64[i];
            }
        }
        const T* pb = p.pulses + (size_t)b * L * 3;
        for (int i = tid; i < L; i += kSu4sThreads) {
            const double p1 = (double)pb[3 * i], p2 = (double)pb[3 * i + 1];
            const double q1 = i > 0 ? (double)pb[3 * i - 3] : 0.0, q2 = i > 0 ? (double)pb[3 * i - 2] : 0.0;
            const double e1 = p1 - q1, e2 = p2 - q2;
            double sa, ca, sb, cb;
            ::sincos(0.5 * (e1 + e2), &sa, &ca);
            ::sincos(0.5 * (e1 - e2), &sb, &cb);
            fw[4 * i] = (T)ca; fw[4 * i + 1] = (T)sa; fw[4 * i + 2] = (T)cb; fw[4 * i + 3] = (T)sb;
            tauv[i] = pb[3 * i + 2];
            b1[4 * i] = (T)(ca * cb - sa * sb); b1[4 * i + 1] = (T)(sa * cb + ca * sb);
            b1[4 * i + 2] = (T)(ca * cb + sa * sb); b1[4 * i + 3] = (T)(sa * cb - ca * sb);
            b2[4 * i] = (T)(ca * ca - sa * sa); b2[4 * i + 1] = (T)(2.0 * sa * ca);
            b2[4 * i + 2] = (T)(cb * cb - sb * sb); b2[4 * i + 3] = (T)(2.0 * sb * cb);
        }
        if (tid < 16) {
            const int i = tid >> 2;
            const double p1 = (double)pb[3 * (L - 1)], p2 = (double)pb[3 * (L - 1) + 1];
            const double gam = (i == 0) ? 0.5 * (p1 + p2) : (i == 1) ? 0.5 * (p1 - p2) : (i == 2) ? -0.5 * (p1 - p2) : -0.5 * (p1 + p2);
            double sg, cg;
            ::sincos(gam, &sg, &cg);
            const double tr_ = (double)p.target[(size_t)b * 32 + 2 * tid], ti_ = (double)p.target[(size_t)b * 32 + 2 * tid + 1];
            tgt[2 * tid] = (T)(cg * tr_ - sg * ti_);
            tgt[2 * tid + 1] = (T)(cg * ti_ + sg * tr_);
            if ((tid & 3) == 0) {
                rl[2 * i] = (T)cg;
                rl[2 * i + 1] = (T)sg;
            }
        }
        if (BWD)
            for (int i = tid; i < L * 3; i += kSu4sThreads) acc[i] = (T)0;
    }
    __syncthreads();
    const size_t Bm = (size_t)p.B * p.M;
    const int k_lo = warp * C, k_hi = (k_lo + C < L) ? k_lo + C : L;       // this warp's pulses [k_lo, k_hi) (may be empty)
    T fsum = (T)0;
    for (int tile = split; tile < p.n_tiles; tile += p.splits) {
        const long long j = (long long)tile * 32 + lane;
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
                if (p.err_out != nullptr && warp == 0) {
                    p.err_out[sidx] = d1; p.err_out[Bm + sidx] = d2; p.err_out[2 * Bm + sidx] = eps;
                }
            }
            su4_make_frame<T>(f, d1, d2, eps, p.J);
        }
        // ---------------- forward over this warp's chunk ----------------
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
            for (int k = k_lo; k < k_hi; ++k) {
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
            for (int k = k_lo; k < k_hi; ++k) {
                T c[4], s[4];
                su4_phases<T>(f, tauv[k], c, s, tab);
                su4e_fwd_step<T>(qr, qi, f, c, s, fw[4 * k], fw[4 * k + 1], fw[4 * k + 2], fw[4 * k + 3]);
            }
        }
        // ---------------- chunk products -> prefix at this chunk's end (P) and the full product (Q_L) ----------------
        {
            T* dst = xq + (size_t)warp * 32 * 32 + lane;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    dst[(size_t)(2 * (4 * i + jj)) * 32] = qr[i][jj];
                    dst[(size_t)(2 * (4 * i + jj) + 1) * 32] = qi[i][jj];
                }
        }
        __syncthreads();
        T pr_[4][4], pi_[4][4];                             // prefix P_w (inclusive of this warp's chunk)
        {
            T rr[4][4], ri[4][4];
#pragma unroll
            for (int w2 = 0; w2 < kSu4sWarps; ++w2) {
                const T* src = xq + (size_t)w2 * 32 * 32 + lane;
                T xr[4][4], xi[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        xr[i][jj] = src[(size_t)(2 * (4 * i + jj)) * 32];
                        xi[i][jj] = src[(size_t)(2 * (4 * i + jj) + 1) * 32];
                    }
                if (w2 == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            rr[i][jj] = xr[i][jj];
                            ri[i][jj] = xi[i][jj];
                        }
                } else {
                    T nr[4][4], ni[4][4];
                    c4_mul(nr, ni, xr, xi, rr, ri);         // later pulses on the left
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            rr[i][jj] = nr[i][jj];
                            ri[i][jj] = ni[i][jj];
                        }
                }
                if (w2 == warp) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            pr_[i][jj] = rr[i][jj];
                            pi_[i][jj] = ri[i][jj];
                        }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    qr[i][jj] = rr[i][jj];                  // Q_L
                    qi[i][jj] = ri[i][jj];
                }
        }
        __syncthreads();                                    // xq is rewritten by the next tile
        if constexpr (sizeof(T) == 4) {
            // one Newton-Schulz step  Q <- Q (3I - Q^dagger Q)/2 on the full product (as su4e_kernel)
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
            c4_mul(ur, ui, qr, qi, nr, ni);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    qr[i][jj] += ur[i][jj];
                    qi[i][jj] += ui[i][jj];
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
        if (valid && warp == 0) {
            fsum += F;
            if (p.F_out != nullptr) p.F_out[sidx] = F;
        }
        if constexpr (BWD) {
            T wgt = (T)0;
            if (valid) wgt = p.weight != nullptr ? p.weight[sidx] : (T)1;
            Herm4<T> A;
            {
                const T fr = wgt * trr * (T)0.1, fi = wgt * tri * (T)0.1;
                T cr_[4][4], ci_[4][4];
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
            // A at the end of this warp's chunk: P (Q_L^dagger A_L Q_L) P^dagger  (the last warp's is A_L itself up to rounding)
            {
                Herm4<T> Mh;
                herm_conj_unitary<T, true>(Mh, A, qr, qi);
                herm_conj_unitary<T, false>(A, Mh, pr_, pi_);
            }
            const T tq = (T)0.25 * (A.dg[0] + A.dg[1] + A.dg[2] + A.dg[3]);
            for (int k = k_hi - 1; k >= k_lo; --k) {
                T wr[6], wi[6];
                su4_phase_diffs<T>(f, tauv[k], wr, wi, tab);
                const T z1a = A.dg[0] + A.dg[1] - A.dg[2] - A.dg[3];
                const T z2a = A.dg[0] - A.dg[1] + A.dg[2] - A.dg[3];
                Herm4<T> E;
                herm_conj_iso<T, false>(E, A, f, tq);
                T g_tau = f.te * (E.dg[0] * f.lam[0] + E.dg[1] * f.lam[1] + E.dg[2] * f.lam[2] + E.dg[3] * f.lam[3]);
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
                    T* dst = acc + (size_t)k * 3;
                    dst[0] += g_p1; dst[1] += g_p2; dst[2] += g_tau;
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
    __syncthreads();
    if (tid == 0 && p.Fsum_part != nullptr) p.Fsum_part[(size_t)split * p.B + b] = fsum;     // warp 0 holds the block's sum
    if constexpr (BWD) {
        T* gout = p.G_part + ((size_t)split * p.B + b) * L * 3;
        for (int i = tid; i < 3 * L; i += kSu4sThreads) gout[i] = acc[i];
    }
}

}  // namespace uqoc
