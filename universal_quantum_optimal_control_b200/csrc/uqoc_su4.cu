// Two-qubit SU(4) disorder-sampled propagation + fidelity, forward and backward (sm_100a).
//
// NOT IN THE REFERENCE (README.md:86,122 promise train/two_qubit/, which does not exist; SURVEY.md
// §8a row A9).  Builder-defined, same callable contract as the SU(2) path; parity is pinned only
// against oracle/uqoc_oracle.py::su4_* (eigendecomposition exponentials), NOT against the reference.
//
//   pulses (B, L, 3) = [phi1, phi2, tau],  error (3, B*M) = [delta1; delta2; eps]
//   H = 1/2 [cos phi1 XI + sin phi1 YI + cos phi2 IX + sin phi2 IY + delta1 ZI + delta2 IZ + J ZZ]
//   U_k = exp(-i H_k tau_k (1+eps)),  U = U_L ... U_1,  F = (|Tr(U^dagger T)|^2 + 4) / 20
//
// One thread owns one error sample; the 4x4 complex matrices live in registers (fully unrolled).
// exp(-iG): scaling and squaring around the [m/0] Pade (Taylor) approximant in Horner form, the
// scaling exponent chosen warp-uniformly from the exact bound ||G|| <= t/2 (|a1| + |a2| + |J|).
// Backward needs no Frechet derivative: rotating phi1 is conjugation by exp(-i phi1 ZI/2), so
//   dU_k/dphi1 = -(i/2) [ZI, U_k],  dU_k/dphi2 = -(i/2) [IZ, U_k],  dU_k/dtau = -i (1+eps) H_k U_k
// and with B_k = P_k W_k^dagger (prefix times adjoint), B_{k-1} = U_k^dagger B_k U_k:
//   dF/dphi1_k = 1/2 Im[Tr(B_k ZI) - Tr(B_{k-1} ZI)],  dF/dtau_k = (1+eps) Im Tr(B_k H_k).
// So the backward sweep is: recompute U_k, two 4x4 products, a few traces; state = one matrix.
#include <mutex>
#include <unordered_map>
#include "uqoc_su2_kernels.cuh"

namespace uqoc {

template <typename T>
struct M4 {
    T re[16], im[16];
};

template <typename T>
__device__ __forceinline__ void m4_identity(M4<T>& A) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        A.re[i] = (i % 5 == 0) ? (T)1 : (T)0;
        A.im[i] = (T)0;
    }
}
// C = A B
template <typename T>
__device__ __forceinline__ void m4_mul(M4<T>& C, const M4<T>& A, const M4<T>& Bm) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            T cr = (T)0, ci = (T)0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                cr += A.re[4 * i + k] * Bm.re[4 * k + j] - A.im[4 * i + k] * Bm.im[4 * k + j];
                ci += A.re[4 * i + k] * Bm.im[4 * k + j] + A.im[4 * i + k] * Bm.re[4 * k + j];
            }
            C.re[4 * i + j] = cr;
            C.im[4 * i + j] = ci;
        }
}
// C = A^dagger B
template <typename T>
__device__ __forceinline__ void m4_mul_hA(M4<T>& C, const M4<T>& A, const M4<T>& Bm) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            T cr = (T)0, ci = (T)0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                cr += A.re[4 * k + i] * Bm.re[4 * k + j] + A.im[4 * k + i] * Bm.im[4 * k + j];
                ci += A.re[4 * k + i] * Bm.im[4 * k + j] - A.im[4 * k + i] * Bm.re[4 * k + j];
            }
            C.re[4 * i + j] = cr;
            C.im[4 * i + j] = ci;
        }
}
// C = A B^dagger
template <typename T>
__device__ __forceinline__ void m4_mul_hB(M4<T>& C, const M4<T>& A, const M4<T>& Bm) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            T cr = (T)0, ci = (T)0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                cr += A.re[4 * i + k] * Bm.re[4 * j + k] + A.im[4 * i + k] * Bm.im[4 * j + k];
                ci += A.im[4 * i + k] * Bm.re[4 * j + k] - A.re[4 * i + k] * Bm.im[4 * j + k];
            }
            C.re[4 * i + j] = cr;
            C.im[4 * i + j] = ci;
        }
}

template <typename T>
struct Su4Traits;
template <>
struct Su4Traits<float> {
    static constexpr int kDeg = 6;                 // |Y|^7/7! = 1.2e-8 at |Y| = 1/4
    __device__ static float theta() { return 0.25f; }
};
template <>
struct Su4Traits<double> {
    static constexpr int kDeg = 12;                // |Y|^13/13! = 2.4e-18 at |Y| = 1/4
    __device__ static double theta() { return 0.25; }
};

template <typename T>
struct Su4Sample {
    T d1, d2, te, nrm;   // detunings, (1+eps), ||H|| bound * (1+eps)
};

// E = exp(-i t H(phi1, phi2, d1, d2, J)),  t = tau * (1+eps).  c1,s1,c2,s2 = cos/sin of the phases.
template <typename T>
__device__ __forceinline__ void su4_pulse_exp(M4<T>& E, T c1, T s1, T c2, T s2, T tau, const Su4Sample<T>& k, T J) {
    constexpr int DEG = Su4Traits<T>::kDeg;
    const T t = tau * k.te;
    // warp-uniform scaling exponent from the norm bound
    const T nrm = fabs(tau) * k.nrm;
    int s = 0;
    {
        T x = nrm;
        while (x > Su4Traits<T>::theta() && s < 30) {
            x *= (T)0.5;
            ++s;
        }
        s = __reduce_max_sync(0xffffffffu, s);
    }
    const T sc = t * (T)0.5 / (T)(1 << s);         // Y = -i sc * (2H)
    // Y = -i sc Hh,  Hh = 2H = diag(d1+d2+J, d1-d2-J, -d1+d2-J, -d1-d2+J) + phase blocks.  Y has 12 non-zeros:
    // a purely imaginary diagonal i*dg[] and four distinct off-diagonal values
    //   a2 = -i sc e^{-i phi2} at (0,1),(2,3);  b2 = -i sc e^{+i phi2} at (1,0),(3,2)
    //   a1 = -i sc e^{-i phi1} at (0,2),(1,3);  b1 = -i sc e^{+i phi1} at (2,0),(3,1)
    const T dg0 = -sc * (k.d1 + k.d2 + J), dg1 = -sc * (k.d1 - k.d2 - J), dg2 = -sc * (-k.d1 + k.d2 - J),
            dg3 = -sc * (-k.d1 - k.d2 + J);
    const T a2r = -sc * s2, b2r = sc * s2, o2i = -sc * c2;     // a2 = a2r + i o2i, b2 = b2r + i o2i
    const T a1r = -sc * s1, b1r = sc * s1, o1i = -sc * c1;
    // Horner: E = I + Y/1 (I + Y/2 (I + ... (I + Y/DEG))), every step a SPARSE product Y*E (160 FMA, not 256)
    {
        const T inv = (T)(1.0 / DEG);
#pragma unroll
        for (int i = 0; i < 16; ++i) E.re[i] = E.im[i] = (T)0;
        E.re[0] = E.re[5] = E.re[10] = E.re[15] = (T)1;
        E.im[0] = dg0 * inv; E.im[5] = dg1 * inv; E.im[10] = dg2 * inv; E.im[15] = dg3 * inv;
        E.re[1] = a2r * inv;  E.im[1] = o2i * inv;   E.re[11] = a2r * inv; E.im[11] = o2i * inv;
        E.re[4] = b2r * inv;  E.im[4] = o2i * inv;   E.re[14] = b2r * inv; E.im[14] = o2i * inv;
        E.re[2] = a1r * inv;  E.im[2] = o1i * inv;   E.re[7] = a1r * inv;  E.im[7] = o1i * inv;
        E.re[8] = b1r * inv;  E.im[8] = o1i * inv;   E.re[13] = b1r * inv; E.im[13] = o1i * inv;
    }
    M4<T> Tm;
#pragma unroll
    for (int d = DEG - 1; d >= 1; --d) {
        const T inv = (T)(1.0 / d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const T e0r = E.re[j], e0i = E.im[j], e1r = E.re[4 + j], e1i = E.im[4 + j];
            const T e2r = E.re[8 + j], e2i = E.im[8 + j], e3r = E.re[12 + j], e3i = E.im[12 + j];
            // row0 = i dg0 E0 + a2 E1 + a1 E2
            Tm.re[j] = -dg0 * e0i + (a2r * e1r - o2i * e1i) + (a1r * e2r - o1i * e2i);
            Tm.im[j] = dg0 * e0r + (a2r * e1i + o2i * e1r) + (a1r * e2i + o1i * e2r);
            // row1 = b2 E0 + i dg1 E1 + a1 E3
            Tm.re[4 + j] = (b2r * e0r - o2i * e0i) - dg1 * e1i + (a1r * e3r - o1i * e3i);
            Tm.im[4 + j] = (b2r * e0i + o2i * e0r) + dg1 * e1r + (a1r * e3i + o1i * e3r);
            // row2 = b1 E0 + i dg2 E2 + a2 E3
            Tm.re[8 + j] = (b1r * e0r - o1i * e0i) - dg2 * e2i + (a2r * e3r - o2i * e3i);
            Tm.im[8 + j] = (b1r * e0i + o1i * e0r) + dg2 * e2r + (a2r * e3i + o2i * e3r);
            // row3 = b1 E1 + b2 E2 + i dg3 E3
            Tm.re[12 + j] = (b1r * e1r - o1i * e1i) + (b2r * e2r - o2i * e2i) - dg3 * e3i;
            Tm.im[12 + j] = (b1r * e1i + o1i * e1r) + (b2r * e2i + o2i * e2r) + dg3 * e3r;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            E.re[i] = Tm.re[i] * inv + ((i % 5 == 0) ? (T)1 : (T)0);
            E.im[i] = Tm.im[i] * inv;
        }
    }
    for (int q = 0; q < s; ++q) {
        m4_mul(Tm, E, E);
        E = Tm;
    }
}

template <typename T>
struct Su4Params {
    const T* pulses;   // (B, L, 3)
    const T* target;   // (B, 4, 4, 2)
    const T* err;      // (3, B*M) or nullptr
    const T* weight;   // (B*M) or nullptr
    int B, L, M, n_tiles, splits;
    long long j0;
    T J, sig_d, sig_e;
    unsigned long long seed;
    unsigned offset;
    T* U_out;      // (B*M, 4, 4, 2) or nullptr
    T* F_out;      // (B*M)
    T* err_out;    // (3, B*M)
    T* Fsum_part;  // [splits][B]
    T* G_part;     // [splits][B][L][3]
    const unsigned long long* rng_dev;   // non-null: {seed, offset} read from device memory (CUDA-graph replay)
    const T* cot;  // non-null (eigenframe backward only): per-sample cotangent dLoss/dU (B*M, 4, 4, 2) seeds the adjoint
};
// Philox (seed, offset) of a launch: immediate values or the device-resident pair of UQOC_FLAG_RNG_FROM_DEVICE
template <typename T>
__device__ __forceinline__ void su4_rng_state(const Su4Params<T>& p, unsigned long long& seed, unsigned& offset) {
    seed = p.rng_dev != nullptr ? p.rng_dev[0] : p.seed;
    offset = p.rng_dev != nullptr ? (unsigned)p.rng_dev[1] : p.offset;
}

template <typename T>
__device__ __forceinline__ void philox_su4(uint64_t j, uint32_t b, uint64_t seed, uint32_t offset, T sig_d, T sig_e,
                                           T& d1, T& d2, T& eps);
template <>
__device__ __forceinline__ void philox_su4<float>(uint64_t j, uint32_t b, uint64_t seed, uint32_t offset, float sig_d,
                                                  float sig_e, float& d1, float& d2, float& eps) {
    const Philox4 x = philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), b, offset, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float U = 2.3283064365386963e-10f, H = 1.1641532182693481e-10f;
    const float r0 = sqrtf(-2.0f * logf(fmaf((float)x.x, U, H))), r1 = sqrtf(-2.0f * logf(fmaf((float)x.z, U, H)));
    float s0, c0, s1, c1;
    sincospif(2.0f * fmaf((float)x.y, U, H), &s0, &c0);
    sincospif(2.0f * fmaf((float)x.w, U, H), &s1, &c1);
    d1 = sig_d * r0 * c0;
    eps = sig_e * r0 * s0;
    d2 = sig_d * r1 * c1;
}
template <>
__device__ __forceinline__ void philox_su4<double>(uint64_t j, uint32_t b, uint64_t seed, uint32_t offset, double sig_d,
                                                   double sig_e, double& d1, double& d2, double& eps) {
    const Philox4 x = philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), b, offset, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double U = 2.3283064365386963e-10;
    const double r0 = ::sqrt(-2.0 * ::log(((double)x.x + 0.5) * U)), r1 = ::sqrt(-2.0 * ::log(((double)x.z + 0.5) * U));
    double s0, c0, s1, c1;
    sincospi(2.0 * ((double)x.y + 0.5) * U, &s0, &c0);
    sincospi(2.0 * ((double)x.w + 0.5) * U, &s1, &c1);
    d1 = sig_d * r0 * c0;
    eps = sig_e * r0 * s0;
    d2 = sig_d * r1 * c1;
}

}  // namespace uqoc
#include "uqoc_su4_eig.cuh"
#include "uqoc_su4_split.cuh"
namespace uqoc {

constexpr int kSu4Threads = 64;
constexpr int kSu4Warps = kSu4Threads / 32;

template <typename T>
__host__ __device__ inline size_t su4_smem_bytes(int L, bool bwd) {
    size_t bytes = (size_t)L * 5 * sizeof(T) + 32 * sizeof(T) + (size_t)kSu4Warps * sizeof(T);
    if (bwd) bytes += (size_t)kSu4Warps * L * 3 * sizeof(T);
    return bytes;
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(kSu4Threads) su4_kernel(const Su4Params<T> p) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    T* tab = reinterpret_cast<T*>(smem_raw);            // [L][5] = cos phi1, sin phi1, cos phi2, sin phi2, tau
    T* tgt = tab + (size_t)p.L * 5;                     // [32]
    T* scratch = tgt + 32;                              // [warps]
    T* acc = scratch + kSu4Warps;                       // [warps][L][3]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int split = blockIdx.x % p.splits, b = blockIdx.x / p.splits;
    const int L = p.L;
    {
        const T* pb = p.pulses + (size_t)b * L * 3;
        for (int i = tid; i < L; i += kSu4Threads) {
            double s1, c1, s2, c2;
            ::sincos((double)pb[3 * i], &s1, &c1);
            ::sincos((double)pb[3 * i + 1], &s2, &c2);
            tab[5 * i] = (T)c1; tab[5 * i + 1] = (T)s1; tab[5 * i + 2] = (T)c2; tab[5 * i + 3] = (T)s2;
            tab[5 * i + 4] = pb[3 * i + 2];
        }
        if (tid < 32) tgt[tid] = p.target[(size_t)b * 32 + tid];
        if (BWD)
            for (int i = tid; i < kSu4Warps * L * 3; i += kSu4Threads) acc[i] = (T)0;
    }
    __syncthreads();
    const size_t Bm = (size_t)p.B * p.M;
    T fsum = (T)0;
    for (int tile = split; tile < p.n_tiles; tile += p.splits) {
        const long long j = (long long)tile * kSu4Threads + tid;
        const bool valid = j < p.M;
        const size_t sidx = (size_t)b * p.M + (size_t)(valid ? j : 0);
        Su4Sample<T> k;
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
            k.d1 = d1; k.d2 = d2; k.te = (T)1 + eps;
            k.nrm = (T)(0.5 * (::sqrt(1.0 + (double)d1 * d1) + ::sqrt(1.0 + (double)d2 * d2) + fabs((double)p.J)) * fabs(1.0 + (double)eps));
        }
        // ---------------- forward ----------------
        M4<T> P, E, Tm;
        m4_identity(P);
        for (int i = 0; i < L; ++i) {
            su4_pulse_exp<T>(E, tab[5 * i], tab[5 * i + 1], tab[5 * i + 2], tab[5 * i + 3], tab[5 * i + 4], k, p.J);
            m4_mul(Tm, E, P);
            P = Tm;
        }
        T trr = (T)0, tri = (T)0;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            trr += P.re[e] * tgt[2 * e] + P.im[e] * tgt[2 * e + 1];
            tri += P.re[e] * tgt[2 * e + 1] - P.im[e] * tgt[2 * e];
        }
        const T F = (trr * trr + tri * tri + (T)4) * (T)0.05;
        if (valid) {
            fsum += F;
            if (p.F_out != nullptr) p.F_out[sidx] = F;
            if (!BWD && p.U_out != nullptr) {
                T* U = p.U_out + sidx * 32;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    U[2 * e] = P.re[e];
                    U[2 * e + 1] = P.im[e];
                }
            }
        }
        if constexpr (BWD) {
            // B_L = P_L W_L^dagger with W_L = (1/10) conj(tr) T  =>  B_L = (tr/10) P_L T^dagger
            T wgt = (T)0;
            if (valid) wgt = p.weight != nullptr ? p.weight[sidx] : (T)1;
            M4<T> Bk;
            {
                M4<T> Tg;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    Tg.re[e] = tgt[2 * e];
                    Tg.im[e] = tgt[2 * e + 1];
                }
                m4_mul_hB(Tm, P, Tg);
                const T fr = wgt * trr * (T)0.1, fi = wgt * tri * (T)0.1;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    Bk.re[e] = fr * Tm.re[e] - fi * Tm.im[e];
                    Bk.im[e] = fr * Tm.im[e] + fi * Tm.re[e];
                }
            }
            for (int i = L - 1; i >= 0; --i) {
                const T c1 = tab[5 * i], s1 = tab[5 * i + 1], c2 = tab[5 * i + 2], s2 = tab[5 * i + 3], tau = tab[5 * i + 4];
                su4_pulse_exp<T>(E, c1, s1, c2, s2, tau, k, p.J);
                // dF/dtau = (1+eps) Im Tr(B H),  2H = diag(...) + phase blocks
                const T hd0 = k.d1 + k.d2 + p.J, hd1 = k.d1 - k.d2 - p.J, hd2 = -k.d1 + k.d2 - p.J, hd3 = -k.d1 - k.d2 + p.J;
                T imtr = Bk.im[0] * hd0 + Bk.im[5] * hd1 + Bk.im[10] * hd2 + Bk.im[15] * hd3;
                // off-diagonal: Tr(B H) += B[i][j] H[j][i];  H[1][0] = e^{i phi2}, H[0][1] = e^{-i phi2}, ...
                // Im(B01 e^{+i p}) = B01.im c + B01.re s ;  Im(B10 e^{-i p}) = B10.im c - B10.re s
                imtr += (Bk.im[1] + Bk.im[11]) * c2 + (Bk.re[1] + Bk.re[11]) * s2 + (Bk.im[4] + Bk.im[14]) * c2 - (Bk.re[4] + Bk.re[14]) * s2;
                imtr += (Bk.im[2] + Bk.im[7]) * c1 + (Bk.re[2] + Bk.re[7]) * s1 + (Bk.im[8] + Bk.im[13]) * c1 - (Bk.re[8] + Bk.re[13]) * s1;
                T g_tau = (T)0.5 * k.te * imtr;
                const T z1a = Bk.im[0] + Bk.im[5] - Bk.im[10] - Bk.im[15];     // Im Tr(B ZI)
                const T z2a = Bk.im[0] - Bk.im[5] + Bk.im[10] - Bk.im[15];     // Im Tr(B IZ)
                m4_mul_hA(Tm, E, Bk);                                           // U^dagger B
                m4_mul(Bk, Tm, E);                                              // ... U
                const T z1b = Bk.im[0] + Bk.im[5] - Bk.im[10] - Bk.im[15];
                const T z2b = Bk.im[0] - Bk.im[5] + Bk.im[10] - Bk.im[15];
                T g_p1 = (T)0.5 * (z1a - z1b), g_p2 = (T)0.5 * (z2a - z2b);
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    g_p1 += __shfl_xor_sync(0xffffffffu, g_p1, d);
                    g_p2 += __shfl_xor_sync(0xffffffffu, g_p2, d);
                    g_tau += __shfl_xor_sync(0xffffffffu, g_tau, d);
                }
                if (lane == 0) {
                    T* dst = acc + ((size_t)warp * L + i) * 3;
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
        for (int w = 0; w < kSu4Warps; ++w) tot += scratch[w];
        p.Fsum_part[(size_t)split * p.B + b] = tot;
    }
    if constexpr (BWD) {
        T* gout = p.G_part + ((size_t)split * p.B + b) * L * 3;
        for (int i = tid; i < 3 * L; i += kSu4Threads) {
            T tot = (T)0;
#pragma unroll
            for (int w = 0; w < kSu4Warps; ++w) tot += acc[(size_t)w * L * 3 + i];
            gout[i] = tot;
        }
    }
}

struct Su4Plan {
    int n_tiles, splits;
    size_t smem;
    int wps;       // 4: pulse train split over the four warps of a 128-thread block (su4e_split_kernel, few samples)
};
// resident blocks per SM of the kernel this plan launches (occupancy API; registers decide: 6 for the FP32
// eigenframe fwd+bwd kernel).  The grid is sized to ONE wave of resident blocks with equal tile counts.
// memoised per (device, kernel, shared-memory size): the occupancy query costs several microseconds per call
template <typename K>
static int su4_blocks_per_sm(K kern, size_t smem, int threads = 64) {
    static std::mutex mu;
    static std::unordered_map<unsigned long long, int> memo;
    const unsigned long long key = ((unsigned long long)(uintptr_t)kern * 1000003ull) ^ ((unsigned long long)smem << 8) ^
                                   (unsigned long long)current_device();
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = memo.find(key);
        if (it != memo.end()) return it->second;
    }
    int n = 0;
    if (smem > 48 * 1024) (void)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem) != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        n = 1;
    }
    std::lock_guard<std::mutex> lk(mu);
    memo[key] = n;
    return n;
}

// The split count is taken from the BACKWARD kernel's occupancy for forward launches too: one workspace size
// (uqoc_su4_workspace_bytes) then fits every launch of a shape, whichever kernel it is.
static Su4Plan su4_plan(int64_t B, int64_t L, int64_t M, int dtype, unsigned flags, bool bwd) {
    const int sms = cached_sm_count();
    Su4Plan pl;
    const bool pade = (flags & UQOC_FLAG_SU4_PADE) != 0, f64 = dtype == UQOC_F64;
    if (pade) pl.smem = f64 ? su4_smem_bytes<double>((int)L, bwd) : su4_smem_bytes<float>((int)L, bwd);
    else pl.smem = f64 ? su4e_smem_bytes<double>((int)L, bwd) : su4e_smem_bytes<float>((int)L, bwd);
    int occ;
    {
        const size_t smem_b = pade ? (f64 ? su4_smem_bytes<double>((int)L, true) : su4_smem_bytes<float>((int)L, true))
                                   : (f64 ? su4e_smem_bytes<double>((int)L, true) : su4e_smem_bytes<float>((int)L, true));
        if (pade) occ = f64 ? su4_blocks_per_sm(su4_kernel<double, true>, smem_b) : su4_blocks_per_sm(su4_kernel<float, true>, smem_b);
        else occ = f64 ? su4_blocks_per_sm(su4e_kernel<double, true, false>, smem_b) : su4_blocks_per_sm(su4e_kernel<float, true, false>, smem_b);
    }
    pl.n_tiles = (int)((M + kSu4Threads - 1) / kSu4Threads);
    pl.wps = 1;
    // few samples (fewer one-sample-per-thread warps than ~3/4 of the SM sub-partitions): split the pulse train over the
    // four warps of 128-thread blocks, 32 samples per block (fused fwd+bwd, eigenframe kernel only).  Every warp repeats
    // the per-sample eigen-decomposition and pays four 4x4 products for its prefix / seed, ~25 % extra work, so at
    // BASELINE config 4 (1024 warps for 592 sub-partitions) the one-sample-per-thread kernel stays faster (measured
    // 0.170 vs 0.221 ms) and keeps the default there.
    if (bwd && !pade && !(flags & UQOC_FLAG_WPS1) && L >= 32 &&
        ((flags & UQOC_FLAG_WPS4) || B * (int64_t)pl.n_tiles * 2 * 4 <= (int64_t)sms * 4 * 3)) {
        pl.wps = 4;
        pl.smem = f64 ? su4s_smem_bytes<double>((int)L) : su4s_smem_bytes<float>((int)L);
        occ = f64 ? su4_blocks_per_sm(su4e_split_kernel<double, true>, pl.smem, kSu4sThreads)
                  : su4_blocks_per_sm(su4e_split_kernel<float, true>, pl.smem, kSu4sThreads);
        pl.n_tiles = (int)((M + 31) / 32);
    }
    // one wave: at most sms*occ blocks; every block of a target walks `rounds` tiles (the last may walk one less)
    int64_t splits = ((int64_t)sms * occ) / B;
    if (splits < 1) splits = 1;
    if (splits > pl.n_tiles) splits = pl.n_tiles;
    const int64_t rounds = (pl.n_tiles + splits - 1) / splits;
    splits = (pl.n_tiles + rounds - 1) / rounds;
    if (splits > 4095) splits = 4095;
    const int fsp = (flags >> 18) & 0xFFF;
    if (fsp) splits = fsp < pl.n_tiles ? fsp : pl.n_tiles;
    pl.splits = (int)splits;
    return pl;
}

template <typename T, bool BWD>
static int su4_run(const void* pulses, const void* target, const void* err, const void* weight, int64_t B, int64_t L,
                   int64_t M, int64_t j0, double J, double sig_d, double sig_e, uint64_t seed, uint64_t offset, void* U_out,
                   void* F_out, void* err_out, void* Fsum, void* G, void* workspace, int64_t workspace_bytes, int dtype,
                   unsigned flags, cudaStream_t stream, const void* cot = nullptr) {
    const Su4Plan pl = su4_plan(B, L, M, dtype, flags, BWD);
    UQOC_CHECK_ARG(B * (int64_t)pl.splits <= 0x7fffffffLL, "grid too large: %lld blocks", (long long)(B * pl.splits));
    UQOC_CHECK_ARG((flags & UQOC_FLAG_RNG_FROM_DEVICE) || offset <= 0xffffffffULL,
                   "Philox offset must be < 2^32 (it is one 32-bit counter word), got %llu", (unsigned long long)offset);
    Su4Params<T> p;
    p.pulses = (const T*)pulses; p.target = (const T*)target; p.err = (const T*)err; p.weight = (const T*)weight;
    p.B = (int)B; p.L = (int)L; p.M = (int)M; p.n_tiles = pl.n_tiles; p.splits = pl.splits;
    p.j0 = j0; p.J = (T)J; p.sig_d = (T)sig_d; p.sig_e = (T)sig_e; p.seed = seed; p.offset = (unsigned)offset;
    p.U_out = (T*)U_out; p.F_out = (T*)F_out; p.err_out = (T*)err_out;
    p.rng_dev = (flags & UQOC_FLAG_RNG_FROM_DEVICE) ? (const unsigned long long*)(uintptr_t)seed : nullptr;
    p.cot = (const T*)cot;
    const int64_t n_g = BWD ? B * L * 3 : 0;
    if (pl.splits > 1) {
        const int64_t need = (int64_t)pl.splits * (B + n_g) * (int64_t)sizeof(T);
        if (workspace == nullptr || workspace_bytes < need) {
            set_error("workspace too small: need %lld bytes, got %lld", (long long)need, (long long)workspace_bytes);
            return UQOC_E_WORKSPACE;
        }
        p.Fsum_part = (T*)workspace;
        p.G_part = (T*)workspace + (size_t)pl.splits * B;
    } else {
        p.Fsum_part = (T*)Fsum;
        p.G_part = (T*)G;
    }
    static_assert(kSu4Threads == kSu4eThreads, "both SU(4) kernels share the launch plan");
    auto kern = (flags & UQOC_FLAG_SU4_PADE) ? su4_kernel<T, BWD> : su4e_kernel<T, BWD, false>;
    if constexpr (BWD) {
        if (cot != nullptr) kern = su4e_kernel<T, true, true>;      // cotangent-seeded adjoint: eigenframe kernel only
    }
    if (pl.smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) {
            set_error("su4 kernel needs %zu bytes of shared memory (L too large): %s", pl.smem, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return UQOC_E_UNSUPPORTED;
        }
    }
    int threads = kSu4Threads;
    if constexpr (BWD) {
        if (pl.wps == 4 && cot == nullptr) {
            kern = su4e_split_kernel<T, true>;
            threads = kSu4sThreads;
            if (pl.smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
                if (e != cudaSuccess) {
                    set_error("su4 kernel needs %zu bytes of shared memory (L too large): %s", pl.smem, cudaGetErrorString(e));
                    (void)cudaGetLastError();
                    return UQOC_E_UNSUPPORTED;
                }
            }
        }
    }
    kern<<<(unsigned)(B * pl.splits), threads, pl.smem, stream>>>(p);
    int rc = launch_status("su4_kernel");
    if (rc) return rc;
    if (pl.splits > 1 && (Fsum != nullptr || n_g > 0)) {
        launch_reduce_partials<T>(p.Fsum_part, p.G_part, pl.splits, (int)B, n_g, (T*)Fsum, (T*)G, stream);
        return launch_status("su4 reduce_partials");
    }
    return 0;
}

}  // namespace uqoc

using namespace uqoc;

extern "C" {

int64_t uqoc_su4_workspace_bytes(int64_t B, int64_t L, int64_t M, int dtype, unsigned flags) {
    if (B < 1 || L < 1 || M < 1) return 0;
    const Su4Plan pb = su4_plan(B, L, M, dtype, flags, true), pf = su4_plan(B, L, M, dtype, flags, false);
    const int splits = pb.splits > pf.splits ? pb.splits : pf.splits;     // the train-split plan of fwd+bwd launches differs
    if (splits <= 1) return 0;
    return (int64_t)splits * (B + B * L * 3) * (dtype == UQOC_F64 ? 8 : 4);
}

static int su4_check(int64_t B, int64_t L, int64_t M, int dtype) {
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "dtype must be UQOC_F32 or UQOC_F64, got %d", dtype);
    UQOC_CHECK_ARG(B >= 1 && B <= (1 << 24), "B out of range: %lld", (long long)B);
    UQOC_CHECK_ARG(L >= 1 && L <= (1 << 16), "L out of range: %lld", (long long)L);
    UQOC_CHECK_ARG(M >= 1 && M <= (1LL << 31) - 1, "M out of range: %lld", (long long)M);
    return 0;
}

int uqoc_su4_fwdbwd(const void* pulses, const void* target, const void* err, const void* weight, int64_t B, int64_t L,
                    int64_t M, int64_t j0, double J, double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                    void* F_out, void* err_out, void* Fsum, void* G, void* workspace, int64_t workspace_bytes, int dtype,
                    unsigned flags, void* stream) {
    int rc = su4_check(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && target && Fsum && G, "pulses, target, Fsum and G must be non-null");
    if (dtype == UQOC_F64)
        return su4_run<double, true>(pulses, target, err, weight, B, L, M, j0, J, sig_d, sig_e, seed, offset, nullptr, F_out,
                                     err_out, Fsum, G, workspace, workspace_bytes, dtype, flags, (cudaStream_t)stream);
    return su4_run<float, true>(pulses, target, err, weight, B, L, M, j0, J, sig_d, sig_e, seed, offset, nullptr, F_out, err_out,
                                Fsum, G, workspace, workspace_bytes, dtype, flags, (cudaStream_t)stream);
}

int uqoc_su4_forward(const void* pulses, const void* target, const void* err, int64_t B, int64_t L, int64_t M, int64_t j0,
                     double J, double sig_d, double sig_e, uint64_t seed, uint64_t offset, void* U_out, void* F_out,
                     void* err_out, void* Fsum, void* workspace, int64_t workspace_bytes, int dtype, unsigned flags,
                     void* stream) {
    int rc = su4_check(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && target, "pulses and target must be non-null");
    UQOC_CHECK_ARG(U_out || F_out || Fsum || err_out, "no output requested");
    if (dtype == UQOC_F64)
        return su4_run<double, false>(pulses, target, err, nullptr, B, L, M, j0, J, sig_d, sig_e, seed, offset, U_out, F_out,
                                      err_out, Fsum, nullptr, workspace, workspace_bytes, dtype, flags, (cudaStream_t)stream);
    return su4_run<float, false>(pulses, target, err, nullptr, B, L, M, j0, J, sig_d, sig_e, seed, offset, U_out, F_out, err_out,
                                 Fsum, nullptr, workspace, workspace_bytes, dtype, flags, (cudaStream_t)stream);
}

/* backward of the strict generator signature: pulses (Bm, L, 3) one row PER SAMPLE, err (3, Bm), gU = dLoss/dU
 * (Bm, 4, 4, 2) in torch's convention for complex tensors (dLoss = Re sum conj(gU) dU) -> g_pulses (Bm, L, 3).
 * One block per sample row (correct, not tuned: the fused entry points are the fast path). */
int uqoc_su4_generator_backward(const void* pulses, const void* err, const void* gU, int64_t Bm, int64_t L, double J,
                                void* g_pulses, int dtype, unsigned flags, void* stream) {
    int rc = su4_check(Bm, L, 1, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && err && gU && g_pulses, "pulses, err, gU and g_pulses must be non-null");
    flags &= ~(unsigned)(UQOC_FLAG_SU4_PADE | UQOC_FLAG_RNG_FROM_DEVICE);
    if (dtype == UQOC_F64)
        return su4_run<double, true>(pulses, gU /* unused target */, err, nullptr, Bm, L, 1, 0, J, 0, 0, 0, 0, nullptr, nullptr,
                                     nullptr, nullptr, g_pulses, nullptr, 0, dtype, flags, (cudaStream_t)stream, gU);
    return su4_run<float, true>(pulses, gU, err, nullptr, Bm, L, 1, 0, J, 0, 0, 0, 0, nullptr, nullptr, nullptr, nullptr,
                                g_pulses, nullptr, 0, dtype, flags, (cudaStream_t)stream, gU);
}

}  // extern "C"
