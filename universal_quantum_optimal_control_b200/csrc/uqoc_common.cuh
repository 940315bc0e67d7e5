// Common device helpers for libuqoc (sm_100a): scalar traits, quaternion algebra,
// sin/cos policies, Philox4x32-10 + Box-Muller, per-sample constants.
//
// Quaternion <-> SU(2):  U = q0 I - i (q1 X + q2 Y + q3 Z); matrix product == Hamilton
// product.  One pulse of SCORE.py:117-127 is q = (cos h, r sin h (cos phi, sin phi, delta))
// with w = sqrt(1+delta^2), r = 1/w, h = tau * (1+eps) w / 2.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/uqoc.h"

namespace uqoc {

// ---------------------------------------------------------------- error handling
void set_error(const char* fmt, ...);

#define UQOC_CHECK_ARG(cond, ...)      \
    do {                               \
        if (!(cond)) {                 \
            uqoc::set_error(__VA_ARGS__); \
            return UQOC_E_BADARG;      \
        }                              \
    } while (0)

// ---- per-device caches of launch-time queries (a training loop makes the same call every 50-200 us: the CUDA runtime
// queries below cost microseconds each) ------------------------------------------------------------------------------
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        dev = 0;
    }
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int cached_sm_count() {
    static std::atomic<int> n_sm[kMaxDevices];
    const int dev = current_device();
    int n = n_sm[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            (void)cudaGetLastError();
            return 148;  // B200
        }
        n_sm[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is sticky per (device, kernel): raise it only when a launch needs more
// than any earlier one.  `slot` = a static array owned by the caller, one per kernel instantiation.
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kern, size_t smem, std::atomic<int>* slot) {
    if (smem <= 48 * 1024) return cudaSuccess;
    const int dev = current_device();
    if ((int)smem <= slot[dev].load(std::memory_order_relaxed)) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) slot[dev].store((int)smem, std::memory_order_relaxed);
    return e;
}

// ---- debug build (-DUQOC_DEBUG_CHECKS, tools/debug_checks.sh): device-side bounds asserts on every shared-memory
// layout and staging / accumulator index, and epoch tags on the chunk-product exchange between the warps of a block
// (a stale read = a missing or misplaced barrier traps instead of returning a plausible number).  The pool this code is
// developed on refuses compute-sanitizer (memcheck / racecheck), so these checks plus the oracle comparisons at small
// sizes are the race / bounds evidence; they compile to nothing in the shipped library.
#ifdef UQOC_DEBUG_CHECKS
#include <cassert>
#define UQOC_ASSERT(c) assert(c)
#else
#define UQOC_ASSERT(c) ((void)0)
#endif
__device__ __forceinline__ unsigned dyn_smem_bytes() {
    unsigned v;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(v));
    return v;
}

inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// ---------------------------------------------------------------- sincos policies
enum SinCosPolicy { SC_POLY = 0, SC_MUFU = 1, SC_LIBM = 2, SC_TABLE = 3 };

// Accurate FP32 sin/cos of h reduced modulo PI (not 2 PI): returns (s, c) = (-1)^k (sin h,
// cos h) and k in the low bits of `kbits`.  The common sign is irrelevant to the fidelity and
// to every gradient (U -> -U), so the fused kernels ignore it; the forward kernel that emits
// U_out tracks the parity.  13 FMA-pipe instructions, no select / branch.
__device__ __forceinline__ void sincos_modpi(float h, float& s, float& c, int& kbits) {
    const float MAGIC = 12582912.0f;  // 1.5 * 2^23: round-to-nearest-integer trick
    float kf = fmaf(h, 0.318309886183790672f, MAGIC);
    kbits = __float_as_int(kf);
    kf -= MAGIC;
    float r = fmaf(kf, -3.14159274101257324f, h);   // pi_hi = float(pi)
    r = fmaf(kf, 8.74227765734758577e-08f, r);      // -(pi - pi_hi)
    const float z = r * r;
    // sin r = r + r z (S0 + S1 z + S2 z^2 + S3 z^3),  |r| <= pi/2  (near-minimax fit)
    float ps = fmaf(z, 2.6325158160034334e-06f, -1.9822049944195896e-04f);
    ps = fmaf(z, ps, 8.3332369104027748e-03f);
    ps = fmaf(z, ps, -1.6666665673255920e-01f);
    const float rz = r * z;
    s = fmaf(rz, ps, r);
    // cos r = 1 + z (C0 + C1 z + ... + C4 z^4)
    float pc = fmaf(z, -2.6282461362825416e-07f, 2.4774040866759606e-05f);
    pc = fmaf(z, pc, -1.3888647081330419e-03f);
    pc = fmaf(z, pc, 4.1666660457849503e-02f);
    pc = fmaf(z, pc, -0.5f);
    c = fmaf(z, pc, 1.0f);
}

template <typename T, int SC>
struct SinCos;
// every policy exposes eval(h, s, c, kbits, tab): `tab` is the shared-memory table {sin[N] | cos[N]} of the
// policies with kTableN > 0 and is ignored by the others.

template <>
struct SinCos<float, SC_POLY> {
    static constexpr bool kTracksParity = true;
    static constexpr int kTableN = 0;
    __device__ __forceinline__ static void eval(float h, float& s, float& c, int& kbits, const float* = nullptr) {
        sincos_modpi(h, s, c, kbits);
    }
};
template <>
struct SinCos<float, SC_MUFU> {
    static constexpr bool kTracksParity = false;
    static constexpr int kTableN = 0;
    __device__ __forceinline__ static void eval(float h, float& s, float& c, int& kbits, const float* = nullptr) {
        kbits = 0;
        __sincosf(h, &s, &c);
    }
};
template <>
struct SinCos<float, SC_LIBM> {
    static constexpr bool kTracksParity = false;
    static constexpr int kTableN = 0;
    __device__ __forceinline__ static void eval(float h, float& s, float& c, int& kbits, const float* = nullptr) {
        kbits = 0;
        sincosf(h, &s, &c);
    }
};
template <>
struct SinCos<double, SC_LIBM> {
    static constexpr bool kTracksParity = false;
    static constexpr int kTableN = 0;
    __device__ __forceinline__ static void eval(double h, double& s, double& c, int& kbits, const double* = nullptr) {
        kbits = 0;
        sincos(h, &s, &c);
    }
};
// FP64 table policy: h = k pi/1024 + r (three-term Cody-Waite, exact for |k| < 2^20), (sin, cos)(k pi/1024) from a
// double table in shared memory, residual rotation to full double accuracy (sin r to r^5, cos r to r^4:
// truncation 4e-24 / 2e-20 at |r| <= pi/2048).  15 DFMA-pipe instructions instead of libm's ~45; like the
// FP32 paths the sign (-1)^(k div 1024) is dropped (kbits = k >> 10 lets the U_out kernel track it).
template <>
struct SinCos<double, SC_TABLE> {
    static constexpr bool kTracksParity = true;
    static constexpr int kTableN = 1024;
    __device__ __forceinline__ static void eval(double h, double& s, double& c, int& kbits, const double* tab) {
        const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
        double kf = ::fma(h, 325.94932345220167, MAGIC);
        const int k = __double2loint(kf);
        kbits = k >> 10;
        kf -= MAGIC;
        double r = ::fma(kf, -0.003067961575652589, h);
        r = ::fma(kf, -1.1869336926769907e-13, r);
        r = ::fma(kf, -6.878046611685938e-30, r);
        const double st = tab[k & 1023], ct = tab[1024 + (k & 1023)];
        const double z = r * r;
        const double sr = ::fma(r * z, ::fma(z, 8.3333333333333332e-03, -1.6666666666666666e-01), r);
        const double cr = ::fma(z, ::fma(z, 4.1666666666666664e-02, -0.5), 1.0);
        s = ::fma(ct, sr, st * cr);
        c = ::fma(-st, sr, ct * cr);
    }
    // Same table, angle given as tau * (index slope ap = a 1024/pi): k = round(tau ap) from the magic-number FMA,
    // the residual frac = fma(tau, ap, -k) with ONE rounding (the Cody-Waite triple is not needed because the angle
    // is never formed), sin r to r^3 (truncation 7e-17 at |r| <= pi/2048).  13 DFMA-pipe instructions including
    // the product with tau (the eval() path needs 16).
    __device__ __forceinline__ static void eval_slope(double tau, double ap, double& s, double& c, int& kbits, const double* tab) {
        const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
        const double kf = ::fma(tau, ap, MAGIC);
        const int k = __double2loint(kf);
        kbits = k >> 10;
        const double frac = ::fma(tau, ap, MAGIC - kf);
        const double r = frac * 0.0030679615757712823;          // pi/1024
        const double st = tab[k & 1023], ct = tab[1024 + (k & 1023)];
        const double z = r * r;
        const double sr = ::fma(r * z, -1.6666666666666666e-01, r);
        const double cr = ::fma(z, ::fma(z, 4.1666666666666664e-02, -0.5), 1.0);
        s = ::fma(ct, sr, st * cr);
        c = ::fma(-st, sr, ct * cr);
    }
    // full-sign variant (the backward sweep looks up the DOUBLE angle): sign (-1)^(k div 1024) restored by an
    // integer XOR on the high words (ALU pipe, not FP64)
    __device__ __forceinline__ static void eval_slope_signed(double tau, double ap, double& s, double& c, const double* tab) {
        int kb;
        eval_slope(tau, ap, s, c, kb, tab);
        const int sg = kb << 31;
        s = __hiloint2double(__double2hiint(s) ^ sg, __double2loint(s));
        c = __hiloint2double(__double2hiint(c) ^ sg, __double2loint(c));
    }
};

// ---------------------------------------------------------------- small math traits
template <typename T>
struct Real;
template <>
struct Real<float> {
    __device__ __forceinline__ static float fma(float a, float b, float c) { return fmaf(a, b, c); }
    __device__ __forceinline__ static float rsqrt_acc(float x) { return 1.0f / sqrtf(x); }
    __device__ __forceinline__ static float sqrt(float x) { return sqrtf(x); }
};
template <>
struct Real<double> {
    __device__ __forceinline__ static double fma(double a, double b, double c) { return ::fma(a, b, c); }
    __device__ __forceinline__ static double rsqrt_acc(double x) { return 1.0 / ::sqrt(x); }
    __device__ __forceinline__ static double sqrt(double x) { return ::sqrt(x); }
};

// 4-wide row type for the shared-memory pulse tables (one LDS.128 / two for double)
template <typename T>
struct alignas(4 * sizeof(T)) Row4 {
    T x, y, z, w;
};

// ---------------------------------------------------------------- quaternions
template <typename T>
struct Quat {
    T a, b, c, d;  // scalar, x, y, z
};

template <typename T>
__device__ __forceinline__ Quat<T> qmul(const Quat<T>& p, const Quat<T>& q) {
    Quat<T> r;
    r.a = p.a * q.a - p.b * q.b - p.c * q.c - p.d * q.d;
    r.b = p.a * q.b + p.b * q.a + p.c * q.d - p.d * q.c;
    r.c = p.a * q.c - p.b * q.d + p.c * q.a + p.d * q.b;
    r.d = p.a * q.d + p.b * q.c - p.c * q.b + p.d * q.a;
    return r;
}
template <typename T>
__device__ __forceinline__ Quat<T> qconj(const Quat<T>& p) {
    return Quat<T>{p.a, -p.b, -p.c, -p.d};
}

// ---------------------------------------------------------------- Philox4x32-10
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// (delta, eps) of sample (b, j): counter = (j lo, j hi, b, offset), key = seed.
// u_k = (x_k + 0.5) 2^-32, Box-Muller pair scaled by (sig_d, sig_e).  Oracle twin:
// oracle/uqoc_oracle.py::philox_errors.
template <typename T>
__device__ __forceinline__ void philox_delta_eps(uint64_t j, uint32_t b, uint64_t seed, uint32_t offset,
                                                 T sig_d, T sig_e, T& delta, T& eps);

template <>
__device__ __forceinline__ void philox_delta_eps<float>(uint64_t j, uint32_t b, uint64_t seed, uint32_t offset,
                                                        float sig_d, float sig_e, float& delta, float& eps) {
    const Philox4 x = philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), b, offset, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u0 = fmaf((float)x.x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float u1 = fmaf((float)x.y, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float rad = sqrtf(-2.0f * logf(u0));
    float sn, cs;
    sincospif(2.0f * u1, &sn, &cs);
    delta = sig_d * rad * cs;
    eps = sig_e * rad * sn;
}
template <>
__device__ __forceinline__ void philox_delta_eps<double>(uint64_t j, uint32_t b, uint64_t seed, uint32_t offset,
                                                         double sig_d, double sig_e, double& delta, double& eps) {
    const Philox4 x = philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), b, offset, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u0 = ((double)x.x + 0.5) * 2.3283064365386963e-10;
    const double u1 = ((double)x.y + 0.5) * 2.3283064365386963e-10;
    const double rad = ::sqrt(-2.0 * ::log(u0));
    double sn, cs;
    sincospi(2.0 * u1, &sn, &cs);
    delta = sig_d * rad * cs;
    eps = sig_e * rad * sn;
}

// ---------------------------------------------------------------- per-sample constants
// Computed in double and rounded once, so that FP32 mode carries a single rounding (not a
// chain of them) into the systematic part of every rotation angle of the sample.
template <typename T>
struct SampleConst {
    T a;      // (1+eps) w / 2  : half-angle per unit tau
    T a2;     // (1+eps) w      : full rotation angle per unit tau
    T r;      // 1/w
    T r2;     // 1/w^2
    T rd;     // delta/w
    T delta;  // detuning
    T ae;     // (1+eps)/2 = a*r : d h / d tau projected on the axis
    T ap;     // a  * 1024/pi : table index of the half angle per unit tau (packed table kernel)
    T a2p;    // a2 * 1024/pi : table index of the full angle per unit tau
};
template <typename T>
__device__ __forceinline__ SampleConst<T> make_sample_const(T delta, T eps) {
    // double arithmetic, one rounding each; rsqrt + multiplies only (no double sqrt / divide)
    const double d = (double)delta, e = (double)eps;
    const double x = ::fma(d, d, 1.0);
    const double r = ::rsqrt(x);          // 1/w, correctly rounded to ~1 ulp in double
    const double w = x * r;
    SampleConst<T> k;
    k.a = (T)(0.5 * (1.0 + e) * w);
    k.a2 = (T)((1.0 + e) * w);
    k.r = (T)r;
    k.r2 = (T)(r * r);
    k.rd = (T)(d * r);
    k.delta = delta;
    k.ae = (T)(0.5 * (1.0 + e));
    k.ap = (T)(0.5 * (1.0 + e) * w * 325.94932345220167);
    k.a2p = (T)((1.0 + e) * w * 325.94932345220167);
    return k;
}

}  // namespace uqoc
