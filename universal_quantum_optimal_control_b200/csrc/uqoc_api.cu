// C ABI of libuqoc.so (see include/uqoc.h): argument checking, launch planning, the small
// epilogue kernels (partials reduction, loss finalize), the Philox sampler and the FP32 peak probe.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "uqoc_su2_kernels.cuh"
#include "uqoc_su2_x2.cuh"

extern "C" int uqoc_loss_finalize(const void* Fsum, int64_t B, double n_total, int loss_kind, double tau, double k, void* G,
                                  int64_t G_numel, void* loss_out, int dtype, void* stream);

namespace uqoc {

// ------------------------------------------------------------------ error string
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

template <typename T, int SC>
int su2_launch(const Su2Params<T>& p, const Su2Plan& plan, bool bwd, cudaStream_t stream);
// defined in uqoc_su2_f32.cu / uqoc_su2_f32_fast.cu / uqoc_su2_f64.cu
template <> int su2_launch<float, SC_POLY>(const Su2Params<float>&, const Su2Plan&, bool, cudaStream_t);
template <> int su2_launch<float, SC_MUFU>(const Su2Params<float>&, const Su2Plan&, bool, cudaStream_t);
template <> int su2_launch<double, SC_LIBM>(const Su2Params<double>&, const Su2Plan&, bool, cudaStream_t);
template <> int su2_launch<double, SC_TABLE>(const Su2Params<double>&, const Su2Plan&, bool, cudaStream_t);

// su2_launch<float, SC_LIBM> / <double, SC_POLY|SC_MUFU> are never instantiated: route by type.
template <>
int su2_launch<float, SC_LIBM>(const Su2Params<float>&, const Su2Plan&, bool, cudaStream_t) {
    set_error("internal: float/libm path not built");
    return UQOC_E_UNSUPPORTED;
}
template <>
int su2_launch<float, SC_TABLE>(const Su2Params<float>&, const Su2Plan&, bool, cudaStream_t) {
    set_error("internal: float/table scalar path not built");
    return UQOC_E_UNSUPPORTED;
}
template <>
int su2_launch<double, SC_POLY>(const Su2Params<double>&, const Su2Plan&, bool, cudaStream_t) {
    set_error("internal: double/poly path not built");
    return UQOC_E_UNSUPPORTED;
}
template <>
int su2_launch<double, SC_MUFU>(const Su2Params<double>&, const Su2Plan&, bool, cudaStream_t) {
    set_error("internal: double/mufu path not built");
    return UQOC_E_UNSUPPORTED;
}

// ------------------------------------------------------------------ launch planning
static int sm_count() { return cached_sm_count(); }

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Choose samples/thread (ST), lanes/sample (LPS) and sample-tile splits per target.
static Su2Plan make_plan(int64_t B, int64_t L, int64_t M, int dtype, unsigned flags, bool bwd) {
    const int sms = sm_count();
    const int64_t N = B * M;                         // samples
    const int64_t one = (int64_t)sms * 4 * 32;       // threads for one warp per SM sub-partition
    int st = 1, lps = 1;
    const int st_max = (dtype == UQOC_F64) ? 2 : 4;
    // more samples per thread (packed f32x2 pairs, amortised staging/reduction) once every
    // sub-partition still gets a warp; below that, split the pulse train over lanes instead
    if (N >= 4 * one && st_max >= 4) st = 4;
    else if (N >= 2 * one) st = 2;
    if (N * 2 <= one) {
        while (lps < 32 && lps * 2 <= L && N * lps * 2 <= 2 * one) lps *= 2;
    }
    const int fst = (flags >> 8) & 0xF, flps = (flags >> 12) & 0x3F;
    if (fst) st = fst;
    if (flps) lps = flps;
    if (lps > 1) st = 1;
    Su2Plan plan;
    // few samples, long train: the packed kernel with the pulse train split over the block's 4 warps
    // (4x the warps for the same samples) beats both ST = 1 and the lane-split scalar kernel
    int wps = 1;
    if (dtype == UQOC_F32 && !(flags & UQOC_FLAG_NO_PACKED) && lps == 1 && !fst && N * 2 > one && N < 4 * one && L >= 32) {
        st = 2;
        wps = 4;
    }
    // many samples but few targets with few 512-sample tiles each (e.g. 100 targets x 1000 samples, L = 400): one
    // block per (target, tile) leaves most SMs with one block; split the train over the 4 warps instead, keeping
    // 4 samples per thread (measured 0.116 vs 0.143 ms on that shape)
    if (dtype == UQOC_F32 && !(flags & UQOC_FLAG_NO_PACKED) && lps == 1 && !fst && wps == 1 && st == 4 && L >= 32 &&
        B * ((M + 4 * kThreads - 1) / (4 * kThreads)) < 2 * (int64_t)sms)
        wps = 4;
    if (flags & UQOC_FLAG_WPS4) wps = 4;
    if (flags & UQOC_FLAG_WPS1) wps = 1;
    plan.st = st;
    plan.lps = lps;
    plan.packed = (dtype == UQOC_F32) && lps == 1 && st >= 2 && !(flags & UQOC_FLAG_NO_PACKED);
    if (!plan.packed) wps = 1;
    plan.wps = wps;
    const int nb = (lps == 1) ? 8 : 1;
    const int chunks = plan.packed ? wps : lps;
    plan.C = round_up((int)((L + chunks - 1) / chunks), nb);
    if (plan.C < nb) plan.C = nb;
    const int ts = plan.packed ? ((wps == 1 ? kThreads : 32) * st) : (kThreads / lps) * st;
    plan.n_tiles = (int)((M + ts - 1) / ts);
    if (plan.n_tiles < 1) plan.n_tiles = 1;
    int64_t want = (int64_t)sms * (plan.wps == 4 ? 8 : 4);   // resident blocks per SM (light blocks when the train is split)
    int64_t splits = (want + B - 1) / B;
    if (splits > plan.n_tiles) splits = plan.n_tiles;
    if (splits < 1) splits = 1;
    const int fsp = (flags >> 18) & 0xFFF;
    if (fsp) splits = fsp < plan.n_tiles ? fsp : plan.n_tiles;
    plan.splits = (int)splits;
    plan.table = plan.packed && !(flags & UQOC_FLAG_FAST_SINCOS) && !(flags & UQOC_FLAG_NO_TABLE);
    // fat block (few targets, train split over warps): one 7 x 128-thread block per SM; its virtual blocks are sample-tile
    // streams of the same target that share the staged pulse train / table, and the target leaves one partial row per
    // block (148 instead of 1024 at BASELINE config 3), few enough for the in-kernel epilogue
    plan.vb = 1;
    plan.cps = plan.splits;
    if (plan.packed && plan.table && bwd && plan.wps == 4 && plan.st == 2 && !fsp && !(flags & UQOC_FLAG_NO_FAT) &&
        B * 2 <= sms && plan.n_tiles >= 2 * kX2FatVB) {
        int64_t cps = sms / B;
        const int64_t need = (plan.n_tiles + kX2FatVB - 1) / kX2FatVB;
        if (cps > need) cps = need;
        plan.vb = kX2FatVB;
        plan.cps = (int)cps;
        plan.splits = (int)cps * kX2FatVB;
    }
    const int fin_thr = bwd ? kThreads * plan.vb : 0;
    plan.fin = false;                                   // decided per call (su2_run)
    plan.smem = (dtype == UQOC_F64) ? su2_smem_bytes<double>(lps, plan.C, bwd, (flags & UQOC_FLAG_NO_TABLE) ? 0 : 1024, bwd)
                                    : (plan.packed ? su2_x2_smem_bytes(plan.C, plan.wps, plan.st, bwd, plan.table, plan.vb, fin_thr)
                                                   : su2_smem_bytes<float>(lps, plan.C, bwd, 0, bwd));
    return plan;
}

// workspace layout: [ticket area | G_part (cps x B*L*2) | Fsum_part (cps x B)]
constexpr int64_t kTicketBytes = 256;

// ------------------------------------------------------------------ small kernels
template <typename T>
__global__ void target_coeffs_kernel(const T* __restrict__ U, long long B, T* __restrict__ out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const T* t = U + b * 8;  // T00 (0,1) T01 (2,3) T10 (4,5) T11 (6,7)
    T* c = out + b * 8;
    // c0 = T00+T11, c1 = i (T01+T10), c2 = T10-T01, c3 = i (T00-T11)
    c[0] = t[0] + t[6];        c[4] = t[1] + t[7];
    c[1] = -(t[3] + t[5]);     c[5] = t[2] + t[4];
    c[2] = t[4] - t[2];        c[6] = t[5] - t[3];
    c[3] = -(t[1] - t[7]);     c[7] = t[0] - t[6];
}

template <typename T>
__global__ void philox_errors_kernel(long long B, long long M, long long j0, T sig_d, T sig_e,
                                     unsigned long long seed, unsigned offset, T* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * M) return;
    const long long b = i / M, j = i % M;
    T d, e;
    philox_delta_eps<T>((uint64_t)(j0 + j), (uint32_t)b, seed, offset, sig_d, sig_e, d, e);
    out[i] = d;
    out[B * M + i] = e;
}

__device__ __forceinline__ void loss_eval(double Fbar, int kind, double tau, double k, double& val, double& dval) {
    if (kind == UQOC_LOSS_SHARP) {
        const double z = exp(-k * (Fbar - tau));
        const double lg = log(1.0 + z);
        val = lg * (1.0 - Fbar);
        dval = -k * z / (1.0 + z) * (1.0 - Fbar) - lg;
    } else if (kind == UQOC_LOSS_NLL) {
        val = -log(Fbar);
        dval = -1.0 / Fbar;
    } else if (kind == UQOC_LOSS_INFIDELITY) {
        val = 1.0 - Fbar;
        dval = -1.0;
    } else {
        val = Fbar;
        dval = 1.0;
    }
}

// Every block recomputes Fbar with the same fixed-order reduction (bit-identical across
// blocks and ranks), then scales its slice of G by dloss/dFbar / n_total.
template <typename T>
__global__ void __launch_bounds__(256) loss_finalize_kernel(const T* __restrict__ Fsum, int B, double n_total, int kind,
                                                            double tau, double k, T* __restrict__ G, long long n_g,
                                                            T* __restrict__ loss_out) {
    __shared__ double red[256];
    __shared__ double s_scale;
    grid_dependency_wait();
    double acc = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) acc += (double)Fsum[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if (threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double Fbar = red[0] / n_total;
        double val, dval;
        loss_eval(Fbar, kind, tau, k, val, dval);
        s_scale = dval / n_total;
        if (blockIdx.x == 0 && loss_out != nullptr) {
            loss_out[0] = (T)val;
            loss_out[1] = (T)Fbar;
            loss_out[2] = (T)dval;
        }
    }
    __syncthreads();
    if (G != nullptr) {
        const T sc = (T)s_scale;
        for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_g; i += (long long)gridDim.x * 256) G[i] *= sc;
    }
}

template <typename T>
__global__ void __launch_bounds__(512) fma_probe_kernel(int iters, T x, T y, T* out) {
    T a0 = (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a0 = Real<T>::fma(a0, x, y); a1 = Real<T>::fma(a1, x, y); a2 = Real<T>::fma(a2, x, y); a3 = Real<T>::fma(a3, x, y);
            a4 = Real<T>::fma(a4, x, y); a5 = Real<T>::fma(a5, x, y); a6 = Real<T>::fma(a6, x, y); a7 = Real<T>::fma(a7, x, y);
        }
    }
    const T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == (T)123456789) out[0] = s;  // keeps the loop alive, never true in practice
}


// FFMA2 (packed f32x2, new on sm_100) probe: two FP32 FMAs per lane per instruction
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__global__ void __launch_bounds__(512) ffma2_probe_kernel(int iters, unsigned long long x, unsigned long long y, unsigned long long* out) {
    unsigned long long a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a0 = ffma2(a0, x, y); a1 = ffma2(a1, x, y); a2 = ffma2(a2, x, y); a3 = ffma2(a3, x, y);
        }
    }
    if ((a0 ^ a1 ^ a2 ^ a3) == 0x123456789abcULL) out[0] = a0;
}

// F = (|Tr(U^dagger T)|^2 + d) / (d (d+1))  for materialised (Bm, d, d) complex tensors
// (SCORE.py:168-183), and its backward  gU = gF * 2/(d(d+1)) * conj(tr) * T.
template <typename T>
__global__ void fidelity_fwd_kernel(const T* __restrict__ U, const T* __restrict__ Tg, long long Bm, int d,
                                    long long t_stride, T* __restrict__ F) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= Bm) return;
    const T* u = U + s * 2 * d * d;
    const T* t = Tg + s * t_stride;
    T re = 0, im = 0;
    for (int e = 0; e < d * d; ++e) {
        const T ur = u[2 * e], ui = u[2 * e + 1], tr_ = t[2 * e], ti = t[2 * e + 1];
        re += ur * tr_ + ui * ti;   // conj(u) * t
        im += ur * ti - ui * tr_;
    }
    F[s] = (re * re + im * im + (T)d) / (T)(d * (d + 1));
}
template <typename T>
__global__ void fidelity_bwd_kernel(const T* __restrict__ U, const T* __restrict__ Tg, const T* __restrict__ gF,
                                    long long Bm, int d, long long t_stride, T* __restrict__ gU) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= Bm) return;
    const T* u = U + s * 2 * d * d;
    const T* t = Tg + s * t_stride;
    T re = 0, im = 0;
    for (int e = 0; e < d * d; ++e) {
        const T ur = u[2 * e], ui = u[2 * e + 1], tr_ = t[2 * e], ti = t[2 * e + 1];
        re += ur * tr_ + ui * ti;
        im += ur * ti - ui * tr_;
    }
    const T sc = gF[s] * (T)2 / (T)(d * (d + 1));
    T* g = gU + s * 2 * d * d;
    for (int e = 0; e < d * d; ++e) {
        const T tr_ = t[2 * e], ti = t[2 * e + 1];
        g[2 * e] = sc * (re * tr_ + im * ti);       // conj(tr) * t
        g[2 * e + 1] = sc * (re * ti - im * tr_);
    }
}

// deterministic two-stage sum: stage 1 -> partial[blockIdx], stage 2 (one block) -> out[0]
template <typename T>
__global__ void __launch_bounds__(256) sum_stage_kernel(const T* __restrict__ x, long long n, T* __restrict__ out) {
    __shared__ double red[256];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) acc += (double)x[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if (threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = (T)red[0];
}

// ---- pulse heads: the element-wise tail of the pulse generators, one launch
// each way instead of 6-8 ATen kernels.  mode 0 = transformer head (model/universal_model.py:131-143):
//   u = sigmoid(x); p = lo + (hi-lo) u; [p = scale*p + base]; tau = relu(tau); phi = wrap(phi + offset_b)
// mode 1 = GRAPE head (model/GRAPE_model.py:76-89): (ux,uy,ut) = sigmoid(x); phi_u = atan2(uy,ux);
//   phi = lo0 + (hi0-lo0) phi_u; tau = relu(lo1 + (hi1-lo1) ut)
template <typename T>
struct HeadParams {
    const T* x;        // (B, L, P_in)   P_in = 2 (mode 0) or 3 (mode 1)
    const T* offset;   // (B) or nullptr   (mode 0: target azimuth added to phi)
    const T* base;     // (L, 2) or nullptr (mode 0: finetune base pulse)
    const T* gout;     // backward: (B, L, 2)
    T* out;            // forward: (B, L, 2);  backward: (B, L, P_in)
    long long n;       // B*L
    int L, mode;
    T lo0, hi0, lo1, hi1, scale;
};
template <typename T>
__device__ __forceinline__ T sigmoid_t(T v) { return (T)1 / ((T)1 + exp(-v)); }

template <typename T, bool BWD>
__global__ void pulse_head_kernel(const HeadParams<T> p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const T PI = (T)3.14159265358979323846;
    if (p.mode == 0) {
        const T u0 = sigmoid_t(p.x[2 * i]), u1 = sigmoid_t(p.x[2 * i + 1]);
        T phi = p.lo0 + (p.hi0 - p.lo0) * u0, tau = p.lo1 + (p.hi1 - p.lo1) * u1;
        if (p.base != nullptr) {
            const int l = (int)(i % p.L);
            phi = p.scale * phi + p.base[2 * l];
            tau = p.scale * tau + p.base[2 * l + 1];
        }
        if (!BWD) {
            tau = tau > (T)0 ? tau : (T)0;
            if (p.offset != nullptr) phi += p.offset[i / p.L];
            // python's float modulo: (phi + pi) % (2 pi) - pi, result of % has the sign of the divisor
            T m = fmod(phi + PI, (T)2 * PI);
            if (m < (T)0) m += (T)2 * PI;
            p.out[2 * i] = m - PI;
            p.out[2 * i + 1] = tau;
        } else {
            const T sc = p.base != nullptr ? p.scale : (T)1;
            p.out[2 * i] = p.gout[2 * i] * sc * (p.hi0 - p.lo0) * u0 * ((T)1 - u0);
            p.out[2 * i + 1] = tau > (T)0 ? p.gout[2 * i + 1] * sc * (p.hi1 - p.lo1) * u1 * ((T)1 - u1) : (T)0;
        }
    } else {
        const T ux = sigmoid_t(p.x[3 * i]), uy = sigmoid_t(p.x[3 * i + 1]), ut = sigmoid_t(p.x[3 * i + 2]);
        const T tau = p.lo1 + (p.hi1 - p.lo1) * ut;
        if (!BWD) {
            p.out[2 * i] = p.lo0 + (p.hi0 - p.lo0) * atan2(uy, ux);
            p.out[2 * i + 1] = tau > (T)0 ? tau : (T)0;
        } else {
            const T den = ux * ux + uy * uy;
            const T gphi = p.gout[2 * i] * (p.hi0 - p.lo0) / den;
            p.out[3 * i] = gphi * (-uy) * ux * ((T)1 - ux);
            p.out[3 * i + 1] = gphi * ux * uy * ((T)1 - uy);
            p.out[3 * i + 2] = tau > (T)0 ? p.gout[2 * i + 1] * (p.hi1 - p.lo1) * ut * ((T)1 - ut) : (T)0;
        }
    }
}

// ------------------------------------------------------------------ typed entry helpers
struct LossSpec {
    double n_total, tau, k;
    int kind;
    void* loss_out;
};
// peer-memory exchange of [G | Fsum] fused into the partials reduction (uqoc_su2_fwdbwd_peer)
struct PeerSpec {
    int rank, world;
    const uint64_t* data;     // host arrays [world] of device addresses
    const uint64_t* flags;
    unsigned epoch;
};

// pulse head folded into the op (uqoc_su2_head_step): host-side description
struct HeadArgs {
    int mode;                 // 1 transformer head (logits (B, L, 2)), 2 GRAPE head (logits (B, L, 3))
    double lo0, hi0, lo1, hi1, scale;
    const void* offset;       // (B) or null
    const void* base;         // (L, 2) or null
    void* pulses_out;         // (B, L, 2) or null
};

template <typename T>
static int su2_run(const void* pulses, const void* target_c, const void* err, const void* weight, int64_t B, int64_t L,
                   int64_t M, int64_t j0, double sig_d, double sig_e, uint64_t seed, uint64_t offset, void* U_out,
                   void* F_out, void* err_out, void* Fsum, void* G, void* workspace, int64_t workspace_bytes, int dtype,
                   unsigned flags, bool bwd, cudaStream_t stream, int grid_ne = 0, const void* sig_tab = nullptr,
                   const LossSpec* ls = nullptr, const PeerSpec* peer = nullptr, int64_t b0 = 0, const HeadArgs* head = nullptr) {
    Su2Plan plan = make_plan(B, L, M, dtype, flags, bwd);
    UQOC_CHECK_ARG(b0 >= 0 && b0 + B <= (1LL << 31), "target offset b0 out of range: %lld", (long long)b0);
    UQOC_CHECK_ARG((int64_t)B * plan.cps <= 0x7fffffffLL, "grid too large: %lld blocks", (long long)(B * plan.cps));
    UQOC_CHECK_ARG((flags & UQOC_FLAG_RNG_FROM_DEVICE) || offset <= 0xffffffffULL,
                   "Philox offset must be < 2^32 (it is one 32-bit counter word), got %llu", (unsigned long long)offset);
    Su2Params<T> p;
    memset(&p, 0, sizeof(p));
    p.pulses = (const T*)pulses;
    p.target_c = (const T*)target_c;
    p.err = (const T*)err;
    p.weight = (const T*)weight;
    p.B = (int)B; p.L = (int)L; p.M = (int)M;
    p.n_tiles = plan.n_tiles; p.splits = plan.splits; p.cps = plan.cps; p.C = plan.C;
    p.j0 = j0;
    p.b0 = (int)b0;
    p.sig_d = (T)sig_d; p.sig_e = (T)sig_e;
    p.seed = seed; p.offset = offset;
    p.rng_dev = (flags & UQOC_FLAG_RNG_FROM_DEVICE) ? (const unsigned long long*)(uintptr_t)seed : nullptr;
    p.U_out = (T*)U_out; p.F_out = (T*)F_out; p.err_out = (T*)err_out;
    p.grid_ne = grid_ne; p.sig_tab = (const T*)sig_tab;
    p.raw_target = (flags & UQOC_FLAG_RAW_TARGET) ? 1 : 0;
    if (head != nullptr) {
        p.head.mode = head->mode;
        p.head.lo0 = (T)head->lo0; p.head.hi0 = (T)head->hi0; p.head.lo1 = (T)head->lo1; p.head.hi1 = (T)head->hi1;
        p.head.scale = (T)head->scale;
        p.head.offset = (const T*)head->offset; p.head.base = (const T*)head->base; p.head.pulses_out = (T*)head->pulses_out;
    }
    const int po = su2_grad_width(p);                      // reals per pulse of a gradient row (3 for the GRAPE head's logits)
    const int64_t n_g = bwd ? B * L * po : 0;
    const int64_t part_bytes = plan.cps > 1 ? (int64_t)plan.cps * (B + n_g) * (int64_t)sizeof(T) : 0;
    if (plan.cps > 1) {
        if (workspace == nullptr || workspace_bytes < kTicketBytes + part_bytes) {
            set_error("workspace too small: need %lld bytes, got %lld", (long long)(kTicketBytes + part_bytes), (long long)workspace_bytes);
            return UQOC_E_WORKSPACE;
        }
        p.G_part = (T*)((char*)workspace + kTicketBytes);          // rows 16-byte aligned when B*L is even
        p.Fsum_part = p.G_part + (size_t)plan.cps * n_g;
    } else {
        p.Fsum_part = (T*)Fsum;
        p.G_part = (T*)G;
    }
    // in-kernel epilogue: the last block reduces the partials, exchanges them with the peers and applies the loss
    const bool want_fin = bwd && Fsum != nullptr && (ls != nullptr || peer != nullptr || plan.cps > 1);
    plan.fin = want_fin && !(flags & UQOC_FLAG_NO_FIN) && workspace != nullptr && workspace_bytes >= kTicketBytes &&
               ((uintptr_t)workspace % 16 == 0) && ((uintptr_t)G % (4 * sizeof(T)) == 0) &&
               su2_fin_supported(B, L, plan.cps, (plan.packed ? kThreads * plan.vb : kThreads), po);
    if (plan.fin) {
        p.fin.ticket = (unsigned*)workspace;
        p.fin.Fsum = (T*)Fsum;
        p.fin.G = (T*)G;
        p.fin.kind = -1;
        if (ls != nullptr) {
            p.fin.kind = ls->kind; p.fin.n_total = ls->n_total; p.fin.tau = ls->tau; p.fin.k = ls->k;
            p.fin.loss_out = (T*)ls->loss_out;
        }
        if (peer != nullptr) {
            for (int q = 0; q < peer->world; ++q) {
                p.fin.pp.data[q] = (T*)(uintptr_t)peer->data[q];
                p.fin.pp.flags[q] = (unsigned*)(uintptr_t)peer->flags[q];
            }
            p.fin.pp.rank = peer->rank; p.fin.pp.world = peer->world; p.fin.pp.epoch = peer->epoch;
            p.fin.pp.n_pad = (n_g + B + 31) / 32 * 32;
        }
    }
    int rc;
    if (dtype == UQOC_F64) rc = (flags & UQOC_FLAG_NO_TABLE) ? su2_launch<T, SC_LIBM>(p, plan, bwd, stream) : su2_launch<T, SC_TABLE>(p, plan, bwd, stream);
    else if (flags & UQOC_FLAG_FAST_SINCOS) rc = su2_launch<T, SC_MUFU>(p, plan, bwd, stream);
    else rc = su2_launch<T, SC_POLY>(p, plan, bwd, stream);
    if (rc != 0 || plan.fin) return rc;
    if (peer != nullptr) {
        PeerParams<T> pp;
        for (int q = 0; q < peer->world; ++q) {
            pp.data[q] = (T*)(uintptr_t)peer->data[q];
            pp.flags[q] = (unsigned*)(uintptr_t)peer->flags[q];
        }
        pp.rank = peer->rank; pp.world = peer->world; pp.epoch = peer->epoch;
        const long long n = n_g + B;
        pp.n_pad = (n + 31) / 32 * 32;
        long long blocks = (n + 31) / 32;
        const long long cap = (long long)sm_count() * 2;          // 1024-thread blocks: 2 resident per SM
        if (blocks > cap) blocks = cap;
        if (blocks > kPeerMaxBlocks) blocks = kPeerMaxBlocks;
        if constexpr (std::is_same<T, float>::value) {
            // FP32: {value, epoch} words, no fence / flag round trip, loss epilogue in the same dependent launch
            launch_dependent(su2_reduce_exchange_ll<32>, (unsigned)blocks, 1024, 0, stream, (const float*)p.Fsum_part,
                             (const float*)p.G_part, plan.cps, (int)B, (long long)n_g, pp, ls ? ls->n_total : 1.0,
                             ls ? ls->kind : -1, ls ? ls->tau : 0.0, ls ? ls->k : 0.0, (float*)Fsum, (float*)G,
                             ls ? (float*)ls->loss_out : (float*)nullptr);
            return launch_status("su2_reduce_exchange_ll");
        }
        launch_dependent(su2_reduce_exchange<T, 32>, (unsigned)blocks, 1024, 0, stream, (const T*)p.Fsum_part, (const T*)p.G_part,
                         plan.cps, (int)B, (long long)n_g, pp, (T*)Fsum, (T*)G);
        rc = launch_status("su2_reduce_exchange");
        if (rc) return rc;
    } else if (plan.cps > 1 && (Fsum != nullptr || n_g > 0)) {
        // few outputs: ONE kernel reduces, evaluates the loss (every block redundantly, from all cps x B partial sums) and
        // scales; many outputs: the redundant part dominates, two dependent-launched kernels are faster (measured)
        if (ls != nullptr && (int64_t)plan.cps * B <= 4096 && (n_g + B + 31) / 32 <= 2 * (int64_t)sm_count() && Fsum != nullptr) {
            const long long n = n_g + B;
            launch_dependent(su2_reduce_finalize<T>, (unsigned)((n + 31) / 32), 1024, 0, stream, (const T*)p.Fsum_part,
                             (const T*)p.G_part, plan.cps, (int)B, (long long)n_g, ls->n_total, ls->kind, ls->tau, ls->k, (T*)Fsum,
                             (T*)G, (T*)ls->loss_out);
            return launch_status("su2_reduce_finalize");
        }
        launch_reduce_partials<T>(p.Fsum_part, p.G_part, plan.cps, (int)B, n_g, (T*)Fsum, (T*)G, stream);
        rc = launch_status("su2_reduce_partials");
        if (rc) return rc;
    }
    if (ls != nullptr)
        return uqoc_loss_finalize(Fsum, B, ls->n_total, ls->kind, ls->tau, ls->k, G, n_g, ls->loss_out, dtype, (void*)stream);
    return 0;
}


static int check_common(int64_t B, int64_t L, int64_t M, int dtype) {
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "dtype must be UQOC_F32 or UQOC_F64, got %d", dtype);
    UQOC_CHECK_ARG(B >= 1 && B <= (1 << 24), "B out of range: %lld", (long long)B);
    UQOC_CHECK_ARG(L >= 1 && L <= (1 << 20), "L out of range: %lld", (long long)L);
    UQOC_CHECK_ARG(M >= 1 && M <= (1LL << 31) - 1, "M out of range: %lld", (long long)M);
    UQOC_CHECK_ARG(B * M <= (1LL << 40), "B*M too large: %lld", (long long)(B * M));
    return 0;
}

template <typename T, bool BWD>
static int pulse_head_run(const void* x, const void* offset, const void* base, const void* gout, void* out, int64_t B,
                          int64_t L, int mode, const double* ranges, double scale, cudaStream_t stream) {
    HeadParams<T> p;
    p.x = (const T*)x; p.offset = (const T*)offset; p.base = (const T*)base; p.gout = (const T*)gout; p.out = (T*)out;
    p.n = B * L; p.L = (int)L; p.mode = mode;
    p.lo0 = (T)ranges[0]; p.hi0 = (T)ranges[1]; p.lo1 = (T)ranges[2]; p.hi1 = (T)ranges[3]; p.scale = (T)scale;
    const unsigned blocks = (unsigned)((p.n + 255) / 256);
    pulse_head_kernel<T, BWD><<<blocks, 256, 0, stream>>>(p);
    return launch_status("pulse_head_kernel");
}


}  // namespace uqoc

using namespace uqoc;

extern "C" {

int uqoc_version(void) { return UQOC_VERSION; }
const char* uqoc_last_error(void) { return g_err; }

int uqoc_su2_target_coeffs(const void* U_target, int64_t B, void* target_c, int dtype, void* stream) {
    UQOC_CHECK_ARG(U_target && target_c, "null pointer");
    UQOC_CHECK_ARG(B >= 1, "B must be >= 1");
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "bad dtype %d", dtype);
    const unsigned blocks = (unsigned)((B + 127) / 128);
    if (dtype == UQOC_F64) target_coeffs_kernel<double><<<blocks, 128, 0, (cudaStream_t)stream>>>((const double*)U_target, B, (double*)target_c);
    else target_coeffs_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>((const float*)U_target, B, (float*)target_c);
    return launch_status("target_coeffs_kernel");
}

int64_t uqoc_su2_workspace_bytes(int64_t B, int64_t L, int64_t M, int dtype, unsigned flags) {
    if (B < 1 || L < 1 || M < 1) return 0;
    const int64_t esz = dtype == UQOC_F64 ? 8 : 4;
    const Su2Plan pb = make_plan(B, L, M, dtype, flags, true), pf = make_plan(B, L, M, dtype, flags, false);
    const int64_t nb = pb.cps > 1 ? (int64_t)pb.cps * (B + B * L * 3) * esz : 0;      // fused fwd+bwd: [Fsum | G] rows (3 reals
                                                                                      // per pulse: the GRAPE head's logit rows)
    const int64_t nf = pf.cps > 1 ? (int64_t)pf.cps * B * esz : 0;                    // forward only: Fsum rows
    return kTicketBytes + (nb > nf ? nb : nf);
}

int uqoc_su2_fwdbwd(const void* pulses, const void* target_c, const void* err, const void* weight, int64_t B, int64_t L,
                    int64_t M, int64_t j0, double sig_d, double sig_e, uint64_t seed, uint64_t offset, void* F_out,
                    void* err_out, void* Fsum, void* G, void* workspace, int64_t workspace_bytes, int dtype,
                    unsigned flags, void* stream) {
    int rc = check_common(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && target_c && Fsum && G, "pulses, target_c, Fsum and G must be non-null");
    if (dtype == UQOC_F64)
        return su2_run<double>(pulses, target_c, err, weight, B, L, M, j0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out,
                               Fsum, G, workspace, workspace_bytes, dtype, flags, true, (cudaStream_t)stream);
    return su2_run<float>(pulses, target_c, err, weight, B, L, M, j0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out, Fsum,
                          G, workspace, workspace_bytes, dtype, flags, true, (cudaStream_t)stream);
}

int uqoc_su2_fwdbwd_slice(const void* pulses, const void* target_c, const void* err, const void* weight, int64_t B, int64_t L,
                          int64_t M, int64_t j0, int64_t b0, double sig_d, double sig_e, uint64_t seed, uint64_t offset, void* F_out,
                          void* err_out, void* Fsum, void* G, void* workspace, int64_t workspace_bytes, int dtype,
                          unsigned flags, void* stream) {
    int rc = check_common(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && target_c && Fsum && G, "pulses, target_c, Fsum and G must be non-null");
    if (dtype == UQOC_F64)
        return su2_run<double>(pulses, target_c, err, weight, B, L, M, j0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out,
                               Fsum, G, workspace, workspace_bytes, dtype, flags, true, (cudaStream_t)stream, 0, nullptr, nullptr,
                               nullptr, b0);
    return su2_run<float>(pulses, target_c, err, weight, B, L, M, j0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out, Fsum,
                          G, workspace, workspace_bytes, dtype, flags, true, (cudaStream_t)stream, 0, nullptr, nullptr, nullptr, b0);
}

int64_t uqoc_peer_data_bytes(int64_t n, int world, int dtype) {
    if (n < 1 || world < 1) return 0;
    (void)dtype;                // 8 bytes per exchanged real either way: FP64 values, or FP32 {value, epoch} words
    return 2 * (int64_t)world * ((n + 31) / 32 * 32) * 8;
}
int64_t uqoc_peer_flag_bytes(int world) { return world < 1 ? 0 : (int64_t)world * kPeerMaxBlocks * (int64_t)sizeof(unsigned); }

int uqoc_su2_fwdbwd_peer(const void* pulses, const void* target_c, const void* err, const void* weight, int64_t B, int64_t L,
                         int64_t M, int64_t j0, double sig_d, double sig_e, uint64_t seed, uint64_t offset, void* F_out,
                         void* err_out, void* Fsum, void* G, void* workspace, int64_t workspace_bytes, int rank, int world,
                         const uint64_t* peer_data, const uint64_t* peer_flags, uint32_t epoch, int dtype, unsigned flags,
                         void* stream) {
    int rc = check_common(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && target_c && Fsum && G, "pulses, target_c, Fsum and G must be non-null");
    UQOC_CHECK_ARG(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bad rank/world %d/%d (max %d ranks)", rank,
                   world, kPeerMaxWorld);
    UQOC_CHECK_ARG(peer_data && peer_flags && epoch != 0, "peer_data, peer_flags must be non-null and epoch non-zero");
    for (int q = 0; q < world; ++q) UQOC_CHECK_ARG(peer_data[q] && peer_flags[q], "null peer pointer for rank %d", q);
    PeerSpec ps{rank, world, peer_data, peer_flags, epoch};
    if (dtype == UQOC_F64)
        return su2_run<double>(pulses, target_c, err, weight, B, L, M, j0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out,
                               Fsum, G, workspace, workspace_bytes, dtype, flags, true, (cudaStream_t)stream, 0, nullptr, nullptr, &ps);
    return su2_run<float>(pulses, target_c, err, weight, B, L, M, j0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out, Fsum,
                          G, workspace, workspace_bytes, dtype, flags, true, (cudaStream_t)stream, 0, nullptr, nullptr, &ps);
}

int uqoc_su2_fwdbwd_peer_loss(const void* pulses, const void* target_c, const void* err, int64_t B, int64_t L, int64_t M,
                              int64_t j0, int64_t M_global, double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                              int loss_kind, double tau, double k, void* F_out, void* err_out, void* Fsum, void* G,
                              void* loss_out, void* workspace, int64_t workspace_bytes, int rank, int world,
                              const uint64_t* peer_data, const uint64_t* peer_flags, uint32_t epoch, int dtype, unsigned flags,
                              void* stream) {
    int rc = check_common(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && target_c && Fsum && G && loss_out, "pulses, target_c, Fsum, G and loss_out must be non-null");
    UQOC_CHECK_ARG(loss_kind >= 0 && loss_kind <= 3, "unknown loss kind %d", loss_kind);
    UQOC_CHECK_ARG(M_global >= M, "M_global (%lld) must be >= M (%lld)", (long long)M_global, (long long)M);
    UQOC_CHECK_ARG(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bad rank/world %d/%d (max %d ranks)", rank,
                   world, kPeerMaxWorld);
    UQOC_CHECK_ARG(peer_data && peer_flags && epoch != 0, "peer_data, peer_flags must be non-null and epoch non-zero");
    for (int q = 0; q < world; ++q) UQOC_CHECK_ARG(peer_data[q] && peer_flags[q], "null peer pointer for rank %d", q);
    PeerSpec ps{rank, world, peer_data, peer_flags, epoch};
    LossSpec ls{(double)B * (double)M_global, tau, k, loss_kind, loss_out};
    if (dtype == UQOC_F64)
        return su2_run<double>(pulses, target_c, err, nullptr, B, L, M, j0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out,
                               Fsum, G, workspace, workspace_bytes, dtype, flags, true, (cudaStream_t)stream, 0, nullptr, &ls, &ps);
    return su2_run<float>(pulses, target_c, err, nullptr, B, L, M, j0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out, Fsum,
                          G, workspace, workspace_bytes, dtype, flags, true, (cudaStream_t)stream, 0, nullptr, &ls, &ps);
}

int uqoc_su2_fwdbwd_loss(const void* pulses, const void* target_c, const void* err, int64_t B, int64_t L, int64_t M,
                         double sig_d, double sig_e, uint64_t seed, uint64_t offset, int loss_kind, double tau, double k,
                         void* F_out, void* err_out, void* Fsum, void* G, void* loss_out, void* workspace,
                         int64_t workspace_bytes, int dtype, unsigned flags, void* stream) {
    int rc = check_common(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && target_c && Fsum && G && loss_out, "pulses, target_c, Fsum, G and loss_out must be non-null");
    UQOC_CHECK_ARG(loss_kind >= 0 && loss_kind <= 3, "unknown loss kind %d", loss_kind);
    LossSpec ls{(double)B * (double)M, tau, k, loss_kind, loss_out};
    if (dtype == UQOC_F64)
        return su2_run<double>(pulses, target_c, err, nullptr, B, L, M, 0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out,
                               Fsum, G, workspace, workspace_bytes, dtype, flags, true, (cudaStream_t)stream, 0, nullptr, &ls);
    return su2_run<float>(pulses, target_c, err, nullptr, B, L, M, 0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out, Fsum,
                          G, workspace, workspace_bytes, dtype, flags, true, (cudaStream_t)stream, 0, nullptr, &ls);
}

int uqoc_su2_head_step(const void* logits, int head_mode, const double* ranges, double scale, const void* phi_offset,
                       const void* base_pulse, const void* target_c, const void* err, int64_t B, int64_t L, int64_t M,
                       double sig_d, double sig_e, uint64_t seed, uint64_t offset, int loss_kind, double tau, double k,
                       void* pulses_out, void* F_out, void* err_out, void* Fsum, void* G, void* loss_out, void* workspace,
                       int64_t workspace_bytes, int dtype, unsigned flags, void* stream) {
    int rc = check_common(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(logits && target_c && Fsum && ranges, "logits, ranges, target_c and Fsum must be non-null");
    UQOC_CHECK_ARG(head_mode == 0 || head_mode == 1, "head_mode must be 0 (transformer) or 1 (GRAPE), got %d", head_mode);
    UQOC_CHECK_ARG(loss_kind >= -1 && loss_kind <= 3, "unknown loss kind %d", loss_kind);
    UQOC_CHECK_ARG(loss_kind < 0 || loss_out != nullptr, "loss_out must be non-null when a loss is requested");
    UQOC_CHECK_ARG(head_mode == 0 || (phi_offset == nullptr && base_pulse == nullptr), "the GRAPE head takes no offset / base pulse");
    const HeadArgs head{head_mode + 1, ranges[0], ranges[1], ranges[2], ranges[3], scale, phi_offset, base_pulse, pulses_out};
    LossSpec ls{(double)B * (double)M, tau, k, loss_kind, loss_out};
    const bool bwd = G != nullptr;
    if (dtype == UQOC_F64)
        rc = su2_run<double>(logits, target_c, err, nullptr, B, L, M, 0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out, Fsum,
                             G, workspace, workspace_bytes, dtype, flags, bwd, (cudaStream_t)stream, 0, nullptr,
                             (bwd && loss_kind >= 0) ? &ls : nullptr, nullptr, 0, &head);
    else
        rc = su2_run<float>(logits, target_c, err, nullptr, B, L, M, 0, sig_d, sig_e, seed, offset, nullptr, F_out, err_out, Fsum,
                            G, workspace, workspace_bytes, dtype, flags, bwd, (cudaStream_t)stream, 0, nullptr,
                            (bwd && loss_kind >= 0) ? &ls : nullptr, nullptr, 0, &head);
    if (rc || bwd || loss_kind < 0) return rc;
    return uqoc_loss_finalize(Fsum, B, ls.n_total, loss_kind, tau, k, nullptr, 0, loss_out, dtype, stream);   // forward only
}

int uqoc_su2_forward(const void* pulses, const void* target_c, const void* err, int64_t B, int64_t L, int64_t M, int64_t j0,
                     double sig_d, double sig_e, uint64_t seed, uint64_t offset, void* U_out, void* F_out, void* err_out,
                     void* Fsum, void* workspace, int64_t workspace_bytes, int dtype, unsigned flags, void* stream) {
    int rc = check_common(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && target_c, "pulses and target_c must be non-null");
    UQOC_CHECK_ARG(U_out || F_out || Fsum || err_out, "no output requested");
    if (dtype == UQOC_F64)
        return su2_run<double>(pulses, target_c, err, nullptr, B, L, M, j0, sig_d, sig_e, seed, offset, U_out, F_out, err_out,
                               Fsum, nullptr, workspace, workspace_bytes, dtype, flags, false, (cudaStream_t)stream);
    return su2_run<float>(pulses, target_c, err, nullptr, B, L, M, j0, sig_d, sig_e, seed, offset, U_out, F_out, err_out, Fsum,
                          nullptr, workspace, workspace_bytes, dtype, flags, false, (cudaStream_t)stream);
}

int uqoc_su2_forward_grid(const void* pulses, const void* target_c, const void* axis_delta, int64_t n_delta,
                          const void* axis_eps, int64_t n_eps, int64_t B, int64_t L, void* U_out, void* F_out,
                          void* Fsum, void* workspace, int64_t workspace_bytes, int dtype, unsigned flags, void* stream) {
    UQOC_CHECK_ARG(n_delta >= 1 && n_eps >= 1 && n_eps < (1LL << 31), "grid axes must be non-empty");
    const int64_t M = n_delta * n_eps;
    int rc = check_common(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && target_c && axis_delta && axis_eps, "null pointer");
    UQOC_CHECK_ARG(U_out || F_out || Fsum, "no output requested");
    const int64_t esz = dtype == UQOC_F64 ? 8 : 4;
    UQOC_CHECK_ARG((const char*)axis_eps == (const char*)axis_delta + n_delta * esz,
                   "axis_eps must directly follow axis_delta in one buffer [delta axis | eps axis]");
    if (dtype == UQOC_F64)
        return su2_run<double>(pulses, target_c, axis_delta, nullptr, B, L, M, 0, 0, 0, 0, 0, U_out, F_out, nullptr, Fsum,
                               nullptr, workspace, workspace_bytes, dtype, flags, false, (cudaStream_t)stream, (int)n_eps);
    return su2_run<float>(pulses, target_c, axis_delta, nullptr, B, L, M, 0, 0, 0, 0, 0, U_out, F_out, nullptr, Fsum, nullptr,
                          workspace, workspace_bytes, dtype, flags, false, (cudaStream_t)stream, (int)n_eps);
}

int uqoc_su2_forward_sigmas(const void* pulses, const void* target_c, const void* sigma_table, int64_t B, int64_t L,
                            int64_t M, int64_t j0, uint64_t seed, uint64_t offset, void* F_out, void* Fsum,
                            void* workspace, int64_t workspace_bytes, int dtype, unsigned flags, void* stream) {
    int rc = check_common(B, L, M, dtype);
    if (rc) return rc;
    UQOC_CHECK_ARG(pulses && target_c && sigma_table, "null pointer");
    UQOC_CHECK_ARG(F_out || Fsum, "no output requested");
    if (dtype == UQOC_F64)
        return su2_run<double>(pulses, target_c, nullptr, nullptr, B, L, M, j0, 0, 0, seed, offset, nullptr, F_out, nullptr,
                               Fsum, nullptr, workspace, workspace_bytes, dtype, flags, false, (cudaStream_t)stream, 0, sigma_table);
    return su2_run<float>(pulses, target_c, nullptr, nullptr, B, L, M, j0, 0, 0, seed, offset, nullptr, F_out, nullptr, Fsum,
                          nullptr, workspace, workspace_bytes, dtype, flags, false, (cudaStream_t)stream, 0, sigma_table);
}

int uqoc_su2_generator_forward(const void* pulses, const void* err, int64_t Bm, int64_t L, void* U_out, int dtype,
                               unsigned flags, void* stream) {
    (void)flags;
    UQOC_CHECK_ARG(pulses && err && U_out, "null pointer");
    UQOC_CHECK_ARG(Bm >= 1 && L >= 1, "Bm and L must be >= 1");
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "bad dtype %d", dtype);
    const unsigned blocks = (unsigned)((Bm + 127) / 128);
    if (dtype == UQOC_F64)
        su2_generator_fwd_kernel<double, SC_LIBM><<<blocks, 128, 0, (cudaStream_t)stream>>>((const double*)pulses, (const double*)err, Bm, (int)L, (double*)U_out);
    else
        su2_generator_fwd_kernel<float, SC_LIBM><<<blocks, 128, 0, (cudaStream_t)stream>>>((const float*)pulses, (const float*)err, Bm, (int)L, (float*)U_out);
    return launch_status("su2_generator_fwd_kernel");
}

int uqoc_su2_generator_backward(const void* pulses, const void* err, const void* grad_U, int64_t Bm, int64_t L,
                                void* grad_pulses, int dtype, unsigned flags, void* stream) {
    (void)flags;
    UQOC_CHECK_ARG(pulses && err && grad_U && grad_pulses, "null pointer");
    UQOC_CHECK_ARG(Bm >= 1 && L >= 1, "Bm and L must be >= 1");
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "bad dtype %d", dtype);
    const unsigned blocks = (unsigned)((Bm + 127) / 128);
    if (dtype == UQOC_F64)
        su2_generator_bwd_kernel<double, SC_LIBM><<<blocks, 128, 0, (cudaStream_t)stream>>>((const double*)pulses, (const double*)err, (const double*)grad_U, Bm, (int)L, (double*)grad_pulses);
    else
        su2_generator_bwd_kernel<float, SC_LIBM><<<blocks, 128, 0, (cudaStream_t)stream>>>((const float*)pulses, (const float*)err, (const float*)grad_U, Bm, (int)L, (float*)grad_pulses);
    return launch_status("su2_generator_bwd_kernel");
}

int uqoc_pulse_head_forward(const void* logits, const void* phi_offset, const void* base_pulse, int64_t B, int64_t L,
                            int mode, const double* ranges, double scale, void* pulses, int dtype, void* stream) {
    UQOC_CHECK_ARG(logits && pulses && ranges, "null pointer");
    UQOC_CHECK_ARG(B >= 1 && L >= 1 && (mode == 0 || mode == 1), "bad B/L/mode");
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "bad dtype %d", dtype);
    if (dtype == UQOC_F64) return pulse_head_run<double, false>(logits, phi_offset, base_pulse, nullptr, pulses, B, L, mode, ranges, scale, (cudaStream_t)stream);
    return pulse_head_run<float, false>(logits, phi_offset, base_pulse, nullptr, pulses, B, L, mode, ranges, scale, (cudaStream_t)stream);
}

int uqoc_pulse_head_backward(const void* logits, const void* base_pulse, const void* grad_pulses, int64_t B, int64_t L,
                             int mode, const double* ranges, double scale, void* grad_logits, int dtype, void* stream) {
    UQOC_CHECK_ARG(logits && grad_pulses && grad_logits && ranges, "null pointer");
    UQOC_CHECK_ARG(B >= 1 && L >= 1 && (mode == 0 || mode == 1), "bad B/L/mode");
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "bad dtype %d", dtype);
    if (dtype == UQOC_F64) return pulse_head_run<double, true>(logits, nullptr, base_pulse, grad_pulses, grad_logits, B, L, mode, ranges, scale, (cudaStream_t)stream);
    return pulse_head_run<float, true>(logits, nullptr, base_pulse, grad_pulses, grad_logits, B, L, mode, ranges, scale, (cudaStream_t)stream);
}

int uqoc_loss_finalize(const void* Fsum, int64_t B, double n_total, int loss_kind, double tau, double k, void* G,
                       int64_t G_numel, void* loss_out, int dtype, void* stream) {
    UQOC_CHECK_ARG(Fsum != nullptr, "Fsum must be non-null");
    UQOC_CHECK_ARG(B >= 1 && n_total > 0, "B and n_total must be positive");
    UQOC_CHECK_ARG(loss_kind >= 0 && loss_kind <= 3, "unknown loss kind %d", loss_kind);
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "bad dtype %d", dtype);
    UQOC_CHECK_ARG(G == nullptr || G_numel >= 0, "bad G_numel");
    long long blocks = G ? (G_numel + 256 * 8 - 1) / (256 * 8) : 1;
    if (blocks < 1) blocks = 1;
    if (blocks > 1184) blocks = 1184;
    if (dtype == UQOC_F64)
        launch_dependent(loss_finalize_kernel<double>, (unsigned)blocks, 256, 0, (cudaStream_t)stream, (const double*)Fsum, (int)B, n_total,
                         loss_kind, tau, k, (double*)G, (long long)G_numel, (double*)loss_out);
    else
        launch_dependent(loss_finalize_kernel<float>, (unsigned)blocks, 256, 0, (cudaStream_t)stream, (const float*)Fsum, (int)B, n_total,
                         loss_kind, tau, k, (float*)G, (long long)G_numel, (float*)loss_out);
    return launch_status("loss_finalize_kernel");
}

int uqoc_philox_errors(int64_t B, int64_t M, int64_t j0, double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                       void* err_out, int dtype, void* stream) {
    UQOC_CHECK_ARG(err_out != nullptr, "err_out must be non-null");
    UQOC_CHECK_ARG(B >= 1 && M >= 1, "B and M must be >= 1");
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "bad dtype %d", dtype);
    const long long n = B * M;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (dtype == UQOC_F64)
        philox_errors_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(B, M, j0, sig_d, sig_e, seed, (unsigned)offset, (double*)err_out);
    else
        philox_errors_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(B, M, j0, (float)sig_d, (float)sig_e, seed, (unsigned)offset, (float*)err_out);
    return launch_status("philox_errors_kernel");
}

int uqoc_fp32_peak_probe(int iters, int dtype, double* tflops, double* ms_out) {
    UQOC_CHECK_ARG(iters >= 1 && tflops != nullptr, "bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device");
        return UQOC_E_NODEVICE;
    }
    const int sms = sm_count();
    const int blocks = sms * 4, threads = 512;
    void* out = nullptr;
    cudaError_t e = cudaMalloc(&out, 64);   // probe-only scratch (the data path never allocates)
    if (e != cudaSuccess) { set_error("cudaMalloc: %s", cudaGetErrorString(e)); return (int)e; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, 0);
        if (dtype == UQOC_F64) fma_probe_kernel<double><<<blocks, threads>>>(iters, 0.999999, 1e-7, (double*)out);
        else if (dtype == 2) ffma2_probe_kernel<<<blocks, threads>>>(iters, 0x3f7fffef3f7fffefULL, 0x33d6bf9533d6bf95ULL, (unsigned long long*)out);
        else fma_probe_kernel<float><<<blocks, threads>>>(iters, 0.999999f, 1e-7f, (float*)out);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    int rc = launch_status("fma_probe_kernel");
    if (rc) return rc;
    const double flops = (double)blocks * threads * (double)iters * 64.0 * 2.0;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}

int uqoc_fidelity_forward(const void* U_out, const void* U_target, int64_t Bm, int d, int64_t target_stride, void* F,
                          int dtype, void* stream) {
    UQOC_CHECK_ARG(U_out && U_target && F, "null pointer");
    UQOC_CHECK_ARG(Bm >= 1 && d >= 1 && d <= 16, "bad Bm/d");
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "bad dtype %d", dtype);
    const unsigned blocks = (unsigned)((Bm + 255) / 256);
    if (dtype == UQOC_F64) fidelity_fwd_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double*)U_out, (const double*)U_target, Bm, d, target_stride, (double*)F);
    else fidelity_fwd_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)U_out, (const float*)U_target, Bm, d, target_stride, (float*)F);
    return launch_status("fidelity_fwd_kernel");
}

int uqoc_fidelity_backward(const void* U_out, const void* U_target, const void* grad_F, int64_t Bm, int d,
                           int64_t target_stride, void* grad_U, int dtype, void* stream) {
    UQOC_CHECK_ARG(U_out && U_target && grad_F && grad_U, "null pointer");
    UQOC_CHECK_ARG(Bm >= 1 && d >= 1 && d <= 16, "bad Bm/d");
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "bad dtype %d", dtype);
    const unsigned blocks = (unsigned)((Bm + 255) / 256);
    if (dtype == UQOC_F64) fidelity_bwd_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double*)U_out, (const double*)U_target, (const double*)grad_F, Bm, d, target_stride, (double*)grad_U);
    else fidelity_bwd_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)U_out, (const float*)U_target, (const float*)grad_F, Bm, d, target_stride, (float*)grad_U);
    return launch_status("fidelity_bwd_kernel");
}

int uqoc_sum(const void* x, int64_t n, void* out, void* workspace, int64_t workspace_bytes, int dtype, void* stream) {
    UQOC_CHECK_ARG(x && out, "null pointer");
    UQOC_CHECK_ARG(n >= 1, "n must be >= 1");
    UQOC_CHECK_ARG(dtype == UQOC_F32 || dtype == UQOC_F64, "bad dtype %d", dtype);
    const int64_t esz = dtype == UQOC_F64 ? 8 : 4;
    long long blocks = (n + 256 * 16 - 1) / (256 * 16);
    if (blocks > 1024) blocks = 1024;
    if (blocks <= 1) {
        if (dtype == UQOC_F64) sum_stage_kernel<double><<<1, 256, 0, (cudaStream_t)stream>>>((const double*)x, n, (double*)out);
        else sum_stage_kernel<float><<<1, 256, 0, (cudaStream_t)stream>>>((const float*)x, n, (float*)out);
        return launch_status("sum_stage_kernel");
    }
    if (workspace == nullptr || workspace_bytes < blocks * esz) {
        set_error("uqoc_sum workspace too small: need %lld bytes", (long long)(blocks * esz));
        return UQOC_E_WORKSPACE;
    }
    if (dtype == UQOC_F64) {
        sum_stage_kernel<double><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const double*)x, n, (double*)workspace);
        sum_stage_kernel<double><<<1, 256, 0, (cudaStream_t)stream>>>((const double*)workspace, blocks, (double*)out);
    } else {
        sum_stage_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)x, n, (float*)workspace);
        sum_stage_kernel<float><<<1, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, blocks, (float*)out);
    }
    return launch_status("sum_stage_kernel");
}

}  // extern "C"
