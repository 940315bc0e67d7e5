// SU(2) disorder-sampled propagation kernels (sm_100a).
//
// One thread owns ST error samples x one chunk of the pulse train; LPS lanes share a sample
// (chunked product + warp-shuffle associative scan) when there are too few samples to fill
// the machine, LPS = 1 (pure register chain) otherwise.  Pulses of the block's target are
// staged once in shared memory as (cos phi, sin phi, tau) rows and read as broadcast LDS.128.
//
// Forward  (SCORE.py:99-145, :168-183):  P <- q_i (x) P, F = ((cr.P)^2 + (ci.P)^2 + 2)/6.
// Backward (autograd of the same, trainer.py:90): with W_i = conj(S_i) (x) lambda (x) conj(P_i)
// (S_i = suffix product, P_i = prefix product, lambda = dF/dP_L) the pulse gradients are
//     dF/dtau_i = (1+eps)/2 * <W_i, n_i>,   dF/dphi_i = c s' B - s'^2 (delta A - W_z)
// and W_{i-1} = conj(q_i) W_i q_i is a 3-vector rotation about the pulse axis by -2h, done in
// the pulse's own (phi-rotated) frame.  W_L comes from the prefix scan alone because
// conj(S_i) = P_i (x) conj(P_L) for unit quaternions, so no L x B x S tensor is ever stored.
#pragma once
#include "uqoc_common.cuh"

namespace uqoc {

#include "uqoc_sincos_table.inc"

constexpr int kThreads = 128;  // 4 warps: one per SM sub-partition
constexpr int kWarps = kThreads / 32;

// ---- programmatic dependent launch: the epilogue kernels (partials reduction, exchange, loss) are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so their blocks are scheduled while the fused kernel drains and
// wait HERE for its completion and memory flush (a no-op for a kernel launched the ordinary way)
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- cross-GPU exchange over NVLink peer memory (uqoc_su2_fwdbwd_peer*) --------------------------------------
constexpr int kPeerMaxWorld = 16;
constexpr int kPeerMaxBlocks = 1024;      // flag row length; grid of su2_reduce_exchange <= this

template <typename T>
struct PeerParams {
    T* data[kPeerMaxWorld];               // rank q's exchange buffer as mapped in this process: [2][world][n_pad]
    unsigned* flags[kPeerMaxWorld];       // rank q's flag buffer: [world][kPeerMaxBlocks]
    int rank, world;                      // world == 0: no exchange
    unsigned epoch;
    long long n_pad;
};

// ---- in-kernel epilogue ("last block done"): partials reduction [+ peer exchange] [+ loss and chain factor] ----
// Every block bumps a ticket after its partials are globally visible; the block that draws the last ticket sums the
// partials in fixed order, optionally exchanges the totals with the peer GPUs, evaluates the loss on the pooled mean
// fidelity and scales the gradient -- the whole step is ONE launch (BASELINE config 3 is a 45 us kernel; the second
// and third launches used to cost another 10 us).  Used while the partials fit one block's L2 bandwidth.
template <typename T>
struct FinParams {
    unsigned* ticket;      // zero on entry; atomicInc wraps it back to zero (caller-owned workspace); null: epilogue off
    T* Fsum;               // (B) final per-target fidelity sums
    T* G;                  // (B, L, 2) final gradient
    T* loss_out;           // (3) {loss, Fbar, dloss/dFbar} or null
    double n_total, tau, k;
    int kind;              // UQOC_LOSS_* or < 0: reduction (and exchange) only
    PeerParams<T> pp;
};

// ---- pulse heads folded into the op ----------------------------------------------------
// The element-wise tail of the reference's pulse generators is applied while the pulse train is staged, and its
// backward while the gradient rows are written: no (B, L, 2) pulses tensor travels between model and op, and the
// 6-8 small ATen launches per direction of model/universal_model.py:131-143 / model/GRAPE_model.py:76-89 disappear.
//   mode 1 (transformer head): u = sigmoid(x); p = lo + (hi - lo) u; [p = scale p + base]; tau = relu(tau);
//                              phi = wrap(phi + offset_b) to [-pi, pi)                       logits (B, L, 2)
//   mode 2 (GRAPE head):       (ux, uy, ut) = sigmoid(x); phi = lo0 + (hi0 - lo0) atan2(uy, ux);
//                              tau = relu(lo1 + (hi1 - lo1) ut)                              logits (B, L, 3)
template <typename T>
struct HeadSpec {
    int mode;          // 0: `pulses` holds the pulses themselves
    T lo0, hi0, lo1, hi1, scale;
    const T* offset;   // (B) target azimuth or nullptr            (mode 1)
    const T* base;     // (L, 2) finetune base pulse or nullptr    (mode 1)
    T* pulses_out;     // (B, L, 2) or nullptr: the pulses the head produced (for logging / saving)
};
template <typename T>
__device__ __forceinline__ T head_sigmoid(T v) { return (T)1 / ((T)1 + exp(-v)); }

template <typename T>
struct Su2Params {
    const T* pulses;    // (B, L, 2), or the head's logits (B, L, 2 | 3) when head.mode != 0
    const T* target_c;  // (B, 8)
    const T* err;       // (2, B*M) or nullptr (Philox)
    const T* weight;    // (B*M) or nullptr
    int B, L, M;
    int n_tiles, splits;   // splits = sample-tile streams per target (blocks per target x virtual blocks per block)
    int cps;            // blocks per target: grid = B * cps, one partial [Fsum | G] per block
    int C;              // chunk length (padded to the gradient-buffer depth)
    long long j0;
    int b0;             // global index of target 0 of this call (Philox counter word; target-chunked launches)
    T sig_d, sig_e;
    unsigned long long seed;
    unsigned long long offset;
    T* U_out;      // (B*M, 2, 2, 2) or nullptr
    T* F_out;      // (B*M) or nullptr
    T* err_out;    // (2, B*M) or nullptr
    T* Fsum_part;  // [cps][B]
    T* G_part;     // [cps][B][L][2]
    // forward-only sweep modes (visualize/util.py:231-249, :313-326)
    const unsigned long long* rng_dev;   // non-null: {seed, offset} read from device memory (CUDA-graph replay)
    int grid_ne;          // > 0: err = [delta axis (M / grid_ne) | eps axis (grid_ne)], sample j -> (j / ne, j % ne)
    const T* sig_tab;     // non-null: per-target (sigma_delta, sigma_eps) rows for the Philox samples
    int raw_target;       // != 0: target_c holds the raw complex targets (B, 2, 2, 2) (UQOC_FLAG_RAW_TARGET)
    FinParams<T> fin;
    HeadSpec<T> head;
};

// reals per pulse of the gradient rows: the head's input width
template <typename T>
__host__ __device__ __forceinline__ int su2_grad_width(const Su2Params<T>& p) { return p.head.mode == 2 ? 3 : 2; }

// (phi, tau) of pulse i of target b: the stored pulse, or the head applied to its logits (same arithmetic as the
// stand-alone pulse_head_kernel, csrc/uqoc_api.cu)
template <typename T>
__device__ __forceinline__ void su2_pulse_at(const Su2Params<T>& p, int b, int i, T& phi, T& tau) {
    const HeadSpec<T>& h = p.head;
    if (h.mode == 0) {
        const T* pb = p.pulses + ((size_t)b * p.L + i) * 2;
        phi = pb[0];
        tau = pb[1];
    } else if (h.mode == 1) {
        const T PI = (T)3.14159265358979323846;
        const T* x = p.pulses + ((size_t)b * p.L + i) * 2;
        const T u0 = head_sigmoid(x[0]), u1 = head_sigmoid(x[1]);
        T ph = h.lo0 + (h.hi0 - h.lo0) * u0, ta = h.lo1 + (h.hi1 - h.lo1) * u1;
        if (h.base != nullptr) {
            ph = h.scale * ph + h.base[2 * i];
            ta = h.scale * ta + h.base[2 * i + 1];
        }
        ta = ta > (T)0 ? ta : (T)0;
        if (h.offset != nullptr) ph += h.offset[b];
        T m = fmod(ph + PI, (T)2 * PI);                    // python's float modulo: the result has the divisor's sign
        if (m < (T)0) m += (T)2 * PI;
        phi = m - PI;
        tau = ta;
    } else {
        const T* x = p.pulses + ((size_t)b * p.L + i) * 3;
        const T ux = head_sigmoid(x[0]), uy = head_sigmoid(x[1]), ut = head_sigmoid(x[2]);
        const T ta = h.lo1 + (h.hi1 - h.lo1) * ut;
        phi = h.lo0 + (h.hi0 - h.lo0) * atan2(uy, ux);
        tau = ta > (T)0 ? ta : (T)0;
    }
}
// gradient row of pulse i w.r.t. the head's inputs from (d/dphi, d/dtau); returns the row width
template <typename T>
__device__ __forceinline__ int su2_head_grad(const Su2Params<T>& p, int b, int i, T gphi, T gtau, T (&o)[3]) {
    const HeadSpec<T>& h = p.head;
    if (h.mode == 1) {
        const T* x = p.pulses + ((size_t)b * p.L + i) * 2;
        const T u0 = head_sigmoid(x[0]), u1 = head_sigmoid(x[1]);
        T ta = h.lo1 + (h.hi1 - h.lo1) * u1;
        if (h.base != nullptr) ta = h.scale * ta + h.base[2 * i + 1];
        const T sc = h.base != nullptr ? h.scale : (T)1;
        o[0] = gphi * sc * (h.hi0 - h.lo0) * u0 * ((T)1 - u0);
        o[1] = ta > (T)0 ? gtau * sc * (h.hi1 - h.lo1) * u1 * ((T)1 - u1) : (T)0;
        o[2] = (T)0;
        return 2;
    }
    const T* x = p.pulses + ((size_t)b * p.L + i) * 3;
    const T ux = head_sigmoid(x[0]), uy = head_sigmoid(x[1]), ut = head_sigmoid(x[2]);
    const T ta = h.lo1 + (h.hi1 - h.lo1) * ut;
    const T g = gphi * (h.hi0 - h.lo0) / (ux * ux + uy * uy);
    o[0] = g * (-uy) * ux * ((T)1 - ux);
    o[1] = g * ux * uy * ((T)1 - uy);
    o[2] = ta > (T)0 ? gtau * (h.hi1 - h.lo1) * ut * ((T)1 - ut) : (T)0;
    return 3;
}

// trace coefficients of target b: Tr(U^dagger T) = (cr + i ci) . q.  Either precomputed rows (uqoc_su2_target_coeffs)
// or formed here from the raw 2x2 complex target: c0 = T00+T11, c1 = i(T01+T10), c2 = T10-T01, c3 = i(T00-T11).
template <typename T>
__device__ __forceinline__ void su2_load_target(const Su2Params<T>& p, int b, T (&cr)[4], T (&ci)[4]) {
    const T* t = p.target_c + (size_t)b * 8;
    if (p.raw_target) {            // T00 (0,1) T01 (2,3) T10 (4,5) T11 (6,7)
        cr[0] = t[0] + t[6];    ci[0] = t[1] + t[7];
        cr[1] = -(t[3] + t[5]); ci[1] = t[2] + t[4];
        cr[2] = t[4] - t[2];    ci[2] = t[5] - t[3];
        cr[3] = -(t[1] - t[7]); ci[3] = t[0] - t[6];
    } else {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            cr[m] = t[m];
            ci[m] = t[4 + m];
        }
    }
}

// (delta, eps) of sample (b, j): explicit tensor, 1-D axes of a meshgrid('ij') sweep, or on-chip Philox
template <typename T>
__device__ __forceinline__ void su2_sample_errors(const Su2Params<T>& p, int b, long long j, size_t sidx, size_t Bm,
                                                  T& delta, T& eps) {
    if (p.grid_ne > 0) {
        const long long nd = p.M / p.grid_ne;
        delta = p.err[j / p.grid_ne];
        eps = p.err[nd + j % p.grid_ne];
    } else if (p.err != nullptr) {
        delta = p.err[sidx];
        eps = p.err[Bm + sidx];
    } else {
        const T sd = p.sig_tab != nullptr ? p.sig_tab[2 * b] : p.sig_d;
        const T se = p.sig_tab != nullptr ? p.sig_tab[2 * b + 1] : p.sig_e;
        const unsigned long long seed = p.rng_dev != nullptr ? p.rng_dev[0] : p.seed;
        const unsigned offset = (unsigned)(p.rng_dev != nullptr ? p.rng_dev[1] : p.offset);   // host rejects >= 2^32
        philox_delta_eps<T>((uint64_t)(p.j0 + j), (uint32_t)(p.b0 + b), seed, offset, sd, se, delta, eps);
    }
}

// gradient-buffer depth: steps buffered in registers before one cross-lane reduction
template <int LPS>
struct GradBuf {
    static constexpr int NB = (LPS == 1) ? 8 : 1;
};

// ---- transpose-reduce of N per-lane values over the lanes that differ in bits >= BIT --------
// While more than one value is left each stage halves the values per lane (a lane keeps the
// half selected by its bit and receives the partner's partial of that half); afterwards plain
// butterflies.  `base` returns the first original index of the values the lane ends up owning.
template <typename T, int N, int BIT>
struct LaneReduce {
    __device__ __forceinline__ static void run(T* v, int lane, int& base) {
        if constexpr (BIT < 32) {
            if constexpr (N > 1) {
                constexpr int H = N / 2;
                const bool hi = (lane & BIT) != 0;
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    const T keep = hi ? v[j + H] : v[j];
                    const T send = hi ? v[j] : v[j + H];
                    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
                }
                if (hi) base += H;
                LaneReduce<T, H, BIT * 2>::run(v, lane, base);
            } else {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], BIT);
                LaneReduce<T, 1, BIT * 2>::run(v, lane, base);
            }
        }
    }
};
// number of values a lane owns at the end and the lane bits over which the result is replicated
constexpr int reduce_final_count(int n, int bit) { return bit >= 32 ? n : reduce_final_count(n > 1 ? n / 2 : 1, bit * 2); }
constexpr int reduce_dup_mask(int n, int bit) {
    return bit >= 32 ? 0 : ((n > 1 ? 0 : bit) | reduce_dup_mask(n > 1 ? n / 2 : 1, bit * 2));
}

template <typename T>
__device__ __forceinline__ T shfl_up_t(T v, int d, int width) { return __shfl_up_sync(0xffffffffu, v, d, width); }
template <typename T>
__device__ __forceinline__ T shfl_t(T v, int src, int width) { return __shfl_sync(0xffffffffu, v, src, width); }

template <typename T>
__device__ __forceinline__ Quat<T> qshfl_up(const Quat<T>& q, int d, int width) {
    return Quat<T>{shfl_up_t(q.a, d, width), shfl_up_t(q.b, d, width), shfl_up_t(q.c, d, width), shfl_up_t(q.d, d, width)};
}
template <typename T>
__device__ __forceinline__ Quat<T> qshfl(const Quat<T>& q, int src, int width) {
    return Quat<T>{shfl_t(q.a, src, width), shfl_t(q.b, src, width), shfl_t(q.c, src, width), shfl_t(q.d, src, width)};
}

__device__ __forceinline__ void su2_loss_eval(double Fbar, int kind, double tau, double k, double& val, double& dval) {
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

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}


// 4 consecutive reals through L2 (other blocks' / peers' stores: never the non-coherent L1)
#ifndef UQOC_FIN_LD
#define UQOC_FIN_LD 0
#endif
__device__ __forceinline__ void ld4cg(const float* p, float (&v)[4]) {
#if UQOC_FIN_LD == 0
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
#elif UQOC_FIN_LD == 1
    const float4 t = *reinterpret_cast<const float4*>(p);
#elif UQOC_FIN_LD == 2
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
#else
    float4 t;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(p));
#endif
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4cg(const double* p, double (&v)[4]) {
    const double2 t0 = __ldcg(reinterpret_cast<const double2*>(p)), t1 = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
    reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
    reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}

// ---- in-kernel epilogue (FinParams): called by every thread of every block after the block's partial [G | Fsum] row
// has been written to G_part / Fsum_part.  `scr_raw` = shared scratch of su2_fin_smem_bytes() (the sweeps' shared memory
// is dead by now).  The output vector is small here (the host enables the epilogue only when n_g / 4 + B <= blockDim.x,
// n_g % 4 == 0): every output column -- 4 gradient reals or one Fsum -- is owned by ONE thread per part-lane, all of a
// thread's loads are independent 16-byte L2 loads, the column total stays in registers through the exchange and the
// loss, and G is written once, already scaled.  Fixed summation order everywhere: bit-reproducible, and with a peer
// exchange bit-identical on every rank (slots are summed in rank order).
__host__ __device__ inline size_t su2_fin_smem_bytes(int nthr, size_t elem);
#ifdef UQOC_FIN_INLINE
#define UQOC_FIN_ATTR __forceinline__
#else
#define UQOC_FIN_ATTR __noinline__     // out of line: keeps the sweeps' register allocation independent of the epilogue
#endif
template <typename T>
__device__ UQOC_FIN_ATTR void su2_block_finalize(const FinParams<T>& f, const T* __restrict__ G_part, const T* __restrict__ Fsum_part,
                                                 const int parts, const int B, const int L, unsigned char* scr_raw,
                                                 const int po = 2 /* reals per pulse of a gradient row */) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    double* dred = reinterpret_cast<double*>(scr_raw);                       // [40]: warp sums, scale, flags
    T* red = reinterpret_cast<T*>(scr_raw + 40 * sizeof(double));            // [nthr][4] part-lane partials
    T* fsm = red + (size_t)nthr * 4;                                         // [B] per-target fidelity sums
#ifdef UQOC_FIN_TIMING
    unsigned long long* stamp = reinterpret_cast<unsigned long long*>(f.ticket) + 2;
    auto now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    const unsigned long long ts_enter = now();
#endif
    __syncthreads();                                               // the block's partial row is complete ...
    if (tid == 0) {
        __threadfence();                                           // ... and (cumulatively) visible device-wide before its
        const unsigned t = atomicInc(f.ticket, gridDim.x - 1);     // ticket is; the counter wraps to 0 after the last one
        const bool last = (t == gridDim.x - 1);
        if (last) __threadfence();
        dred[39] = last ? 1.0 : 0.0;
    }
    __syncthreads();
    if (dred[39] == 0.0) return;
#ifdef UQOC_FIN_TIMING
    if (tid == 0) { stamp[0] = ts_enter; stamp[1] = now(); }
#endif
    const int n_g = B * L * po, ncol4 = n_g / 4, ncol = ncol4 + B;
    UQOC_ASSERT(ncol <= nthr && n_g % 4 == 0 && su2_fin_smem_bytes(nthr, sizeof(T)) <= dyn_smem_bytes());
    int Y = nthr / ncol;                                           // part-lanes per column
    if (Y > parts) Y = parts;
    const int cpp = nthr / Y;                                      // columns per part-lane (>= ncol)
    const int y = tid / cpp, cx = tid % cpp;
    const bool owner = (y == 0) && cx < ncol;                      // holds the column's total from here on
    // ---- 1. fixed-order sum over the blocks' partial rows
    T a[4] = {(T)0, (T)0, (T)0, (T)0};
    if (y < Y && cx < ncol) {
        if (cx < ncol4) {
            const T* src = G_part + 4 * cx;
#pragma unroll 8
            for (int s = y; s < parts; s += Y) {
                T v[4];
                ld4cg(src + (size_t)s * n_g, v);
                a[0] += v[0]; a[1] += v[1]; a[2] += v[2]; a[3] += v[3];
            }
        } else {
            const T* src = Fsum_part + (cx - ncol4);
#pragma unroll 8
            for (int s = y; s < parts; s += Y) a[0] += __ldcg(src + (size_t)s * B);
        }
    }
    if (Y > 1) {
        st4(red + 4 * (size_t)tid, a);
        __syncthreads();
        if (owner) {
            for (int yy = 1; yy < Y; ++yy) {
                const T* r = red + 4 * (size_t)(yy * cpp + cx);
                a[0] += r[0]; a[1] += r[1]; a[2] += r[2]; a[3] += r[3];
            }
        }
    }
#ifdef UQOC_FIN_TIMING
    __syncthreads();
    if (tid == 0) stamp[2] = now();
#endif
    // ---- 2. peer exchange: push into slot[rank] of every rank (NVLink stores), flags up, wait, sum in rank order
    if (f.pp.world > 0) {
        const size_t set_off = (size_t)(f.pp.epoch & 1u) * f.pp.world * f.pp.n_pad;
        const size_t col_off = cx < ncol4 ? 4 * (size_t)cx : (size_t)n_g + (cx - ncol4);
        if (owner) {
            const size_t o = set_off + (size_t)f.pp.rank * f.pp.n_pad + col_off;
            for (int q = 0; q < f.pp.world; ++q) {
                if (cx < ncol4) st4(f.pp.data[q] + o, a);
                else f.pp.data[q][o] = a[0];
            }
            __threadfence_system();
        }
        __syncthreads();
        if (tid < f.pp.world) st_release_sys(f.pp.flags[tid] + (size_t)f.pp.rank * kPeerMaxBlocks, f.pp.epoch);
        if (tid == 0) dred[38] = 0.0;
        __syncthreads();
        if (tid < f.pp.world) {
            const unsigned* fl = f.pp.flags[f.pp.rank] + (size_t)tid * kPeerMaxBlocks;
            unsigned long long t0 = 0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while ((int)(ld_acquire_sys(fl) - f.pp.epoch) < 0) {   // epochs only grow (mod 2^32); a peer may be one call ahead
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 10000000000ull) {                    // 10 s: a peer never made the matching call
                    dred[38] = 1.0;                                // -> NaN outputs instead of a hung GPU
                    break;
                }
            }
        }
        __syncthreads();
        if (owner) {
            const T* mine = f.pp.data[f.pp.rank] + set_off + col_off;
            a[0] = a[1] = a[2] = a[3] = (T)0;
            for (int q = 0; q < f.pp.world; ++q) {
                if (cx < ncol4) {
                    T v[4];
                    ld4cg(mine + (size_t)q * f.pp.n_pad, v);
                    a[0] += v[0]; a[1] += v[1]; a[2] += v[2]; a[3] += v[3];
                } else {
                    a[0] += __ldcg(mine + (size_t)q * f.pp.n_pad);
                }
            }
            if (dred[38] != 0.0) a[0] = a[1] = a[2] = a[3] = (T)NAN;
        }
    }
    // ---- 3. per-target sums out; pooled mean fidelity (double, fixed order), loss, chain factor
    if (owner && cx >= ncol4) {
        f.Fsum[cx - ncol4] = a[0];
        fsm[cx - ncol4] = a[0];
    }
    T sc = (T)1;
    if (f.kind >= 0) {
        __syncthreads();
        double t = 0.0;
        for (int i = tid; i < B; i += nthr) t += (double)fsm[i];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if ((tid & 31) == 0) dred[tid >> 5] = t;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < (nthr + 31) / 32; ++w) tot += dred[w];
            double val, dval;
            const double Fbar = tot / f.n_total;
            su2_loss_eval(Fbar, f.kind, f.tau, f.k, val, dval);
            dred[37] = dval / f.n_total;
            if (f.loss_out != nullptr) {
                f.loss_out[0] = (T)val;
                f.loss_out[1] = (T)Fbar;
                f.loss_out[2] = (T)dval;
            }
        }
        __syncthreads();
        sc = (T)dred[37];
    }
#ifdef UQOC_FIN_TIMING
    if (tid == 0) stamp[3] = now();
#endif
    if (owner && cx < ncol4) {
        a[0] *= sc; a[1] *= sc; a[2] *= sc; a[3] *= sc;
        st4(f.G + 4 * (size_t)cx, a);
    }
#ifdef UQOC_FIN_TIMING
    __syncthreads();
    if (tid == 0) stamp[4] = now();
#endif
}
// bytes of shared scratch su2_block_finalize needs for a block of `nthr` threads, and the launch shapes it supports
__host__ __device__ inline size_t su2_fin_smem_bytes(int nthr, size_t elem) { return 40 * sizeof(double) + (size_t)nthr * 5 * elem; }
__host__ __device__ inline bool su2_fin_supported(long long B, long long L, int parts, int nthr, int po = 2) {
    const long long n_g = B * L * po, ncol = n_g / 4 + B;
    // <= 4 16-byte loads per thread: beyond that one block's pass over the partial rows (measured 11 us for the 148 rows
    // of BASELINE config 3) loses to a dependent-launched reduction kernel spread over many SMs (5.7 us)
    return (n_g % 4 == 0) && ncol <= nthr && (long long)parts * ncol <= 4LL * nthr;
}

// shared-memory footprint (bytes) of one block
template <typename T>
__host__ __device__ inline size_t su2_smem_bytes(int LPS, int C, bool bwd, int table_n = 0, bool fin = false) {
    const size_t rows = (size_t)LPS * (C + 1);
    size_t bytes = rows * sizeof(Row4<T>);                             // forward table
    bytes += 2 * (size_t)table_n * sizeof(T);                          // sin/cos table of the table policies
    if (bwd) bytes += rows * sizeof(Row4<T>);                          // backward table
    if (bwd) bytes += (size_t)kWarps * LPS * C * 2 * sizeof(T);        // per-warp gradient accumulators
    bytes += 32 * sizeof(T);                                           // block-reduction scratch
    const size_t fb = fin ? su2_fin_smem_bytes(kThreads, sizeof(T)) : 0;   // in-kernel epilogue (reuses the block's memory)
    return bytes > fb ? bytes : fb;
}

template <typename T, int ST, int LPS, int SC, bool BWD>
__global__ void __launch_bounds__(kThreads) su2_kernel(const Su2Params<T> p) {
    using R = Real<T>;
    using SCP = SinCos<T, SC>;
    constexpr bool kSlopeTable = (sizeof(T) == 8) && (SC == SC_TABLE);   // FP64 table policy: slope-indexed look-ups
    constexpr int NB = GradBuf<LPS>::NB;
    constexpr int SPB = kThreads / LPS;  // sample slots per block
    constexpr int TS = SPB * ST;         // samples per tile
    constexpr int NV = 2 * NB;

    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int C = p.C;
    const int rows = LPS * (C + 1);
    Row4<T>* fwd_tab = reinterpret_cast<Row4<T>*>(smem_raw);
    Row4<T>* bwd_tab = fwd_tab + (BWD ? rows : 0);
    T* acc = reinterpret_cast<T*>(bwd_tab + rows);       // [kWarps][LPS*C][2] (BWD only)
    T* scratch = acc + (BWD ? (size_t)kWarps * LPS * C * 2 : 0);
    T* sctab = scratch + 32;                             // {sin[N] | cos[N]} of the table sin/cos policies

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int k = tid % LPS;      // chunk index inside the sample's lane group
    const int slot = tid / LPS;   // sample slot inside the block
    const int split = blockIdx.x % p.splits;
    const int b = blockIdx.x / p.splits;
    const int L = p.L;

    grid_dependency_wait();       // launched as a programmatic dependent launch (su2_launch_one): no input is read above
    // ---- stage the target's pulse train: coalesced loads, trig in double, rounded once ----
    {
        for (int i = tid; i < LPS * C; i += kThreads) {
            const int ic = i < L ? i : L - 1;
            const int im = (i - 1) < 0 ? 0 : ((i - 1) < L ? (i - 1) : L - 1);
            T ph_c, ta_c, ph_m, ta_m;
            su2_pulse_at<T>(p, b, ic, ph_c, ta_c);
            su2_pulse_at<T>(p, b, im, ph_m, ta_m);
            if (p.head.pulses_out != nullptr && split == 0 && i < L) {
                p.head.pulses_out[((size_t)b * L + i) * 2] = ph_c;
                p.head.pulses_out[((size_t)b * L + i) * 2 + 1] = ta_c;
            }
            const double phi = (double)ph_c;
            const double phim = (double)ph_m;
            const T tau = i < L ? ta_c : (T)0;
            double sn, cs;
            ::sincos(phi, &sn, &cs);
            const int row = i + i / C;
            UQOC_ASSERT(row < LPS * (C + 1) && ic < L && im < L);
            fwd_tab[row] = Row4<T>{(T)cs, (T)sn, tau, (T)0};
            if (BWD) {
                double sd, cd;
                ::sincos(i == 0 ? 0.0 : phi - phim, &sd, &cd);
                bwd_tab[row] = Row4<T>{(T)cd, (T)sd, tau, (T)0};
            }
        }
        if (BWD) {
            for (int i = tid; i < kWarps * LPS * C * 2; i += kThreads) acc[i] = (T)0;
        }
        if constexpr (SCP::kTableN > 0) {
            for (int i = tid; i < SCP::kTableN; i += kThreads) {
                sctab[i] = (T)g_sin_table_f64[i];
                sctab[SCP::kTableN + i] = (T)g_cos_table_f64[i];
            }
        }
    }
    __syncthreads();

    // target coefficients: Tr(U^dagger T) = (cr + i ci) . P
    T cr[4], ci[4];
    su2_load_target<T>(p, b, cr, ci);

    const int rowbase = k * (C + 1);
    const size_t Bm = (size_t)p.B * p.M;
    T fsum = (T)0;

    for (int tile = split; tile < p.n_tiles; tile += p.splits) {
        SampleConst<T> kc[ST];
        bool valid[ST];
        size_t sidx[ST];
#pragma unroll
        for (int u = 0; u < ST; ++u) {
            const long long j = (long long)tile * TS + u * SPB + slot;
            valid[u] = j < p.M;
            sidx[u] = (size_t)b * p.M + (size_t)(valid[u] ? j : 0);
            T delta = (T)0, eps = (T)0;
            if (valid[u]) {
                su2_sample_errors<T>(p, b, j, sidx[u], Bm, delta, eps);
                if (p.err_out != nullptr && k == 0) {
                    p.err_out[sidx[u]] = delta;
                    p.err_out[Bm + sidx[u]] = eps;
                }
            }
            kc[u] = make_sample_const<T>(delta, eps);
        }

        // ---------------- forward sweep over this lane's chunk ----------------
        Quat<T> P[ST];
        int parity[ST];
#pragma unroll
        for (int u = 0; u < ST; ++u) {
            P[u] = Quat<T>{(T)1, (T)0, (T)0, (T)0};
            parity[u] = 0;
        }
#pragma unroll 2
        for (int jj = 0; jj < C; ++jj) {
            const Row4<T> row = fwd_tab[rowbase + jj];
#pragma unroll
            for (int u = 0; u < ST; ++u) {
                T s, c;
                int kb;
                if constexpr (kSlopeTable) {
                    SCP::eval_slope(row.z, kc[u].ap, s, c, kb, sctab);
                } else {
                    const T h = row.z * kc[u].a;
                    SCP::eval(h, s, c, kb, sctab);
                }
                if (SCP::kTracksParity && !BWD) parity[u] ^= kb;
                const T sp = s * kc[u].r;
                const T q1 = sp * row.x, q2 = sp * row.y, q3 = sp * kc[u].delta;
                const Quat<T> o = P[u];
                P[u].a = c * o.a - q1 * o.b - q2 * o.c - q3 * o.d;
                P[u].b = c * o.b + q1 * o.a + q2 * o.d - q3 * o.c;
                P[u].c = c * o.c - q1 * o.d + q2 * o.a + q3 * o.b;
                P[u].d = c * o.d + q1 * o.c - q2 * o.b + q3 * o.a;
            }
        }

        // ---------------- associative scan over the LPS lanes of a sample ----------------
        Quat<T> PL[ST];
#pragma unroll
        for (int u = 0; u < ST; ++u) {
            if constexpr (LPS > 1) {
#pragma unroll
                for (int d = 1; d < LPS; d <<= 1) {
                    const Quat<T> o = qshfl_up(P[u], d, LPS);
                    const Quat<T> n = qmul(P[u], o);   // later pulses on the left
                    if (k >= d) P[u] = n;
                }
                PL[u] = qshfl(P[u], LPS - 1, LPS);
                if (SCP::kTracksParity && !BWD) {
#pragma unroll
                    for (int d = 1; d < LPS; d <<= 1) parity[u] ^= __shfl_xor_sync(0xffffffffu, parity[u], d);
                }
            } else {
                PL[u] = P[u];
            }
            // unit-norm projection: removes the radial part of the accumulated rounding error
            const T n2 = PL[u].a * PL[u].a + PL[u].b * PL[u].b + PL[u].c * PL[u].c + PL[u].d * PL[u].d;
            const T inv = R::rsqrt_acc(n2);
            PL[u].a *= inv; PL[u].b *= inv; PL[u].c *= inv; PL[u].d *= inv;
        }

        // ---------------- fidelity epilogue ----------------
        T trr[ST], tri[ST];
#pragma unroll
        for (int u = 0; u < ST; ++u) {
            trr[u] = cr[0] * PL[u].a + cr[1] * PL[u].b + cr[2] * PL[u].c + cr[3] * PL[u].d;
            tri[u] = ci[0] * PL[u].a + ci[1] * PL[u].b + ci[2] * PL[u].c + ci[3] * PL[u].d;
            const T F = (trr[u] * trr[u] + tri[u] * tri[u] + (T)2) * (T)(1.0 / 6.0);
            if (valid[u] && k == 0) {
                fsum += F;
                if (p.F_out != nullptr) p.F_out[sidx[u]] = F;
                if (!BWD && p.U_out != nullptr) {
                    const T sg = (parity[u] & 1) ? (T)-1 : (T)1;
                    T* U = p.U_out + sidx[u] * 8;
                    // U = q0 I - i (q1 X + q2 Y + q3 Z), row-major interleaved (re, im)
                    U[0] = sg * PL[u].a;  U[1] = -sg * PL[u].d;
                    U[2] = -sg * PL[u].c; U[3] = -sg * PL[u].b;
                    U[4] = sg * PL[u].c;  U[5] = -sg * PL[u].b;
                    U[6] = sg * PL[u].a;  U[7] = sg * PL[u].d;
                }
            }
        }

        if constexpr (BWD) {
            // ---------------- adjoint seed at the end of this lane's chunk ----------------
            T A[ST], Bq[ST], W3[ST];
            {
                const Row4<T> rowL = fwd_tab[rowbase + C - 1];
#pragma unroll
                for (int u = 0; u < ST; ++u) {
                    T wgt = (T)0;
                    if (valid[u]) wgt = p.weight != nullptr ? p.weight[sidx[u]] : (T)1;
                    const T fr = wgt * trr[u] * (T)(1.0 / 3.0), fi = wgt * tri[u] * (T)(1.0 / 3.0);
                    const Quat<T> lam{fr * cr[0] + fi * ci[0], fr * cr[1] + fi * ci[1],
                                      fr * cr[2] + fi * ci[2], fr * cr[3] + fi * ci[3]};
                    Quat<T> Wq;
                    if constexpr (LPS > 1) {
                        const Quat<T> Lam = qmul(qconj(PL[u]), lam);
                        Wq = qmul(qmul(P[u], Lam), qconj(P[u]));
                    } else {
                        Wq = qmul(lam, qconj(PL[u]));
                    }
                    A[u] = Wq.b * rowL.x + Wq.c * rowL.y;
                    Bq[u] = Wq.c * rowL.x - Wq.b * rowL.y;
                    W3[u] = Wq.d;
                }
            }
            // ---------------- backward sweep ----------------
            for (int jb = C / NB - 1; jb >= 0; --jb) {
                T v[NV];
#pragma unroll
                for (int e = NB - 1; e >= 0; --e) {
                    const Row4<T> row = bwd_tab[rowbase + jb * NB + e];
                    T gp = (T)0, gt = (T)0;
#pragma unroll
                    for (int u = 0; u < ST; ++u) {
                        T C2, Sr;
                        if constexpr (kSlopeTable) {
                            T S2;                                     // (sin 2h, cos 2h) straight from the table
                            SCP::eval_slope_signed(row.z, kc[u].a2p, S2, C2, sctab);
                            Sr = S2 * kc[u].r;
                        } else {
                            const T h = row.z * kc[u].a;
                            T s, c;
                            int kb;
                            SCP::eval(h, s, c, kb, sctab);
                            const T s2 = s + s;
                            C2 = R::fma(-s2, s, (T)1);                // cos 2h
                            Sr = (s2 * kc[u].r) * c;                  // sin 2h / w
                        }
                        const T k1 = R::fma(-C2, kc[u].r2, kc[u].r2);  // (1 - cos 2h)/w^2
                        const T dl = kc[u].delta;
                        const T t = R::fma(dl, W3[u], A[u]);      // w <W, n>
                        const T uu = R::fma(dl, A[u], -W3[u]);
                        gt = R::fma(kc[u].ae, t, gt);
                        gp = R::fma(Sr, Bq[u], gp);
                        gp = R::fma(-k1, uu, gp);
                        const T K = k1 * t, BS = Bq[u] * Sr;
                        T A1 = R::fma(A[u], C2, K);
                        A1 = R::fma(dl, BS, A1);
                        T B1 = Bq[u] * C2;
                        B1 = R::fma(-uu, Sr, B1);
                        T Wz = R::fma(W3[u], C2, -BS);
                        W3[u] = R::fma(dl, K, Wz);
                        // into the frame of pulse i-1: rotate by (phi_i - phi_{i-1})
                        A[u] = R::fma(-B1, row.y, A1 * row.x);
                        Bq[u] = R::fma(B1, row.x, A1 * row.y);
                    }
                    v[2 * e] = gp;
                    v[2 * e + 1] = gt;
                }
                int base = 0;
                LaneReduce<T, NV, LPS>::run(v, lane, base);
                constexpr int NF = reduce_final_count(NV, LPS);
                constexpr int DUP = reduce_dup_mask(NV, LPS);
                if ((lane & DUP) == 0) {
                    T* dst = acc + ((size_t)(warp * LPS + k) * C + (size_t)jb * NB) * 2 + base;
                    UQOC_ASSERT(base >= 0 && ((size_t)(warp * LPS + k) * C + (size_t)jb * NB) * 2 + base < (size_t)kWarps * LPS * C * 2);
#pragma unroll
                    for (int m = 0; m < NF; ++m) dst[m] += v[m];
                }
            }
        }
    }

    // ---------------- block epilogue: deterministic fixed-order reductions ----------------
    {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
        if (lane == 0) scratch[warp] = fsum;
    }
    __syncthreads();
    if (tid == 0 && p.Fsum_part != nullptr) {
        T tot = (T)0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tot += scratch[w];
        p.Fsum_part[(size_t)split * p.B + b] = tot;
    }
    if constexpr (BWD) {
        const int LC2 = LPS * C * 2;
        if (p.head.mode == 0) {
            T* gout = p.G_part + ((size_t)split * p.B + b) * L * 2;
            for (int i = tid; i < 2 * L; i += kThreads) {
                T tot = (T)0;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) tot += acc[(size_t)w * LC2 + i];
                gout[i] = (i & 1) ? tot : tot * (T)0.5;   // d/dphi carries the 1/2 of sin 2h = 2 s c
            }
        } else {
            // head backward: one thread per pulse, (d/dphi, d/dtau) -> the head's input row
            const int po = su2_grad_width(p);
            T* gout = p.G_part + ((size_t)split * p.B + b) * L * po;
            for (int l = tid; l < L; l += kThreads) {
                T t0 = (T)0, t1 = (T)0;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) {
                    t0 += acc[(size_t)w * LC2 + 2 * l];
                    t1 += acc[(size_t)w * LC2 + 2 * l + 1];
                }
                T o[3];
                su2_head_grad<T>(p, b, l, t0 * (T)0.5, t1, o);
                for (int c = 0; c < po; ++c) gout[(size_t)po * l + c] = o[c];
            }
        }
        if (p.fin.ticket != nullptr) {
            __syncthreads();                          // the accumulators are dead: the epilogue reuses shared memory
            const FinParams<T> fin = p.fin;           // a copy: the kernel parameters themselves stay in the constant bank
            su2_block_finalize<T>(fin, p.G_part, p.Fsum_part, p.cps, p.B, p.L, smem_raw, su2_grad_width(p));
        }
    }
}

// ---- second stage: fixed-order sum of the per-split partials --------------------------------
// Block = 32 outputs x YL split-lanes: lane y sums splits y, y+YL, ... (coalesced 128 B rows, loads
// independent so they pipeline), then a fixed-order combine over y through shared memory.  Fsum rides
// along as outputs [n_g, n_g + B).  Deterministic: the summation order depends only on `splits`.
template <typename T, int YL>
__global__ void __launch_bounds__(32 * YL) su2_reduce_partials(const T* __restrict__ Fsum_part, const T* __restrict__ G_part,
                                                               int splits, int B, long long n_g /* B*L*P or 0 */,
                                                               T* __restrict__ Fsum, T* __restrict__ G) {
    __shared__ T red[YL][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * 32 + x;
    const long long n_f = (Fsum != nullptr) ? B : 0;
    grid_dependency_wait();
    T tot = (T)0;
    if (i < n_g) {
#pragma unroll 8
        for (int s = y; s < splits; s += YL) tot += G_part[(size_t)s * n_g + i];
    } else if (i < n_g + n_f) {
#pragma unroll 8
        for (int s = y; s < splits; s += YL) tot += Fsum_part[(size_t)s * B + (i - n_g)];
    }
    red[y][x] = tot;
    __syncthreads();
    if (y == 0) {
        T t = red[0][x];
#pragma unroll
        for (int yy = 1; yy < YL; ++yy) t += red[yy][x];
        if (i < n_g) G[i] = t;
        else if (i < n_g + n_f) Fsum[i - n_g] = t;
    }
}

// ---- fused second stage + loss epilogue (single GPU, no exchange step in between) -----------------
// Every block first recomputes Fbar = sum_{split,b} Fsum_part / n_total with the same fixed-order tree
// (bit-identical across blocks), evaluates the loss and d loss / d Fbar, then reduces its 32 outputs over
// the splits and scales them.  Saves one launch per step, which matters at BASELINE config 3 size.
template <typename T>
__global__ void __launch_bounds__(1024) su2_reduce_finalize(const T* __restrict__ Fsum_part, const T* __restrict__ G_part,
                                                            int splits, int B, long long n_g, double n_total, int kind,
                                                            double tau, double k, T* __restrict__ Fsum, T* __restrict__ G,
                                                            T* __restrict__ loss_out) {
    constexpr int YL = 32;
    __shared__ T red[YL][33];
    __shared__ double fred[1024];
    __shared__ double s_scale;
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    grid_dependency_wait();
    // Fbar: fixed-order strided partial sums + tree
    {
        double a = 0.0;
        const long long nf = (long long)splits * B;
        for (long long i = threadIdx.x; i < nf; i += 1024) a += (double)Fsum_part[i];
        fred[threadIdx.x] = a;
        __syncthreads();
        for (int d = 512; d >= 1; d >>= 1) {
            if ((int)threadIdx.x < d) fred[threadIdx.x] += fred[threadIdx.x + d];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            double val, dval;
            const double Fbar = fred[0] / n_total;
            su2_loss_eval(Fbar, kind, tau, k, val, dval);
            s_scale = dval / n_total;
            if (blockIdx.x == 0 && loss_out != nullptr) {
                loss_out[0] = (T)val;
                loss_out[1] = (T)Fbar;
                loss_out[2] = (T)dval;
            }
        }
        __syncthreads();
    }
    const long long i = (long long)blockIdx.x * 32 + x;
    const long long n_f = (Fsum != nullptr) ? B : 0;
    T tot = (T)0;
    if (i < n_g) {
#pragma unroll 8
        for (int s = y; s < splits; s += YL) tot += G_part[(size_t)s * n_g + i];
    } else if (i < n_g + n_f) {
#pragma unroll 8
        for (int s = y; s < splits; s += YL) tot += Fsum_part[(size_t)s * B + (i - n_g)];
    }
    red[y][x] = tot;
    __syncthreads();
    if (y == 0) {
        T t = red[0][x];
#pragma unroll
        for (int yy = 1; yy < YL; ++yy) t += red[yy][x];
        if (i < n_g) G[i] = t * (T)s_scale;
        else if (i < n_g + n_f) Fsum[i - n_g] = t;
    }
}

// ---- fused second stage + cross-GPU exchange over NVLink peer memory -------------------------------------
// Replaces  su2_reduce_partials -> ncclAllReduce(SUM)  of the multi-GPU step for small exchange vectors (the
// NCCL all-reduce of a 2 KB buffer costs ~45 us of latency at 8 GPUs; BASELINE config 3 is a 50 us kernel).
// Every rank maps every other rank's exchange buffer (CUDA VMM peer mapping; the host side gets the pointers
// from torch.distributed._symmetric_memory).  One-shot, push model, no grid-wide or cross-block sync:
//   block g reduces its outputs over the sample-tile partials (fixed order), STORES them into slot[rank] of
//   every rank's buffer (coalesced 128 B rows over NVLink), fences at system scope, raises flag[rank][g] on every
//   rank, spins on its own flag[q][g] for all q, then sums the slots in rank order 0..R-1 -- the same order on
//   every rank, so all ranks end with bit-identical [G | Fsum].
// Two slot sets alternate with the epoch (a rank can only be one call ahead of the slowest peer).  The grid is
// capped at the resident capacity so every waiting block has its remote partner running.
template <typename T, int YL>
__global__ void __launch_bounds__(32 * YL) su2_reduce_exchange(const T* __restrict__ Fsum_part, const T* __restrict__ G_part,
                                                               int splits, int B, long long n_g, const PeerParams<T> pp,
                                                               T* __restrict__ Fsum, T* __restrict__ G) {
    static_assert(YL >= kPeerMaxWorld, "one warp per peer in the push / pull phases");
    __shared__ T red[YL][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const long long n = n_g + B;
    const long long n_groups = (n + 31) / 32;
    const size_t set_off = (size_t)(pp.epoch & 1u) * pp.world * pp.n_pad;
    grid_dependency_wait();
    // ---- phase 1: reduce over the sample-tile partials, push to every rank's slot[rank]
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long i = g * 32 + x;
        T tot = (T)0;
        if (i < n_g) {
#pragma unroll 8
            for (int s = y; s < splits; s += YL) tot += G_part[(size_t)s * n_g + i];
        } else if (i < n) {
#pragma unroll 8
            for (int s = y; s < splits; s += YL) tot += Fsum_part[(size_t)s * B + (i - n_g)];
        }
        red[y][x] = tot;
        __syncthreads();
        if (y == 0) {
            T t = red[0][x];
#pragma unroll
            for (int yy = 1; yy < YL; ++yy) t += red[yy][x];
            red[0][x] = t;
        }
        __syncthreads();
        if (y < pp.world) pp.data[y][set_off + (size_t)pp.rank * pp.n_pad + i] = red[0][x];   // i < n_pad always
        __syncthreads();
    }
    if (y < pp.world) __threadfence_system();       // only the warps that stored to peers
    __syncthreads();
    if ((int)threadIdx.x < pp.world)
        st_release_sys(pp.flags[threadIdx.x] + (size_t)pp.rank * kPeerMaxBlocks + blockIdx.x, pp.epoch);
    // ---- phase 2: wait for block blockIdx.x of every rank, then sum the slots in rank order
    __shared__ int s_timeout;
    if (threadIdx.x == 0) s_timeout = 0;
    __syncthreads();
    if ((int)threadIdx.x < pp.world) {
        const unsigned* f = pp.flags[pp.rank] + (size_t)threadIdx.x * kPeerMaxBlocks + blockIdx.x;
        unsigned long long t0 = 0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(ld_acquire_sys(f) - pp.epoch) < 0) {      // epochs only grow; a peer may already be one call ahead
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 10000000000ull) {                    // 10 s: a peer never made the matching call
                s_timeout = 1;                                 // -> NaN outputs instead of a hung GPU
                break;
            }
        }
    }
    __syncthreads();
    const bool timed_out = s_timeout != 0;
    const T* mine = pp.data[pp.rank] + set_off;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long i = g * 32 + x;
        if (y < pp.world) red[y][x] = __ldcg(mine + (size_t)y * pp.n_pad + i);
        __syncthreads();
        if (y == 0) {
            T t = red[0][x];
            for (int q = 1; q < pp.world; ++q) t += red[q][x];
            if (timed_out) t = (T)NAN;
            if (i < n_g) G[i] = t;
            else if (i < n) Fsum[i - n_g] = t;
        }
        __syncthreads();
    }
}

// ---- the same exchange with the data and its "arrived" mark in ONE 8-byte store (FP32) ---------------------------
// su2_reduce_exchange pays two serialised NVLink transactions per call: the slot stores have to be fenced at system
// scope before the flag may follow.  Here every exchanged real travels as {value, epoch} in one aligned 64-bit store
// (delivered atomically), the receiver polls the word itself until it carries this call's epoch: one one-way NVLink
// latency, no fence, no flag buffer.  The loss epilogue rides along: every block forms the pooled mean fidelity from
// the B x world exchanged Fsum words in the same fixed order (bit-identical in every block and on every rank),
// evaluates the loss and scales its gradient columns -- the multi-GPU step is main kernel + this one dependent launch.
// Slot layout per rank buffer: u64 [2 sets][world][n_pad]; sets alternate with the epoch's parity (a rank can be at
// most one call ahead of the slowest peer, see su2_reduce_exchange).  kind < 0: reduction + exchange only.
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// poll one {value, epoch} word; false after 10 s (a peer never made the matching call)
__device__ __forceinline__ bool ll_wait(const unsigned long long* p, unsigned epoch, float& val) {
    unsigned long long w = ld_relaxed_sys_u64(p);
    if ((int)((unsigned)(w >> 32) - epoch) < 0) {
        // tight poll; the wall clock (a slow system-level read) is consulted every 4096 polls only
        unsigned long long t0 = 0, t1;
        unsigned spins = 0;
        do {
            w = ld_relaxed_sys_u64(p);
            if ((++spins & 4095u) == 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t0 == 0) t0 = t1;
                else if (t1 - t0 > 10000000000ull) return false;
            }
        } while ((int)((unsigned)(w >> 32) - epoch) < 0);
    }
    val = __uint_as_float((unsigned)w);
    return true;
}

#ifdef UQOC_LL_TIMING
// timing build (tools/ll_timing.py): globaltimer stamps in the workspace's ticket area, 128 bytes below the partial rows
#define UQOC_LL_STAMP(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
    (reinterpret_cast<unsigned long long*>(const_cast<float*>(G_part)) - 16)[k] = t_; } } while (0)
#else
#define UQOC_LL_STAMP(k) ((void)0)
#endif
template <int YL>   // a template only so that the header can be included by several translation units
__global__ void __launch_bounds__(32 * YL) su2_reduce_exchange_ll(const float* __restrict__ Fsum_part, const float* __restrict__ G_part,
                                                               int splits, int B, long long n_g, const PeerParams<float> pp,
                                                               double n_total, int kind, double tau, double k,
                                                               float* __restrict__ Fsum, float* __restrict__ G,
                                                               float* __restrict__ loss_out) {
    static_assert(YL == 32, "1024-thread blocks: one warp per part-lane / per peer");
    __shared__ float red[YL][33];
    __shared__ double fred[32];
    __shared__ float s_scale;
    __shared__ int s_timeout;
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const long long n = n_g + B;
    const long long n_groups = (n + 31) / 32;
    const size_t set_off = (size_t)(pp.epoch & 1u) * pp.world * pp.n_pad;
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(pp.data[pp.rank]) + set_off;
    if (threadIdx.x == 0) s_timeout = 0;
    UQOC_LL_STAMP(0);
    grid_dependency_wait();
    UQOC_LL_STAMP(1);
    // ---- phase 1: reduce over the sample-tile partials, push {value, epoch} to every rank's slot[rank]
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long i = g * 32 + x;
        float tot = 0.0f;
        if (i < n_g) {
#pragma unroll 8
            for (int s = y; s < splits; s += YL) tot += G_part[(size_t)s * n_g + i];
        } else if (i < n) {
#pragma unroll 8
            for (int s = y; s < splits; s += YL) tot += Fsum_part[(size_t)s * B + (i - n_g)];
        }
        red[y][x] = tot;
        __syncthreads();
        if (y == 0) {
            float t = red[0][x];
#pragma unroll
            for (int yy = 1; yy < YL; ++yy) t += red[yy][x];
            red[0][x] = t;
        }
        __syncthreads();
        if (y < pp.world) {
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(pp.data[y]) + set_off + (size_t)pp.rank * pp.n_pad + i;
            st_relaxed_sys_u64(dst, ((unsigned long long)pp.epoch << 32) | (unsigned long long)__float_as_uint(red[0][x]));
        }
        __syncthreads();
    }
    // ---- phase 2: wait for every rank's words of this block's outputs, sum the slots in rank order.  The first group's
    // totals stay in registers (the usual case: one group per block); further groups park theirs, unscaled, in G.
    UQOC_LL_STAMP(2);
    float first_tot = 0.0f;
    __shared__ float fs[1024];
    // few targets (the usual case here): the warps that poll no column fetch the B x world fidelity-sum words of phase 3
    // at the same time, so the loss epilogue does not add a second wait
    const bool fs_prefetch = kind >= 0 && B * pp.world <= 1024 - 32 * pp.world;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long i = g * 32 + x;
        if (y < pp.world) {
            float v = 0.0f;
            if (!ll_wait(mine + (size_t)y * pp.n_pad + i, pp.epoch, v)) s_timeout = 1;
            red[y][x] = v;
        } else if (fs_prefetch && g == blockIdx.x) {
            const int idx = (int)threadIdx.x - 32 * pp.world;          // = b * world + q
            if (idx < B * pp.world) {
                float v = 0.0f;
                if (!ll_wait(mine + (size_t)(idx % pp.world) * pp.n_pad + n_g + idx / pp.world, pp.epoch, v)) s_timeout = 1;
                fs[idx] = v;
            }
        }
        __syncthreads();
        if (y == 0) {
            float t = red[0][x];
            for (int q = 1; q < pp.world; ++q) t += red[q][x];
            if (s_timeout) t = NAN;
            if (g == blockIdx.x) first_tot = t;
            if (i >= n_g && i < n) Fsum[i - n_g] = t;
            else if (i < n_g && (kind < 0 || g != blockIdx.x)) G[i] = t;
        }
        __syncthreads();
    }
    UQOC_LL_STAMP(3);
    if (kind < 0) return;
    // ---- phase 3: pooled mean fidelity from ALL exchanged Fsum words (other blocks' columns included: polled here too,
    // so no cross-block synchronisation).  One thread per (target, rank) word -- the polls of a pass are in flight
    // together -- then a fixed-order sum: the same bits in every block and on every rank.
    {
        const int tpp = 1024 / pp.world;                       // targets per pass
        double a = 0.0;
        for (int b0 = 0; b0 < B; b0 += tpp) {
            const int bl = (int)threadIdx.x / pp.world, q = (int)threadIdx.x % pp.world;
            if (!fs_prefetch) {
                if (bl < tpp && b0 + bl < B) {
                    float v = 0.0f;
                    if (!ll_wait(mine + (size_t)q * pp.n_pad + n_g + b0 + bl, pp.epoch, v)) s_timeout = 1;
                    fs[threadIdx.x] = v;
                }
                __syncthreads();
            }
            if ((int)threadIdx.x < tpp && b0 + (int)threadIdx.x < B) {
                float t = fs[threadIdx.x * pp.world];
                for (int q2 = 1; q2 < pp.world; ++q2) t += fs[threadIdx.x * pp.world + q2];
                a += (double)t;
            }
            __syncthreads();
        }
        // only the first tpp threads hold a non-zero partial: when they all sit in warp 0 (B <= 32) one butterfly does it
        if (B > 32) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
            if (x == 0) fred[y] = a;
            __syncthreads();
            a = (y == 0) ? fred[x] : 0.0;
        }
        if (y == 0) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
            if (x == 0) {
                // double, like the single-GPU epilogues: the sharp loss amplifies a rounding of Fbar by k = 100, and the
                // FP32 alternative saves only 0.35 us
                double val, dval;
                const double Fbar = a / n_total;
                su2_loss_eval(Fbar, kind, tau, k, val, dval);
                s_scale = s_timeout ? NAN : (float)(dval / n_total);
                if (blockIdx.x == 0 && loss_out != nullptr) {
                    loss_out[0] = s_timeout ? NAN : (float)val;
                    loss_out[1] = (float)Fbar;
                    loss_out[2] = (float)dval;
                }
            }
        }
        __syncthreads();
    }
    UQOC_LL_STAMP(4);
    if (y == 0) {
        const float sc = s_scale;
        for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
            const long long i = g * 32 + x;
            if (i < n_g) G[i] = (g == blockIdx.x ? first_tot : G[i]) * sc;
        }
    }
    UQOC_LL_STAMP(5);
}

template <typename T>
inline void launch_reduce_partials(const T* Fsum_part, const T* G_part, int splits, int B, long long n_g, T* Fsum, T* G,
                                   cudaStream_t stream) {
    const long long n = n_g + (Fsum != nullptr ? B : 0);
    const unsigned blocks = (unsigned)((n + 31) / 32);
    if (splits > 64) launch_dependent(su2_reduce_partials<T, 32>, blocks, 1024, 0, stream, Fsum_part, G_part, splits, B, n_g, Fsum, G);
    else launch_dependent(su2_reduce_partials<T, 8>, blocks, 256, 0, stream, Fsum_part, G_part, splits, B, n_g, Fsum, G);
}

// =============================================================================================
// Strict reference signature: one pulse row per sample (SCORE.py:77-145) and its backward.
// One thread per sample; pulse rows are walked through L1 (each 128 B line serves 16 steps).
// =============================================================================================
template <typename T, int SC>
__global__ void __launch_bounds__(128) su2_generator_fwd_kernel(const T* __restrict__ pulses, const T* __restrict__ err,
                                                                long long Bm, int L, T* __restrict__ U_out) {
    using SCP = SinCos<T, SC_LIBM>;
    const long long s_idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s_idx >= Bm) return;
    const SampleConst<T> kc = make_sample_const<T>(err[s_idx], err[Bm + s_idx]);
    const T* row = pulses + (size_t)s_idx * L * 2;
    Quat<T> P{(T)1, (T)0, (T)0, (T)0};
    for (int i = 0; i < L; ++i) {
        const T phi = row[2 * i], tau = row[2 * i + 1];
        T sp_, cp_, s, c;
        int kb;
        SCP::eval(phi, sp_, cp_, kb);
        SCP::eval(tau * kc.a, s, c, kb);
        const T sp = s * kc.r;
        const Quat<T> q{c, sp * cp_, sp * sp_, sp * kc.delta};
        P = qmul(q, P);
    }
    const T n2 = P.a * P.a + P.b * P.b + P.c * P.c + P.d * P.d;
    const T inv = Real<T>::rsqrt_acc(n2);
    P.a *= inv; P.b *= inv; P.c *= inv; P.d *= inv;
    T* U = U_out + (size_t)s_idx * 8;
    U[0] = P.a;  U[1] = -P.d;
    U[2] = -P.c; U[3] = -P.b;
    U[4] = P.c;  U[5] = -P.b;
    U[6] = P.a;  U[7] = P.d;
}

template <typename T, int SC>
__global__ void __launch_bounds__(128) su2_generator_bwd_kernel(const T* __restrict__ pulses, const T* __restrict__ err,
                                                                const T* __restrict__ grad_U, long long Bm, int L,
                                                                T* __restrict__ grad_pulses) {
    using SCP = SinCos<T, SC_LIBM>;
    using R = Real<T>;
    const long long s_idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s_idx >= Bm) return;
    const SampleConst<T> kc = make_sample_const<T>(err[s_idx], err[Bm + s_idx]);
    const T* row = pulses + (size_t)s_idx * L * 2;
    T* grow = grad_pulses + (size_t)s_idx * L * 2;
    Quat<T> P{(T)1, (T)0, (T)0, (T)0};
    for (int i = 0; i < L; ++i) {
        const T phi = row[2 * i], tau = row[2 * i + 1];
        T sp_, cp_, s, c;
        int kb;
        SCP::eval(phi, sp_, cp_, kb);
        SCP::eval(tau * kc.a, s, c, kb);
        const T sp = s * kc.r;
        const Quat<T> q{c, sp * cp_, sp * sp_, sp * kc.delta};
        P = qmul(q, P);
    }
    // cotangent on U (torch convention dL = Re sum conj(g) dU) -> cotangent on the quaternion
    const T* g = grad_U + (size_t)s_idx * 8;
    const Quat<T> lam{g[0] + g[6], -(g[3] + g[5]), g[4] - g[2], g[7] - g[1]};
    const Quat<T> Wq = qmul(lam, qconj(P));
    T W1 = Wq.b, W2 = Wq.c, W3 = Wq.d;
    for (int i = L - 1; i >= 0; --i) {
        const T phi = row[2 * i], tau = row[2 * i + 1];
        T sphi, cphi, s, c;
        int kb;
        SCP::eval(phi, sphi, cphi, kb);
        SCP::eval(tau * kc.a, s, c, kb);
        const T A = W1 * cphi + W2 * sphi;
        const T Bq = W2 * cphi - W1 * sphi;
        const T s2 = s + s;
        const T C2 = R::fma(-s2, s, (T)1);
        const T Sr = (s2 * kc.r) * c;
        const T k1 = R::fma(-C2, kc.r2, kc.r2);
        const T dl = kc.delta;
        const T t = R::fma(dl, W3, A);
        const T uu = R::fma(dl, A, -W3);
        grow[2 * i] = (T)0.5 * (Sr * Bq - k1 * uu);
        grow[2 * i + 1] = kc.ae * t;
        const T K = k1 * t, BS = Bq * Sr;
        const T A1 = R::fma(dl, BS, R::fma(A, C2, K));
        const T B1 = R::fma(-uu, Sr, Bq * C2);
        W3 = R::fma(dl, K, R::fma(W3, C2, -BS));
        W1 = A1 * cphi - B1 * sphi;
        W2 = A1 * sphi + B1 * cphi;
    }
}

// ---- host-side launch plan shared by all instantiation units ---------------------------------
struct Su2Plan {
    int st, lps, splits, n_tiles, C;
    size_t smem;
    bool packed;   // FP32 only: f32x2 (FFMA2) kernel, two samples per register pair
    int wps;       // packed kernel: warps per sample group (1, or 4 = pulse train split over the block's warps)
    bool table;    // packed kernel: table-lookup sin/cos instead of the polynomial pair
    int vb;        // packed kernel: virtual blocks per block (kX2FatVB = fat block, one per SM), else 1
    int cps;       // blocks per target = splits / vb (partial rows per target)
    bool fin;      // in-kernel epilogue (su2_block_finalize) instead of separate reduction / loss launches
};

}  // namespace uqoc
