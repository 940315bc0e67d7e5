// Packed-FP32 (f32x2 -> SASS FFMA2/FMUL2/FADD2, new on sm_100) variant of the fused SU(2) kernel.
//
// Two error samples ride in the two halves of every 64-bit register pair for the per-pulse scalars
// (table sin/cos, q) and for the whole backward sweep, so one instruction advances two propagations;
// the forward running product keeps two COMPONENTS of one sample per pair instead (per-sample scalar
// in the 32-bit broadcast slot, swap / sign operand modifiers).  The FP32 FMA pipe does the same FLOPs
// either way, but the packed form needs half the issue slots -- the scalar kernel is issue/dispatch
// limited (ncu: 81 % issue-active), this one is limited by the FMA pipe and its operand delivery
// (74 % pipe-active).  Same algorithm as su2_kernel (uqoc_su2_kernels.cuh), one thread
// owns its samples' whole pulse train (WPS = 1) or a quarter of it (WPS = 4).
// sin/cos: shared-memory table indexed by the per-sample slope (sincos2_tab), half angle in the forward
// sweep, full angle (full-period table) in the backward sweep.
#pragma once
#include "uqoc_su2_kernels.cuh"

namespace uqoc {

constexpr int kTabN = UQOC_SINCOS_TABLE_N;

typedef unsigned long long u64;

struct F2 {
    u64 v;
};

__device__ __forceinline__ F2 f2(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 f2b(float s) { return f2(s, s); }
__device__ __forceinline__ float f2lo(F2 a) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
    return lo;
}
__device__ __forceinline__ float f2hi(F2 a) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
    return hi;
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    F2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return d;
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
    F2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
    F2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
__device__ __forceinline__ F2 sub2(F2 a, F2 b) {
    F2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
// ptxas folds these into the consumer's operand modifiers (no instruction is emitted): an FFMA2 / FMUL2 source takes
// a whole-pair negate, a half swap (.LO_HI) and -- in the FIRST source slot only -- a per-half sign pattern (.NP)
// for free.  So the sign-pattern forms must be passed as the first operand of fma2 / mul2; in the second slot
// ptxas materialises them (FADD + MOV).
__device__ __forceinline__ F2 neg2(F2 a) { return f2(-f2lo(a), -f2hi(a)); }
__device__ __forceinline__ F2 swp2(F2 a) { return f2(f2hi(a), f2lo(a)); }        // (hi, lo)
__device__ __forceinline__ F2 swp2_np(F2 a) { return f2(-f2hi(a), f2lo(a)); }    // (-hi, lo)
__device__ __forceinline__ F2 swp2_nn(F2 a) { return f2(-f2hi(a), -f2lo(a)); }   // (-hi, -lo)
__device__ __forceinline__ F2 sgn2_np(F2 a) { return f2(-f2lo(a), f2hi(a)); }    // (-lo, hi)
__device__ __forceinline__ F2 sgn2_pn(F2 a) { return f2(f2lo(a), -f2hi(a)); }    // (lo, -hi)

// packed twin of sincos_modpi (uqoc_common.cuh): (s, c) = (-1)^k (sin h, cos h), k in kb_*.
__device__ __forceinline__ void sincos_modpi2(F2 h, F2& s, F2& c, int& kb_lo, int& kb_hi) {
    const float MAGIC = 12582912.0f;
    F2 kf = fma2(h, f2b(0.318309886183790672f), f2b(MAGIC));
    kb_lo = __float_as_int(f2lo(kf));
    kb_hi = __float_as_int(f2hi(kf));
    kf = add2(kf, f2b(-MAGIC));
    F2 r = fma2(kf, f2b(-3.14159274101257324f), h);
    r = fma2(kf, f2b(8.74227765734758577e-08f), r);
    const F2 z = mul2(r, r);
    F2 ps = fma2(z, f2b(2.6325158160034334e-06f), f2b(-1.9822049944195896e-04f));
    ps = fma2(z, ps, f2b(8.3332369104027748e-03f));
    ps = fma2(z, ps, f2b(-1.6666665673255920e-01f));
    const F2 rz = mul2(r, z);
    s = fma2(rz, ps, r);
    F2 pc = fma2(z, f2b(-2.6282461362825416e-07f), f2b(2.4774040866759606e-05f));
    pc = fma2(z, pc, f2b(-1.3888647081330419e-03f));
    pc = fma2(z, pc, f2b(4.1666660457849503e-02f));
    pc = fma2(z, pc, f2b(-0.5f));
    c = fma2(z, pc, f2b(1.0f));
}

template <int SC>
__device__ __forceinline__ void sincos2(F2 h, F2& s, F2& c, int& kb_lo, int& kb_hi) {
    if constexpr (SC == SC_MUFU) {
        float s0, c0, s1, c1;
        __sincosf(f2lo(h), &s0, &c0);
        __sincosf(f2hi(h), &s1, &c1);
        s = f2(s0, s1);
        c = f2(c0, c1);
        kb_lo = kb_hi = 0;
    } else {
        sincos_modpi2(h, s, c, kb_lo, kb_hi);
    }
}

// Stage-major evaluation for NP independent pairs: consecutive instructions are independent
// (NP pairs x {sin, cos} Horner chains), which is what keeps the 2-cycle FFMA2 pipe fed from a
// single warp instead of relying on other warps to hide each dependent-issue latency.
// SC_TABLE: h = k pi/N + r with |r| <= pi/2N (N = 1024).  (sin, cos)(k pi/N) come from a shared-memory
// table (one LDS.32 per half), the residual is applied to FIRST order, (s, c) = (st + r ct, ct - r st):
// the angle error is r^3/3 <= 1.2e-9; the norm error r^2/2 <= 1.2e-6 is radial (removed exactly by the
// final re-normalisation of P_L) and has zero mean because the table is pre-scaled by 1 - (pi/2N)^2/6.
// 6 FMA-pipe instructions (4 with immediates) instead of 14; the sign (-1)^(k div N) is dropped like in
// the polynomial path (kb returns k >> 10 so the U_out kernel can track it).
// FULL: index the full-period table (k mod 2N): the signs of (s, c) are then exact and kb = 0 (the backward
// sweep looks up the DOUBLE angle 2h this way: cos 2h / sin 2h are what the adjoint rotation needs, and taking
// them from the table instead of the double-angle identities saves 3 FMA-pipe instructions per pair-step).
template <int NP, int SC, bool FULL = false>
__device__ __forceinline__ void sincos2_n(const F2 (&h)[NP], F2 (&s)[NP], F2 (&c)[NP], int (&kb)[2 * NP],
                                          const float* __restrict__ tsin, const float* __restrict__ tcos) {
    if constexpr (SC == SC_TABLE) {
        constexpr int MASK = FULL ? (2 * kTabN - 1) : (kTabN - 1);
        const float MAGIC = 12582912.0f;
        F2 kf[NP], r[NP], st[NP], ct[NP];
#pragma unroll
        for (int u = 0; u < NP; ++u) kf[u] = fma2(h[u], f2b(325.94931f), f2b(MAGIC));
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const int k0 = __float_as_int(f2lo(kf[u])), k1 = __float_as_int(f2hi(kf[u]));
            kb[2 * u] = FULL ? 0 : (k0 >> 10);
            kb[2 * u + 1] = FULL ? 0 : (k1 >> 10);
            const int i0 = k0 & MASK, i1 = k1 & MASK;
            st[u] = f2(tsin[i0], tsin[i1]);
            ct[u] = f2(tcos[i0], tcos[i1]);
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) kf[u] = add2(kf[u], f2b(-MAGIC));
#pragma unroll
        for (int u = 0; u < NP; ++u) r[u] = fma2(kf[u], f2b(-0.003067961661145091f), h[u]);
#pragma unroll
        for (int u = 0; u < NP; ++u) r[u] = fma2(kf[u], f2b(8.537380524753502e-11f), r[u]);
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            s[u] = fma2(r[u], ct[u], st[u]);
            c[u] = fma2(neg2(r[u]), st[u], ct[u]);
        }
    } else if constexpr (SC == SC_MUFU) {
#pragma unroll
        for (int u = 0; u < NP; ++u) sincos2<SC>(h[u], s[u], c[u], kb[2 * u], kb[2 * u + 1]);
    } else {
        const float MAGIC = 12582912.0f;
        F2 kf[NP], r[NP], z[NP], rz[NP], ps[NP], pc[NP];
#pragma unroll
        for (int u = 0; u < NP; ++u) kf[u] = fma2(h[u], f2b(0.318309886183790672f), f2b(MAGIC));
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            kb[2 * u] = __float_as_int(f2lo(kf[u]));
            kb[2 * u + 1] = __float_as_int(f2hi(kf[u]));
            kf[u] = add2(kf[u], f2b(-MAGIC));
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) r[u] = fma2(kf[u], f2b(-3.14159274101257324f), h[u]);
#pragma unroll
        for (int u = 0; u < NP; ++u) r[u] = fma2(kf[u], f2b(8.74227765734758577e-08f), r[u]);
#pragma unroll
        for (int u = 0; u < NP; ++u) z[u] = mul2(r[u], r[u]);
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            ps[u] = fma2(z[u], f2b(2.6325158160034334e-06f), f2b(-1.9822049944195896e-04f));
            pc[u] = fma2(z[u], f2b(-2.6282461362825416e-07f), f2b(2.4774040866759606e-05f));
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            ps[u] = fma2(z[u], ps[u], f2b(8.3332369104027748e-03f));
            pc[u] = fma2(z[u], pc[u], f2b(-1.3888647081330419e-03f));
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            rz[u] = mul2(r[u], z[u]);
            ps[u] = fma2(z[u], ps[u], f2b(-1.6666665673255920e-01f));
            pc[u] = fma2(z[u], pc[u], f2b(4.1666660457849503e-02f));
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            s[u] = fma2(rz[u], ps[u], r[u]);
            pc[u] = fma2(z[u], pc[u], f2b(-0.5f));
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) c[u] = fma2(z[u], pc[u], f2b(1.0f));
    }
}

// Table sin/cos of tau * a for NP pairs, from the per-sample table-index slope kap = a * N/pi:
//   kf = fma(tau, kap, MAGIC)  (k in the low mantissa bits),  frac = fma(tau, kap, -(kf - MAGIC))  (one rounding),
//   r = frac * pi/N,  (s, c) = (st + r ct, ct - r st).
// 6 FMA-pipe instructions; the angle tau*a itself is never formed (the Cody-Waite pair of sincos2_n's table
// path is replaced by the exact FMA residual).  FULL as in sincos2_n.
template <int NP, bool FULL>
__device__ __forceinline__ void sincos2_tab(F2 tau, const F2 (&kap)[NP], F2 (&s)[NP], F2 (&c)[NP], int (&kb)[2 * NP],
                                            const float* __restrict__ tsin, const float* __restrict__ tcos) {
    constexpr int MASK = FULL ? (2 * kTabN - 1) : (kTabN - 1);
    const float MAGIC = 12582912.0f;
    F2 kf[NP], fr[NP], st[NP], ct[NP];
#pragma unroll
    for (int u = 0; u < NP; ++u) kf[u] = fma2(tau, kap[u], f2b(MAGIC));
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        const int k0 = __float_as_int(f2lo(kf[u])), k1 = __float_as_int(f2hi(kf[u]));
        kb[2 * u] = FULL ? 0 : (k0 >> 10);
        kb[2 * u + 1] = FULL ? 0 : (k1 >> 10);
        const int i0 = k0 & MASK, i1 = k1 & MASK;
        st[u] = f2(tsin[i0], tsin[i1]);
        ct[u] = f2(tcos[i0], tcos[i1]);
    }
#pragma unroll
    for (int u = 0; u < NP; ++u) kf[u] = sub2(f2b(MAGIC), kf[u]);
#pragma unroll
    for (int u = 0; u < NP; ++u) fr[u] = fma2(tau, kap[u], kf[u]);
#pragma unroll
    for (int u = 0; u < NP; ++u) fr[u] = mul2(fr[u], f2b(0.0030679615757712823f));
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        s[u] = fma2(fr[u], ct[u], st[u]);
        c[u] = fma2(neg2(fr[u]), st[u], ct[u]);
    }
}

// shared memory of one block of the packed kernel (bytes).  C = pulses per chunk, WPS chunks.
__host__ __device__ inline size_t su2_x2_smem_bytes(int C, int wps, int st, bool bwd, bool table) {
    const size_t rows = (size_t)C * wps;
    size_t bytes = rows * 16;                                     // {cos phi, sin phi, tau, -} rows
    if (bwd) bytes += rows * 16;                                  // {cos dphi, sin dphi, tau, -} rows
    if (bwd) bytes += (size_t)kWarps * C * 2 * sizeof(float);     // gradient accumulators (per warp / per chunk)
    bytes += 32 * sizeof(float);
    if (table) bytes += 2 * (size_t)(bwd ? UQOC_SINCOS_TABLE_LEN : UQOC_SINCOS_TABLE_N) * sizeof(float);   // bwd: full period
    if (wps > 1) bytes += (size_t)kWarps * st * 5 * 32 * sizeof(float);   // chunk-product (+ parity) exchange
    return bytes;
}

// resident-block hint of the table kernel: 5 blocks/SM (<= 96 registers) measured best (8.33 ms vs 8.50 at 6, 8.37 at 7,
// 8.58 at 4 on the 4096 x 4096 x 256 launch); tools/variants.sh builds such variants side by side
#ifndef UQOC_X2_MINB
#define UQOC_X2_MINB 5
#endif
#ifndef UQOC_X2_FWD_UNROLL
#define UQOC_X2_FWD_UNROLL 2
#endif
constexpr int kX2FwdUnroll = UQOC_X2_FWD_UNROLL;
constexpr int x2_min_blocks(int NP, int SC, int WPS) { return (WPS == 4 && NP == 1) ? 7 : ((SC == SC_TABLE) ? UQOC_X2_MINB : 1); }

// NP  = sample PAIRS per thread (1 or 2)
// WPS = warps per sample group.  1: every warp owns its own 32*ST samples and the whole pulse train.
//       4: the block's four warps share 32*ST samples and each owns a quarter of the pulse train
//       (chunk products exchanged through shared memory; the prefix alone seeds every chunk's backward
//       sweep, see uqoc_su2_kernels.cuh) -- 4x the warps for the same samples, for small sample counts.
template <int NP, int SC, bool BWD, int WPS>
__global__ void __launch_bounds__(kThreads, x2_min_blocks(NP, SC, WPS)) su2_kernel_x2(const Su2Params<float> p) {
    constexpr int ST = 2 * NP;
    constexpr int NB = 8;
    constexpr int SLOTS = (WPS == 1) ? kThreads : 32;     // sample slots per block
    constexpr int TS = SLOTS * ST;
    constexpr int NV = 2 * NB;

    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int C = p.C;                 // pulses per chunk (multiple of NB)
    const int CT = C * WPS;            // staged rows
    float4* fwd4 = reinterpret_cast<float4*>(smem_raw);
    float4* bwd4 = fwd4 + CT;
    float* acc = reinterpret_cast<float*>(bwd4 + (BWD ? CT : 0));
    float* scratch = acc + (BWD ? (size_t)kWarps * C * 2 : 0);
    constexpr int TLEN = (SC == SC_TABLE) ? (BWD ? UQOC_SINCOS_TABLE_LEN : kTabN) : 0;   // table entries staged
    float* tsin = scratch + 32;
    float* tcos = tsin + TLEN;
    float* xq = tcos + TLEN;                                // [kWarps][ST][5][32], WPS > 1 only

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int split = blockIdx.x % p.splits;
    const int b = blockIdx.x / p.splits;
    const int L = p.L;
    const int rb = (WPS == 1) ? 0 : warp * C;              // first staged row of this warp's chunk
    const int slot = (WPS == 1) ? tid : lane;
    const bool lead = (WPS == 1) || warp == 0;             // the warp that reports per-sample outputs

    {
        // sin/cos table: all 16 loads of a thread in flight before the pulse trigonometry, stored after it
        // (a load -> store loop serialises 8 L2 round trips per thread: 4 us of a 55 us launch at BASELINE config 3)
        float tv[2][(TLEN > 0 ? TLEN : kThreads) / kThreads];
        if (SC == SC_TABLE) {
#pragma unroll
            for (int u = 0; u < TLEN / kThreads; ++u) {
                tv[0][u] = g_sin_table[tid + u * kThreads];
                tv[1][u] = g_cos_table[tid + u * kThreads];
            }
        }
        const float* pb = p.pulses + (size_t)b * L * 2;
        for (int i = tid; i < CT; i += kThreads) {
            const int ic = i < L ? i : L - 1;
            const int im = (i - 1) < 0 ? 0 : ((i - 1) < L ? (i - 1) : L - 1);
            const double phi = (double)pb[2 * ic];
            const double phim = (double)pb[2 * im];
            const float tau = i < L ? pb[2 * ic + 1] : 0.0f;
            double sn, cs;
            ::sincos(phi, &sn, &cs);
            fwd4[i] = make_float4((float)cs, (float)sn, tau, 0.0f);
            if (BWD) {
                double sd, cd;
                ::sincos(i == 0 ? 0.0 : phi - phim, &sd, &cd);
                bwd4[i] = make_float4((float)cd, (float)sd, tau, 0.0f);
            }
        }
        if (BWD) {
            for (int i = tid; i < kWarps * C * 2; i += kThreads) acc[i] = 0.0f;
        }
        if (SC == SC_TABLE) {
#pragma unroll
            for (int u = 0; u < TLEN / kThreads; ++u) {
                tsin[tid + u * kThreads] = tv[0][u];
                tcos[tid + u * kThreads] = tv[1][u];
            }
        }
    }
    __syncthreads();

    float cr[4], ci[4];
    su2_load_target<float>(p, b, cr, ci);
    const size_t Bm = (size_t)p.B * p.M;
    float fsum = 0.0f;

    for (int tile = split; tile < p.n_tiles; tile += p.splits) {
        // ---- per-sample constants, packed pairwise: pair u = samples (2u, 2u+1) of this thread
        F2 ka[NP], ka2[NP], kr[NP], kr2[NP], kdl[NP], kae[NP];   // ka / ka2: table-index slopes when SC == SC_TABLE
        bool valid[ST];
        size_t sidx[ST];
        {
            SampleConst<float> kc[ST];
#pragma unroll
            for (int u = 0; u < ST; ++u) {
                const long long j = (long long)tile * TS + u * SLOTS + slot;
                valid[u] = j < p.M;
                sidx[u] = (size_t)b * p.M + (size_t)(valid[u] ? j : 0);
                float delta = 0.0f, eps = 0.0f;
                if (valid[u]) {
                    su2_sample_errors<float>(p, b, j, sidx[u], Bm, delta, eps);
                    if (p.err_out != nullptr && lead) {
                        p.err_out[sidx[u]] = delta;
                        p.err_out[Bm + sidx[u]] = eps;
                    }
                }
                kc[u] = make_sample_const<float>(delta, eps);
            }
#pragma unroll
            for (int u = 0; u < NP; ++u) {
                if (SC == SC_TABLE) {
                    ka[u] = f2(kc[2 * u].ap, kc[2 * u + 1].ap);
                    ka2[u] = f2(kc[2 * u].a2p, kc[2 * u + 1].a2p);
                } else {
                    ka[u] = f2(kc[2 * u].a, kc[2 * u + 1].a);
                    ka2[u] = f2(kc[2 * u].a2, kc[2 * u + 1].a2);
                }
                kr[u] = f2(kc[2 * u].r, kc[2 * u + 1].r);
                kr2[u] = f2(kc[2 * u].r2, kc[2 * u + 1].r2);
                kdl[u] = f2(kc[2 * u].delta, kc[2 * u + 1].delta);
                kae[u] = f2(kc[2 * u].ae, kc[2 * u + 1].ae);
            }
        }

        // ---------------- forward sweep over this warp's chunk ----------------
        // The per-pulse scalars (sin/cos, q) are computed with two SAMPLES per register pair; the running product
        // keeps two COMPONENTS of one sample per pair, X = (a, b), Y = (c, d), so that every Hamilton-product
        // instruction is  acc += {X | Y with a free swap / sign modifier} * scalar: the per-sample scalar rides in
        // the 32-bit broadcast slot and X / Y come from the operand-reuse cache -- one fresh 64-bit register read
        // per FFMA2 instead of two (tools/ubench/fma_ubench.cu: 95 % vs 76 % of the pipe).
        F2 X[ST], Y[ST];
        int par[ST];
#pragma unroll
        for (int u = 0; u < ST; ++u) {
            X[u] = f2(1.0f, 0.0f);
            Y[u] = f2b(0.0f);
            par[u] = 0;
        }
#pragma unroll kX2FwdUnroll
        for (int jj = 0; jj < C; ++jj) {
            // per-pulse values are warp-uniform: f2b() lets ptxas use the 32-bit broadcast operand form (R.F32),
            // which costs no 64-bit register-file read (tools/ubench/fma_ubench.cu modes 5/6)
            const float4 row = fwd4[rb + jj];
            const F2 cc = f2b(row.x), ss = f2b(row.y), tau = f2b(row.z);
            F2 h[NP], s[NP], c[NP], sp[NP], q1[NP], q2[NP], q3[NP];
            int kb[ST];
            if constexpr (SC == SC_TABLE) {
                sincos2_tab<NP, false>(tau, ka, s, c, kb, tsin, tcos);
            } else {
#pragma unroll
                for (int u = 0; u < NP; ++u) h[u] = mul2(tau, ka[u]);
                sincos2_n<NP, SC>(h, s, c, kb, tsin, tcos);
            }
            if ((SC == SC_POLY || SC == SC_TABLE) && !BWD) {
#pragma unroll
                for (int u = 0; u < ST; ++u) par[u] ^= kb[u];
            }
#pragma unroll
            for (int u = 0; u < NP; ++u) sp[u] = mul2(s[u], kr[u]);
#pragma unroll
            for (int u = 0; u < NP; ++u) {
                q1[u] = mul2(sp[u], cc);
                q2[u] = mul2(sp[u], ss);
                q3[u] = mul2(sp[u], kdl[u]);
            }
#pragma unroll
            for (int v = 0; v < ST; ++v) {
                const int u = v >> 1;
                const F2 cs = f2b((v & 1) ? f2hi(c[u]) : f2lo(c[u]));
                const F2 a1 = f2b((v & 1) ? f2hi(q1[u]) : f2lo(q1[u]));
                const F2 a2 = f2b((v & 1) ? f2hi(q2[u]) : f2lo(q2[u]));
                const F2 a3 = f2b((v & 1) ? f2hi(q3[u]) : f2lo(q3[u]));
                // (na, nb) = c (a, b) + q1 (-b, a) + q2 (-c, d) + q3 (-d, -c)
                // (nc, nd) = c (c, d) + q1 (-d, c) + q2 (a, -b) + q3 (b, a)
                F2 nX = mul2(X[v], cs);
                F2 nY = mul2(Y[v], cs);
                nX = fma2(swp2_np(X[v]), a1, nX);
                nY = fma2(sgn2_pn(X[v]), a2, nY);
                nY = fma2(swp2(X[v]), a3, nY);
                nY = fma2(swp2_np(Y[v]), a1, nY);
                nX = fma2(sgn2_np(Y[v]), a2, nX);
                nX = fma2(swp2_nn(Y[v]), a3, nX);
                X[v] = nX;
                Y[v] = nY;
            }
        }

        // ---------------- chunk products -> prefix at this chunk's end and the full product ----------------
        Quat<float> PL[ST], Pin[ST];
#pragma unroll
        for (int u = 0; u < ST; ++u) {
            Pin[u] = Quat<float>{f2lo(X[u]), f2hi(X[u]), f2lo(Y[u]), f2hi(Y[u])};
            PL[u] = Pin[u];
        }
        if constexpr (WPS > 1) {
#pragma unroll
            for (int u = 0; u < ST; ++u) {
                float* dst = xq + ((size_t)(warp * ST + u) * 5) * 32 + lane;
                dst[0] = Pin[u].a; dst[32] = Pin[u].b; dst[64] = Pin[u].c; dst[96] = Pin[u].d;
                dst[128] = __int_as_float(par[u] & 1);
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < ST; ++u) {
                const float* s0 = xq + ((size_t)u * 5) * 32 + lane;
                Quat<float> run{s0[0], s0[32], s0[64], s0[96]};
                int ptot = __float_as_int(s0[128]);
                if (warp == 0) Pin[u] = run;
#pragma unroll
                for (int w2 = 1; w2 < WPS; ++w2) {
                    const float* sw = xq + ((size_t)(w2 * ST + u) * 5) * 32 + lane;
                    const Quat<float> Qw{sw[0], sw[32], sw[64], sw[96]};
                    ptot ^= __float_as_int(sw[128]);
                    run = qmul(Qw, run);                   // later pulses on the left
                    if (w2 == warp) Pin[u] = run;
                }
                PL[u] = run;
                par[u] = ptot;                             // parity of the whole train (U_out sign)
            }
            __syncthreads();                               // xq is rewritten by the next tile
        }

        // ---------------- fidelity epilogue (scalar, once per sample) ----------------
        float trr[ST], tri[ST];
#pragma unroll
        for (int u = 0; u < ST; ++u) {
            Quat<float> q = PL[u];
            const float n2 = q.a * q.a + q.b * q.b + q.c * q.c + q.d * q.d;
            const float inv = 1.0f / sqrtf(n2);
            q.a *= inv; q.b *= inv; q.c *= inv; q.d *= inv;
            PL[u] = q;
            trr[u] = cr[0] * q.a + cr[1] * q.b + cr[2] * q.c + cr[3] * q.d;
            tri[u] = ci[0] * q.a + ci[1] * q.b + ci[2] * q.c + ci[3] * q.d;
            const float F = (trr[u] * trr[u] + tri[u] * tri[u] + 2.0f) * (1.0f / 6.0f);
            if (valid[u] && lead) {
                fsum += F;
                if (p.F_out != nullptr) p.F_out[sidx[u]] = F;
                if (!BWD && p.U_out != nullptr) {
                    const float sg = (par[u] & 1) ? -1.0f : 1.0f;
                    float* U = p.U_out + sidx[u] * 8;
                    U[0] = sg * q.a;  U[1] = -sg * q.d;
                    U[2] = -sg * q.c; U[3] = -sg * q.b;
                    U[4] = sg * q.c;  U[5] = -sg * q.b;
                    U[6] = sg * q.a;  U[7] = sg * q.d;
                }
            }
        }

        if constexpr (BWD) {
            // ---------------- adjoint seed at the end of this warp's chunk ----------------
            F2 A[NP], Bq[NP], W3[NP];
            {
                const float4 rowL = fwd4[rb + C - 1];      // (cos phi, sin phi) of the chunk's last pulse
                float a_[ST], b_[ST], w_[ST];
#pragma unroll
                for (int u = 0; u < ST; ++u) {
                    float wgt = 0.0f;
                    if (valid[u]) wgt = p.weight != nullptr ? p.weight[sidx[u]] : 1.0f;
                    const float fr = wgt * trr[u] * (1.0f / 3.0f), fi = wgt * tri[u] * (1.0f / 3.0f);
                    const Quat<float> lam{fr * cr[0] + fi * ci[0], fr * cr[1] + fi * ci[1], fr * cr[2] + fi * ci[2],
                                          fr * cr[3] + fi * ci[3]};
                    Quat<float> Wq;
                    if constexpr (WPS > 1) {
                        const Quat<float> Lam = qmul(qconj(PL[u]), lam);
                        Wq = qmul(qmul(Pin[u], Lam), qconj(Pin[u]));
                    } else {
                        Wq = qmul(lam, qconj(PL[u]));
                    }
                    a_[u] = Wq.b * rowL.x + Wq.c * rowL.y;
                    b_[u] = Wq.c * rowL.x - Wq.b * rowL.y;
                    w_[u] = Wq.d;
                }
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    A[u] = f2(a_[2 * u], a_[2 * u + 1]);
                    Bq[u] = f2(b_[2 * u], b_[2 * u + 1]);
                    W3[u] = f2(w_[2 * u], w_[2 * u + 1]);
                }
            }
            // ---------------- backward sweep ----------------
            const F2 one = f2b(1.0f);
            for (int jb = C / NB - 1; jb >= 0; --jb) {
                float v[NV];
#pragma unroll
                for (int e = NB - 1; e >= 0; --e) {
                    const float4 row = bwd4[rb + jb * NB + e];
                    const F2 cd = f2b(row.x), sd = f2b(row.y), tau = f2b(row.z);
                    F2 gp = f2b(0.0f), gt = f2b(0.0f);
                    F2 h[NP], s[NP], c[NP], s2[NP], C2[NP], Sr[NP], k1_[NP], t[NP], uu[NP], K[NP], BS[NP], A1[NP], B1[NP], Wz[NP];
                    int kb[ST];
                    if constexpr (SC == SC_TABLE) {
                        // (sin 2h, cos 2h) straight from the full-period table
                        sincos2_tab<NP, true>(tau, ka2, s2, C2, kb, tsin, tcos);
#pragma unroll
                        for (int u = 0; u < NP; ++u) {
                            t[u] = fma2(kdl[u], W3[u], A[u]);
                            uu[u] = fma2(kdl[u], A[u], neg2(W3[u]));
                        }
#pragma unroll
                        for (int u = 0; u < NP; ++u) {
                            Sr[u] = mul2(s2[u], kr[u]);
                            gt = fma2(kae[u], t[u], gt);
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < NP; ++u) h[u] = mul2(tau, ka[u]);
                        sincos2_n<NP, SC>(h, s, c, kb, tsin, tcos);
#pragma unroll
                        for (int u = 0; u < NP; ++u) {
                            s2[u] = add2(s[u], s[u]);
                            t[u] = fma2(kdl[u], W3[u], A[u]);
                            uu[u] = fma2(kdl[u], A[u], neg2(W3[u]));
                        }
#pragma unroll
                        for (int u = 0; u < NP; ++u) {
                            C2[u] = fma2(neg2(s2[u]), s[u], one);
                            Sr[u] = mul2(mul2(s2[u], kr[u]), c[u]);
                            gt = fma2(kae[u], t[u], gt);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        k1_[u] = fma2(neg2(C2[u]), kr2[u], kr2[u]);
                        BS[u] = mul2(Bq[u], Sr[u]);
                        B1[u] = mul2(Bq[u], C2[u]);
                        gp = fma2(Sr[u], Bq[u], gp);
                    }
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        K[u] = mul2(k1_[u], t[u]);
                        gp = fma2(neg2(k1_[u]), uu[u], gp);
                        B1[u] = fma2(neg2(uu[u]), Sr[u], B1[u]);
                        Wz[u] = fma2(W3[u], C2[u], neg2(BS[u]));
                    }
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        A1[u] = fma2(A[u], C2[u], K[u]);
                        W3[u] = fma2(kdl[u], K[u], Wz[u]);
                    }
#pragma unroll
                    for (int u = 0; u < NP; ++u) A1[u] = fma2(kdl[u], BS[u], A1[u]);
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        A[u] = fma2(neg2(B1[u]), sd, mul2(A1[u], cd));
                        Bq[u] = fma2(B1[u], cd, mul2(A1[u], sd));
                    }
                    v[2 * e] = f2lo(gp) + f2hi(gp);
                    v[2 * e + 1] = f2lo(gt) + f2hi(gt);
                }
                int base = 0;
                LaneReduce<float, NV, 1>::run(v, lane, base);
                constexpr int NF = reduce_final_count(NV, 1);
                constexpr int DUP = reduce_dup_mask(NV, 1);
                if ((lane & DUP) == 0) {
                    // WPS = 1: per-warp slice of the whole train;  WPS = 4: this warp's chunk of the train
                    float* dst = acc + ((size_t)warp * C + (size_t)jb * NB) * 2 + base;
#pragma unroll
                    for (int m = 0; m < NF; ++m) dst[m] += v[m];
                }
            }
        }
    }

    {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
        if (lane == 0) scratch[warp] = fsum;
    }
    __syncthreads();
    if (tid == 0 && p.Fsum_part != nullptr) {
        float tot = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tot += scratch[w];
        p.Fsum_part[(size_t)split * p.B + b] = tot;
    }
    if constexpr (BWD) {
        float* gout = p.G_part + ((size_t)split * p.B + b) * L * 2;
        const int LC2 = C * 2;
        for (int i = tid; i < 2 * L; i += kThreads) {
            float tot;
            if constexpr (WPS == 1) {
                tot = 0.0f;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) tot += acc[(size_t)w * LC2 + i];
            } else {
                tot = acc[i];                              // chunks are consecutive: flat index == pulse index
            }
            gout[i] = (i & 1) ? tot : tot * 0.5f;
        }
    }
}

}  // namespace uqoc
