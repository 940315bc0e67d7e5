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

// Stage-major evaluation for NP independent pairs (polynomial / MUFU policies): consecutive instructions are
// independent (NP pairs x {sin, cos} Horner chains), which is what keeps the 2-cycle FFMA2 pipe fed from a
// single warp instead of relying on other warps to hide each dependent-issue latency.
template <int NP, int SC>
__device__ __forceinline__ void sincos2_n(const F2 (&h)[NP], F2 (&s)[NP], F2 (&c)[NP], int (&kb)[2 * NP]) {
    if constexpr (SC == SC_MUFU) {
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

// Inner-loop formulations under test (tools/variants.sh builds them side by side; the winner becomes the default):
//   UQOC_X2_FWD_FORM  0 = per-pulse scalars two SAMPLES per register pair, separate sin / cos tables (round 1)
//                     1 = two COMPONENTS of one sample per pair after the table-index arithmetic, interleaved table
//   UQOC_X2_BWD_FORM  0 = two samples per pair, separate tables (round 1)
//                     1 = two samples per pair, table step per sample on the interleaved {sin, cos} pair, re-paired
//                     2 = (A, W3) component pair + scalar FFMA for the rest
#ifndef UQOC_X2_FWD_FORM
#define UQOC_X2_FWD_FORM 1
#endif
#ifndef UQOC_X2_BWD_FORM
#define UQOC_X2_BWD_FORM 0
#endif
#ifndef UQOC_X2_BWD_ORDER
#define UQOC_X2_BWD_ORDER 0      // 1 = reuse-ordered backward core (BWD_FORM 0 / 1 only)
#endif
//   UQOC_X2_BWD_CORE  0 = d/dphi accumulated per pulse as Sr B - k1 uu, A' from its own two FMAs (rounds 1-2)
//                     1 = two identities of the adjoint rotation about the pulse axis n = (1, 0, delta):
//                         (a) d/dphi of a pulse is the z-torque, W3(before) - W3(after) (expand W3' = C2 W3 - Sr B + delta k1 t
//                             with k1 (1 + delta^2) = 1 - C2), and z is untouched by the frame change between pulses, so the
//                             sample sum telescopes: only S_i = sum_s W3_s entering pulse i is accumulated (one FADD2 per
//                             pulse instead of four three-register FFMA2) and d/dphi_i = (S_i - S_{i-1}) / 2 is formed once
//                             per block in the epilogue (S_{-1}: the sweep's exit value, kept per warp);
//                         (b) the axis component t = A + delta W3 is invariant, so A' = t - delta W3' (one FFMA2 for two).
//                         19 -> 16.5 FMA-pipe instructions per sample pair and pulse in the backward core.
#ifndef UQOC_X2_BWD_CORE
#define UQOC_X2_BWD_CORE 1
#endif
constexpr int kFwdForm = UQOC_X2_FWD_FORM, kBwdForm = UQOC_X2_BWD_FORM, kBwdOrder = UQOC_X2_BWD_ORDER, kBwdCore = UQOC_X2_BWD_CORE;
static_assert(kBwdCore == 0 || (kBwdForm <= 1 && kBwdOrder == 0), "the telescoped backward core exists for the two-samples-per-pair sweep only");
//   UQOC_X2_IDX3      three-instruction table-index arithmetic on the shifted-node tables (g_*_table_sh,
//                     tools/gen_sincos_table.py): kf = fma(tau, a N/pi, M3), kc = fma(kf, pi/N, -C0) = (pi/N) k + e0 with ONE
//                     rounding, r = fma(tau, a, -kc): the residual comes out in radians and the multiplication of the
//                     index-unit residual by pi/N (one FMUL2 per sample pair and pulse) disappears.  The rounding of kc
//                     (half an ulp of the ANGLE) is an extra, per-pulse random angle error of the size of the slope's own
//                     rounding.  Bit 0 = forward sweep, bit 1 = backward sweep.  Measured on the bench launch (4096
//                     targets x 4096 samples x L = 256) and the GPU suite (profiles/r2_bwd_core_identities.txt):
//                       0  neither         7.84 ms   max |dF| 2.9e-6 (L = 256 sweep), 8.8e-6 (worst case of the suite)
//                       2  backward only   7.81 ms   the same F bit for bit; rel dG 1.2e-6 (bound 1e-4)
//                       3  both (default)  7.69 ms   max |dF| 3.3e-6 / 9.4e-6 (bound 1e-5), rel dG 1.1e-6
//                     Build with -DUQOC_X2_IDX3=2 for the exact forward residual (kf, M - kf, fma, * pi/N, plain tables).
#ifndef UQOC_X2_IDX3
#define UQOC_X2_IDX3 3
#endif
constexpr bool kIdx3Fwd = (UQOC_X2_IDX3 & 1) != 0, kIdx3Bwd = (UQOC_X2_IDX3 & 2) != 0;
constexpr int kExitSlot = 8;          // scratch[kExitSlot + warp]: sum over this warp's samples of W3 leaving its sweep (kBwdCore = 1)
// table entries staged for a kernel: interleaved {sin, cos} pairs and / or separate sin[] / cos[] arrays
constexpr int x2_tab_il(bool table, bool bwd) {
    return !table ? 0 : ((bwd && kBwdForm >= 1) ? UQOC_SINCOS_TABLE_LEN : ((kFwdForm == 1) ? kTabN : 0));
}
constexpr int x2_tab_sep(bool table, bool bwd) {
    return !table ? 0 : ((bwd && kBwdForm == 0) ? UQOC_SINCOS_TABLE_LEN : ((kFwdForm == 0) ? kTabN : 0));
}

// Table sin/cos of tau * a for NP sample pairs, from the per-sample table-index slope kap = a * N/pi:
//   kf = fma(tau, kap, MAGIC)  (k in the low mantissa bits),  frac = fma(tau, kap, -(kf - MAGIC))  (one rounding),
//   r = frac * pi/N,  (s, c) = (st + r ct, ct - r st).
// 6 FMA-pipe instructions; the angle tau*a itself is never formed.  FULL: index the full-period table (k mod 2N), signs
// exact (the backward sweep looks up the DOUBLE angle 2h this way); otherwise k mod N and the common sign
// (-1)^(k div N) is dropped (kb returns k >> 10 so the U_out kernel can track it).
// Separate sin[] / cos[] arrays, two samples per register pair throughout:
template <int NP, bool FULL, bool kIdx3>
__device__ __forceinline__ void sincos2_tab_sep(F2 tau, const F2 (&kap)[NP], const F2 (&kar)[NP], F2 (&s)[NP], F2 (&c)[NP],
                                                int (&kb)[2 * NP], const float* __restrict__ tsin, const float* __restrict__ tcos) {
    constexpr int MASK = FULL ? (2 * kTabN - 1) : (kTabN - 1);
    const float MAGIC = kIdx3 ? UQOC_SINCOS_M3 : 12582912.0f;
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
    if constexpr (kIdx3 == 1) {
        // kar = the slope in radians: r = tau a - ((pi/N) k + e0), the nodes of the shifted tables
#pragma unroll
        for (int u = 0; u < NP; ++u) kf[u] = fma2(kf[u], f2b(UQOC_SINCOS_STEP), f2b(-UQOC_SINCOS_C0));
#pragma unroll
        for (int u = 0; u < NP; ++u) fr[u] = fma2(tau, kar[u], neg2(kf[u]));
    } else {
#pragma unroll
        for (int u = 0; u < NP; ++u) kf[u] = sub2(f2b(MAGIC), kf[u]);
#pragma unroll
        for (int u = 0; u < NP; ++u) fr[u] = fma2(tau, kap[u], kf[u]);
#pragma unroll
        for (int u = 0; u < NP; ++u) fr[u] = mul2(fr[u], f2b(0.0030679615757712823f));
    }
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        s[u] = fma2(fr[u], ct[u], st[u]);
        c[u] = fma2(neg2(fr[u]), st[u], ct[u]);
    }
}
// INTERLEAVED {sin, cos} table (one LDS.64 per sample); the residual step is done per SAMPLE on the loaded component
// pair V = (st, ct):  V' = V + (-r) (-ct, st)  -- one FFMA2 whose multiplier rides in the 32-bit broadcast slot and whose
// other two sources are the SAME register pair (a swap / sign modifier apart).  Outputs V[v] = (sin, cos), v = 2u + half.
template <int NP, bool FULL, bool kIdx3>
__device__ __forceinline__ void sincos2_tab_il(F2 tau, const F2 (&kap)[NP], const F2 (&kar)[NP], F2 (&V)[2 * NP],
                                               int (&kb)[2 * NP], const u64* __restrict__ tsc) {
    constexpr int MASK = FULL ? (2 * kTabN - 1) : (kTabN - 1);
    const float MAGIC = kIdx3 ? UQOC_SINCOS_M3 : 12582912.0f;
    F2 kf[NP], nr[NP], T[2 * NP];
#pragma unroll
    for (int u = 0; u < NP; ++u) kf[u] = fma2(tau, kap[u], f2b(MAGIC));
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        const int k0 = __float_as_int(f2lo(kf[u])), k1 = __float_as_int(f2hi(kf[u]));
        kb[2 * u] = FULL ? 0 : (k0 >> 10);
        kb[2 * u + 1] = FULL ? 0 : (k1 >> 10);
#if defined(UQOC_X2_PROBE) && UQOC_X2_PROBE == 1      // timing probe (wrong results): conflict-free look-ups
        T[2 * u].v = tsc[(k0 & 0) + (threadIdx.x & 31)];
        T[2 * u + 1].v = tsc[(k1 & 0) + 32 + (threadIdx.x & 31)];
#elif defined(UQOC_X2_PROBE) && UQOC_X2_PROBE == 2    // timing probe (wrong results): no look-ups at all
        T[2 * u] = f2(__int_as_float(k0 & MASK) * 1e-9f, 1.0f);
        T[2 * u + 1] = f2(__int_as_float(k1 & MASK) * 1e-9f, 1.0f);
#else
        UQOC_ASSERT((unsigned)(k0 & MASK) < (unsigned)(FULL ? UQOC_SINCOS_TABLE_LEN : kTabN));
        T[2 * u].v = tsc[k0 & MASK];
        T[2 * u + 1].v = tsc[k1 & MASK];
#endif
    }
    if constexpr (kIdx3 == 1) {
#pragma unroll
        for (int u = 0; u < NP; ++u) kf[u] = fma2(kf[u], f2b(UQOC_SINCOS_STEP), f2b(-UQOC_SINCOS_C0));   // (pi/N) k + e0
#pragma unroll
        for (int u = 0; u < NP; ++u) nr[u] = fma2(tau, neg2(kar[u]), kf[u]);                             // -r
    } else {
#pragma unroll
        for (int u = 0; u < NP; ++u) kf[u] = sub2(f2b(MAGIC), kf[u]);
#pragma unroll
        for (int u = 0; u < NP; ++u) nr[u] = fma2(tau, kap[u], kf[u]);
#pragma unroll
        for (int u = 0; u < NP; ++u) nr[u] = mul2(nr[u], f2b(-0.0030679615757712823f));    // -r = -frac * pi/N
    }
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        V[2 * u] = fma2(swp2_np(T[2 * u]), f2b(f2lo(nr[u])), T[2 * u]);                 // (st + r ct, ct - r st)
        V[2 * u + 1] = fma2(swp2_np(T[2 * u + 1]), f2b(f2hi(nr[u])), T[2 * u + 1]);
    }
}

// shared memory of one block of the packed kernel (bytes).  C = pulses per chunk, WPS chunks, VB virtual blocks
// (128-thread sample groups that share the staged pulse train and the table); fin_threads > 0 reserves the scratch of
// the in-kernel epilogue (which reuses the block's shared memory from offset 0 once the sweeps are done).
__host__ __device__ inline size_t su2_x2_smem_bytes(int C, int wps, int st, bool bwd, bool table, int vb = 1, int fin_threads = 0) {
    const size_t rows = (size_t)C * wps;
    size_t bytes = 0;
    bytes += (size_t)(x2_tab_il(table, bwd) + x2_tab_sep(table, bwd)) * 2 * sizeof(float);   // bwd: full period
    bytes += rows * 16;                                           // {cos phi, sin phi, tau, -} rows
    if (bwd) bytes += rows * 16;                                  // {cos dphi, sin dphi, tau, -} rows
    size_t per_vb = 32 * sizeof(float);                           // block-reduction scratch
    if (bwd) per_vb += (size_t)kWarps * C * 2 * sizeof(float);    // gradient accumulators (per warp / per chunk)
    if (wps > 1) per_vb += (size_t)kWarps * st * 5 * 32 * sizeof(float);   // chunk-product (+ parity) exchange
    bytes += per_vb * vb;
    const size_t fin = fin_threads > 0 ? su2_fin_smem_bytes(fin_threads, sizeof(float)) : 0;
    return bytes > fin ? bytes : fin;
}

// resident-block hint of the table kernel: 5 blocks/SM (<= 96 registers) measured best (8.33 ms vs 8.50 at 6, 8.37 at 7,
// 8.58 at 4 on the 4096 x 4096 x 256 launch); tools/variants.sh builds such variants side by side
#ifndef UQOC_X2_MINB
#define UQOC_X2_MINB 5
#endif
#ifndef UQOC_X2_FWD_UNROLL
#define UQOC_X2_FWD_UNROLL 8
#endif
constexpr int kX2FwdUnroll = UQOC_X2_FWD_UNROLL;
constexpr int kX2FatVB = 7;           // virtual blocks of the fat-block variant (NP = 1, WPS = 4): 896 threads x 72 registers
constexpr int x2_min_blocks(int NP, int SC, int WPS, int VB) {
    return VB > 1 ? 1 : ((WPS == 4 && NP == 1) ? 7 : ((SC == SC_TABLE) ? UQOC_X2_MINB : 1));
}

// barrier over the 128 threads of one virtual block (VB > 1: named barrier 1 + vb; VB == 1: the block barrier)
template <int VB>
__device__ __forceinline__ void vb_sync(int vb) {
    if constexpr (VB == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(vb + 1), "n"(kThreads) : "memory");
}

// NP  = sample PAIRS per thread (1 or 2)
// WPS = warps per sample group.  1: every warp owns its own 32*ST samples and the whole pulse train.
//       4: the four warps of a (virtual) block share 32*ST samples and each owns a quarter of the pulse train
//       (chunk products exchanged through shared memory; the prefix alone seeds every chunk's backward
//       sweep, see uqoc_su2_kernels.cuh) -- 4x the warps for the same samples, for small sample counts.
// VB  = virtual blocks per block ("fat block", few targets): VB x 128 threads, one block per SM; the virtual blocks
//       are independent sample-tile streams of the SAME target that share one staged pulse train / trig / table
//       (staged once per SM instead of once per 128 threads) and whose gradient slices are summed in shared memory,
//       so a target leaves `cps` partial rows (one per block) instead of `splits`.
template <int NP, int SC, bool BWD, int WPS, int VB>
__global__ void __launch_bounds__(kThreads * VB, x2_min_blocks(NP, SC, WPS, VB)) su2_kernel_x2(const Su2Params<float> p) {
    constexpr int ST = 2 * NP;
    constexpr int NB = 8;
    constexpr int SLOTS = (WPS == 1) ? kThreads : 32;     // sample slots per virtual block
    constexpr int TS = SLOTS * ST;
    constexpr int NV = 2 * NB;
    constexpr int NT = kThreads * VB;

    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int C = p.C;                 // pulses per chunk (multiple of NB)
    const int CT = C * WPS;            // staged rows
    constexpr int TIL = x2_tab_il(SC == SC_TABLE, BWD), TSEP = x2_tab_sep(SC == SC_TABLE, BWD);   // table entries staged
    // which sweep reads which staged table decides its node set (plain k pi/N or shifted, see UQOC_X2_IDX3)
    constexpr bool kIlBwd = BWD && kBwdForm >= 1, kIlFwd = kFwdForm == 1, kSepBwd = BWD && kBwdForm == 0, kSepFwd = kFwdForm == 0;
    static_assert(!(kIlBwd && kIlFwd) || kIdx3Fwd == kIdx3Bwd, "one interleaved table serves both sweeps: same node set");
    static_assert(!(kSepBwd && kSepFwd) || kIdx3Fwd == kIdx3Bwd, "one sin[] / cos[] table serves both sweeps: same node set");
    constexpr bool kIlShift = kIlBwd ? kIdx3Bwd : kIdx3Fwd, kSepShift = kSepBwd ? kIdx3Bwd : kIdx3Fwd;
    // layout: [{sin, cos} pairs @ 0][fwd rows][bwd rows][sin[] | cos[]][per virtual block: scratch, acc, xq].  The separate
    // tables sit BEHIND the run-time sized rows on purpose: ptxas then addresses them as [R + UR + imm]; at a
    // compile-time offset it materialises the base in a vector register and adds it per look-up (+4 % on the step)
    const u64* tsc = reinterpret_cast<const u64*>(smem_raw);
    float4* fwd4 = reinterpret_cast<float4*>(smem_raw + (size_t)TIL * 8);
    float4* bwd4 = fwd4 + CT;
    float* tsin = reinterpret_cast<float*>(bwd4 + (BWD ? CT : 0));
    float* tcos = tsin + TSEP;
    float* vb_base = tcos + TSEP;
    const int acc_len = BWD ? kWarps * C * 2 : 0;
    const int xq_len = (WPS > 1) ? kWarps * ST * 5 * 32 : 0;
    const int vb_len = 32 + acc_len + xq_len;

    const int tid = threadIdx.x;
#ifdef UQOC_LL_TIMING
    if (blockIdx.x == 0 && tid == 0 && p.cps > 1) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); (reinterpret_cast<unsigned long long*>(p.G_part) - 16)[6] = t_; }
#endif
#ifdef UQOC_FIN_TIMING
    if (p.fin.ticket != nullptr && blockIdx.x == 0 && tid == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        reinterpret_cast<unsigned long long*>(p.fin.ticket)[7] = t;
    }
#endif
    const int vb = (VB == 1) ? 0 : tid / kThreads;         // virtual block
    const int vt = (VB == 1) ? tid : tid % kThreads;       // thread inside it
    const int lane = vt & 31;
    const int warp = vt >> 5;
    float* scratch = vb_base + (size_t)vb * vb_len;
    float* acc = scratch + 32;
    float* xq = acc + acc_len;                              // [kWarps][ST][5][32], WPS > 1 only

    UQOC_ASSERT((size_t)(reinterpret_cast<unsigned char*>(vb_base + (size_t)VB * vb_len) - smem_raw) <= dyn_smem_bytes());
    UQOC_ASSERT(C % NB == 0 && C >= NB && (long long)C * WPS >= p.L);

    const int cblk = blockIdx.x % p.cps;                   // this block among the target's blocks
    const int b = blockIdx.x / p.cps;
    const int split = cblk * VB + vb;                      // sample-tile stream
    const int L = p.L;
    const int rb = (WPS == 1) ? 0 : warp * C;              // first staged row of this warp's chunk
    const int slot = (WPS == 1) ? vt : lane;
    const bool lead = (WPS == 1) || warp == 0;             // the warp that reports per-sample outputs

    {
        // sin/cos tables: all loads of a thread in flight before the pulse trigonometry, stored after it
        // (a load -> store loop serialises the L2 round trips: 4 us of a 55 us launch at BASELINE config 3)
        constexpr int TVI = (TIL * 2 / 4 + NT - 1) / NT, TVS = (TSEP / 4 + NT - 1) / NT;   // float4 per thread
        float4 tvi[TVI > 0 ? TVI : 1], tvs[2][TVS > 0 ? TVS : 1];
#pragma unroll
        for (int u = 0; u < TVI; ++u) {
            const int i = tid + u * NT;
            if (i < TIL / 2) tvi[u] = reinterpret_cast<const float4*>(kIlShift ? g_sincos_table_sh : g_sincos_table)[i];
        }
#pragma unroll
        for (int u = 0; u < TVS; ++u) {
            const int i = tid + u * NT;
            if (i < TSEP / 4) {
                tvs[0][u] = reinterpret_cast<const float4*>(kSepShift ? g_sin_table_sh : g_sin_table)[i];
                tvs[1][u] = reinterpret_cast<const float4*>(kSepShift ? g_cos_table_sh : g_cos_table)[i];
            }
        }
        // everything above read immutable tables only; the previous kernel of the stream (the optimiser step that wrote the
        // pulses, or the previous step's epilogue still reading the partial rows this kernel will overwrite) must be complete
        // from here on (the launch is a programmatic dependent launch, su2_launch_x2w)
        grid_dependency_wait();
        for (int i = tid; i < CT; i += NT) {
            const int ic = i < L ? i : L - 1;
            const int im = (i - 1) < 0 ? 0 : ((i - 1) < L ? (i - 1) : L - 1);
            float ph_c, ta_c, ph_m, ta_m;                   // the stored pulse, or the head applied to its logits
            su2_pulse_at<float>(p, b, ic, ph_c, ta_c);
            su2_pulse_at<float>(p, b, im, ph_m, ta_m);
            if (p.head.pulses_out != nullptr && cblk == 0 && i < L) {
                p.head.pulses_out[((size_t)b * L + i) * 2] = ph_c;
                p.head.pulses_out[((size_t)b * L + i) * 2 + 1] = ta_c;
            }
            const double phi = (double)ph_c;
            const double phim = (double)ph_m;
            const float tau = i < L ? ta_c : 0.0f;
            double sn, cs;
            ::sincos(phi, &sn, &cs);
            UQOC_ASSERT(i < CT && ic < L && im < L);
            fwd4[i] = make_float4((float)cs, (float)sn, tau, 0.0f);
            if (BWD) {
                double sd, cd;
                ::sincos(i == 0 ? 0.0 : phi - phim, &sd, &cd);
                bwd4[i] = make_float4((float)cd, (float)sd, tau, 0.0f);
            }
        }
        if (BWD) {
            for (int i = vt; i < acc_len; i += kThreads) acc[i] = 0.0f;
            if (vt < 32) scratch[vt] = 0.0f;
        }
#pragma unroll
        for (int u = 0; u < TVI; ++u) {
            const int i = tid + u * NT;
            if (i < TIL / 2) reinterpret_cast<float4*>(smem_raw)[i] = tvi[u];
        }
#pragma unroll
        for (int u = 0; u < TVS; ++u) {
            const int i = tid + u * NT;
            if (i < TSEP / 4) {
                reinterpret_cast<float4*>(tsin)[i] = tvs[0][u];
                reinterpret_cast<float4*>(tcos)[i] = tvs[1][u];
            }
        }
    }
    __syncthreads();

    float cr[4], ci[4];
    su2_load_target<float>(p, b, cr, ci);
    const size_t Bm = (size_t)p.B * p.M;
    float fsum = 0.0f;

    for (int tile = split; tile < p.n_tiles; tile += p.splits) {
        // ---- per-sample constants; the table-index slopes ride two SAMPLES per register pair (pair u = samples 2u, 2u+1)
        F2 ka[NP], ka2[NP];      // half / full angle per unit tau: table-index slopes (SC_TABLE) or radians
        F2 kar[NP], kar2[NP];    // the same slopes in radians (SC_TABLE: the residual of the three-instruction index arithmetic)
        F2 kr[NP], kr2[NP], kdl[NP], kae[NP];   // two-samples-per-pair constants
        F2 RD[ST];               // per sample (1/w, delta/w): (s', q3) = sin h * RD
        float kdl_s[ST], kr_s[ST], kr2_s[ST], kae_s[ST];
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
                RD[u] = f2(kc[u].r, kc[u].rd);
                kdl_s[u] = kc[u].delta;
                kr_s[u] = kc[u].r;
                kr2_s[u] = kc[u].r2;
                kae_s[u] = kc[u].ae;
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
                kar[u] = f2(kc[2 * u].a, kc[2 * u + 1].a);
                kar2[u] = f2(kc[2 * u].a2, kc[2 * u + 1].a2);
                kr[u] = f2(kc[2 * u].r, kc[2 * u + 1].r);
                kr2[u] = f2(kc[2 * u].r2, kc[2 * u + 1].r2);
                kdl[u] = f2(kc[2 * u].delta, kc[2 * u + 1].delta);
                kae[u] = f2(kc[2 * u].ae, kc[2 * u + 1].ae);
            }
        }

        // ---------------- forward sweep over this warp's chunk ----------------
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
            const float4 row = fwd4[rb + jj];
            const F2 tau = f2b(row.z);
            float cs_[ST], a1_[ST], a2_[ST], a3_[ST];      // per sample: cos h, q1, q2, q3 (32-bit broadcast operands)
            int kb[ST];
            if constexpr (kFwdForm == 0 || SC != SC_TABLE) {
                // two samples per register pair (per-pulse values are warp-uniform: f2b() lets ptxas use the 32-bit
                // broadcast operand form, which costs no 64-bit register-file read)
                const F2 cc = f2b(row.x), ss = f2b(row.y);
                F2 h[NP], sn[NP], c[NP], sp[NP], q1[NP], q2[NP], q3[NP];
                if constexpr (SC == SC_TABLE) {
                    sincos2_tab_sep<NP, false, kIdx3Fwd>(tau, ka, kar, sn, c, kb, tsin, tcos);
                } else {
#pragma unroll
                    for (int u = 0; u < NP; ++u) h[u] = mul2(tau, ka[u]);
                    sincos2_n<NP, SC>(h, sn, c, kb);
                }
#pragma unroll
                for (int u = 0; u < NP; ++u) sp[u] = mul2(sn[u], kr[u]);
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    q1[u] = mul2(sp[u], cc);
                    q2[u] = mul2(sp[u], ss);
                    q3[u] = mul2(sp[u], kdl[u]);
                }
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    cs_[2 * u] = f2lo(c[u]);   cs_[2 * u + 1] = f2hi(c[u]);
                    a1_[2 * u] = f2lo(q1[u]);  a1_[2 * u + 1] = f2hi(q1[u]);
                    a2_[2 * u] = f2lo(q2[u]);  a2_[2 * u + 1] = f2hi(q2[u]);
                    a3_[2 * u] = f2lo(q3[u]);  a3_[2 * u + 1] = f2hi(q3[u]);
                }
            } else {
                // two components of one sample per pair after the table-index arithmetic: V = (sin h, cos h),
                // (s', q3) = sin h (1/w, delta/w), (q1, q2) = s' (cos phi, sin phi)
                const F2 CS = f2(row.x, row.y);
                F2 V[ST], SQ[ST], Q12[ST];
                sincos2_tab_il<NP, false, kIdx3Fwd>(tau, ka, kar, V, kb, tsc);
#pragma unroll
                for (int v = 0; v < ST; ++v) SQ[v] = mul2(RD[v], f2b(f2lo(V[v])));
#pragma unroll
                for (int v = 0; v < ST; ++v) Q12[v] = mul2(CS, f2b(f2lo(SQ[v])));
#pragma unroll
                for (int v = 0; v < ST; ++v) {
                    cs_[v] = f2hi(V[v]);
                    a1_[v] = f2lo(Q12[v]);
                    a2_[v] = f2hi(Q12[v]);
                    a3_[v] = f2hi(SQ[v]);
                }
            }
            if ((SC == SC_POLY || SC == SC_TABLE) && !BWD) {
#pragma unroll
                for (int u = 0; u < ST; ++u) par[u] ^= kb[u];
            }
            // running product: two COMPONENTS of one sample per pair, X = (a, b), Y = (c, d); every instruction is
            // acc += {X | Y with a free swap / sign modifier} * scalar (ptxas folds the modifiers only when the modified
            // pair is the FIRST source operand)
#pragma unroll
            for (int v = 0; v < ST; ++v) {
                const F2 cs = f2b(cs_[v]), a1 = f2b(a1_[v]), a2 = f2b(a2_[v]), a3 = f2b(a3_[v]);
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
                UQOC_ASSERT((size_t)(dst + 128 - xq) < (size_t)xq_len);
                dst[0] = Pin[u].a; dst[32] = Pin[u].b; dst[64] = Pin[u].c; dst[96] = Pin[u].d;
#ifdef UQOC_DEBUG_CHECKS
                dst[128] = __int_as_float((par[u] & 1) | (tile << 1));      // epoch tag: which tile this product belongs to
#else
                dst[128] = __int_as_float(par[u] & 1);
#endif
            }
            vb_sync<VB>(vb);
#pragma unroll
            for (int u = 0; u < ST; ++u) {
                const float* s0 = xq + ((size_t)u * 5) * 32 + lane;
                Quat<float> run{s0[0], s0[32], s0[64], s0[96]};
                UQOC_ASSERT((__float_as_int(s0[128]) >> 1) == tile);       // a stale slot = a missing barrier
                int ptot = __float_as_int(s0[128]) & 1;
                if (warp == 0) Pin[u] = run;
#pragma unroll
                for (int w2 = 1; w2 < WPS; ++w2) {
                    const float* sw = xq + ((size_t)(w2 * ST + u) * 5) * 32 + lane;
                    const Quat<float> Qw{sw[0], sw[32], sw[64], sw[96]};
                    UQOC_ASSERT((__float_as_int(sw[128]) >> 1) == tile);
                    ptot ^= __float_as_int(sw[128]) & 1;
                    run = qmul(Qw, run);                   // later pulses on the left
                    if (w2 == warp) Pin[u] = run;
                }
                PL[u] = run;
                par[u] = ptot;                             // parity of the whole train (U_out sign)
            }
            vb_sync<VB>(vb);                               // xq is rewritten by the next tile
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
            float a_[ST], b_[ST], w_[ST];
            {
                const float4 rowL = fwd4[rb + C - 1];      // (cos phi, sin phi) of the chunk's last pulse
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
            }
            // ---------------- backward sweep ----------------
            if constexpr (kBwdForm <= 1 || SC != SC_TABLE) {
                // two SAMPLES per register pair (A, Bq, W3 of samples 2u, 2u+1)
                F2 A[NP], Bq[NP], W3[NP];
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    A[u] = f2(a_[2 * u], a_[2 * u + 1]);
                    Bq[u] = f2(b_[2 * u], b_[2 * u + 1]);
                    W3[u] = f2(w_[2 * u], w_[2 * u + 1]);
                }
                const F2 one = f2b(1.0f);
                for (int jb = C / NB - 1; jb >= 0; --jb) {
                    float v[NV];
#pragma unroll
                    for (int e = NB - 1; e >= 0; --e) {
                        const float4 row = bwd4[rb + jb * NB + e];
                        const F2 cd = f2b(row.x), sd = f2b(row.y), tau = f2b(row.z);
                        F2 gp = f2b(0.0f), gt = f2b(0.0f);
                        F2 s2[NP], C2[NP], Sr[NP], k1_[NP], t[NP], uu[NP], K[NP], BS[NP], A1[NP], B1[NP], Wz[NP];
                        int kb[ST];
                        if constexpr (SC == SC_TABLE && kBwdForm == 0) {
                            // (sin 2h, cos 2h) straight from the full-period table
                            sincos2_tab_sep<NP, true, kIdx3Bwd>(tau, ka2, kar2, s2, C2, kb, tsin, tcos);
                        } else if constexpr (SC == SC_TABLE) {
                            F2 V2[ST];                      // per sample on the interleaved pair, then re-paired
                            sincos2_tab_il<NP, true, kIdx3Bwd>(tau, ka2, kar2, V2, kb, tsc);
#pragma unroll
                            for (int u = 0; u < NP; ++u) {
                                s2[u] = f2(f2lo(V2[2 * u]), f2lo(V2[2 * u + 1]));
                                C2[u] = f2(f2hi(V2[2 * u]), f2hi(V2[2 * u + 1]));
                            }
                        } else {
                            F2 h[NP], s[NP], c[NP];
#pragma unroll
                            for (int u = 0; u < NP; ++u) h[u] = mul2(tau, ka[u]);
                            sincos2_n<NP, SC>(h, s, c, kb);
#pragma unroll
                            for (int u = 0; u < NP; ++u) {
                                const F2 sh = add2(s[u], s[u]);
                                C2[u] = fma2(neg2(sh), s[u], one);   // cos 2h
                                s2[u] = mul2(sh, c[u]);              // sin 2h
                            }
                        }
                        if constexpr (kBwdOrder == 1) {
                            // one sample pair at a time, consecutive instructions sharing a register in the SAME operand
                            // slot: ptxas flags it .reuse and the FFMA2 then reads two fresh register pairs instead of
                            // three (tools/ubench/operand_ubench.cu: 2.2 instead of 3.06 cycles; core 2.44 vs 2.55)
                            const F2 nsd = f2b(-row.y);
#pragma unroll
                            for (int u = 0; u < NP; ++u) {
                                t[u] = fma2(kdl[u], W3[u], A[u]);
                                uu[u] = fma2(kdl[u], A[u], neg2(W3[u]));
                                Sr[u] = mul2(kr[u], s2[u]);
                                k1_[u] = fma2(kr2[u], neg2(C2[u]), kr2[u]);
                                B1[u] = mul2(Bq[u], C2[u]);
                                BS[u] = mul2(Bq[u], Sr[u]);
                                gp = fma2(Bq[u], Sr[u], gp);
                                gt = fma2(t[u], kae[u], gt);
                                K[u] = mul2(t[u], k1_[u]);
                                gp = fma2(uu[u], neg2(k1_[u]), gp);
                                B1[u] = fma2(uu[u], neg2(Sr[u]), B1[u]);
                                Wz[u] = fma2(C2[u], W3[u], neg2(BS[u]));
                                A1[u] = fma2(C2[u], A[u], K[u]);
                                W3[u] = fma2(kdl[u], K[u], Wz[u]);
                                A1[u] = fma2(kdl[u], BS[u], A1[u]);
                                const F2 tA = mul2(A1[u], cd), tB = mul2(A1[u], sd);
                                A[u] = fma2(B1[u], nsd, tA);
                                Bq[u] = fma2(B1[u], cd, tB);
                            }
                        } else if constexpr (kBwdCore == 1) {
                            // S_i: the z-components entering this pulse, summed over the thread's samples
                            gp = W3[0];
#pragma unroll
                            for (int u = 1; u < NP; ++u) gp = add2(gp, W3[u]);
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
#pragma unroll
                            for (int u = 0; u < NP; ++u) {
                                k1_[u] = fma2(neg2(C2[u]), kr2[u], kr2[u]);
                                BS[u] = mul2(Bq[u], Sr[u]);
                                B1[u] = mul2(Bq[u], C2[u]);
                            }
#pragma unroll
                            for (int u = 0; u < NP; ++u) {
                                K[u] = mul2(k1_[u], t[u]);
                                B1[u] = fma2(neg2(uu[u]), Sr[u], B1[u]);
                                Wz[u] = fma2(W3[u], C2[u], neg2(BS[u]));
                            }
#pragma unroll
                            for (int u = 0; u < NP; ++u) W3[u] = fma2(kdl[u], K[u], Wz[u]);
#pragma unroll
                            for (int u = 0; u < NP; ++u) A1[u] = fma2(neg2(kdl[u]), W3[u], t[u]);     // t is invariant
#pragma unroll
                            for (int u = 0; u < NP; ++u) {
                                A[u] = fma2(neg2(B1[u]), sd, mul2(A1[u], cd));
                                Bq[u] = fma2(B1[u], cd, mul2(A1[u], sd));
                            }
                        } else {
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
                        UQOC_ASSERT(base >= 0 && ((size_t)warp * C + (size_t)jb * NB) * 2 + base + NF <= (size_t)acc_len);
#pragma unroll
                        for (int m = 0; m < NF; ++m) dst[m] += v[m];
                    }
                }
                if constexpr (kBwdCore == 1) {
                    // S_{-1} of this warp's sweep: what its samples' z-components are after the sweep's first pulse
                    F2 ex = W3[0];
#pragma unroll
                    for (int u = 1; u < NP; ++u) ex = add2(ex, W3[u]);
                    float e1 = f2lo(ex) + f2hi(ex);
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) e1 += __shfl_xor_sync(0xffffffffu, e1, d);
                    if (lane == 0) scratch[kExitSlot + warp] += e1;
                }
            } else {
                // per sample: Z = (A, W3) as one register pair, Bq as a scalar.
                //   (t, -uu) = Z + (-delta)(-W3, A);  Sr = sin 2h / w, k1 = (1 - cos 2h)/w^2;
                //   d/dtau += ae t,  d/dphi += Sr Bq + k1 (-uu);  Y = (k1 t, -Bq Sr);
                //   Z <- cos 2h Z + Y + delta (-Y.hi, Y.lo),  B1 = Bq cos 2h + (-uu) Sr;  then the frame change by dphi.
                F2 Z[ST];
                float Bs[ST];
#pragma unroll
                for (int u = 0; u < ST; ++u) {
                    Z[u] = f2(a_[u], w_[u]);
                    Bs[u] = b_[u];
                }
                for (int jb = C / NB - 1; jb >= 0; --jb) {
                    float v[NV];
#pragma unroll
                    for (int e = NB - 1; e >= 0; --e) {
                        const float4 row = bwd4[rb + jb * NB + e];
                        const float cd = row.x, sd = row.y;
                        const F2 tau = f2b(row.z);
                        F2 V2[ST];                         // (sin 2h, cos 2h)
                        int kb[ST];
                        sincos2_tab_il<NP, true, kIdx3Bwd>(tau, ka2, kar2, V2, kb, tsc);
                        float gp = 0.0f, gt = 0.0f;
                        F2 T2[ST], Yk[ST], Z1[ST];
                        float Sr[ST], k1[ST], B1[ST];
#pragma unroll
                        for (int u = 0; u < ST; ++u) T2[u] = fma2(swp2_np(Z[u]), f2b(-kdl_s[u]), Z[u]);      // (t, -uu)
#pragma unroll
                        for (int u = 0; u < ST; ++u) {
                            Sr[u] = f2lo(V2[u]) * kr_s[u];
                            k1[u] = fmaf(-f2hi(V2[u]), kr2_s[u], kr2_s[u]);
                        }
#pragma unroll
                        for (int u = 0; u < ST; ++u) {
                            gt = fmaf(kae_s[u], f2lo(T2[u]), gt);
                            gp = fmaf(Sr[u], Bs[u], gp);
                            Yk[u] = f2(k1[u] * f2lo(T2[u]), -(Bs[u] * Sr[u]));
                            B1[u] = Bs[u] * f2hi(V2[u]);
                        }
#pragma unroll
                        for (int u = 0; u < ST; ++u) {
                            gp = fmaf(k1[u], f2hi(T2[u]), gp);
                            B1[u] = fmaf(f2hi(T2[u]), Sr[u], B1[u]);
                            Z1[u] = fma2(Z[u], f2b(f2hi(V2[u])), Yk[u]);
                        }
#pragma unroll
                        for (int u = 0; u < ST; ++u) Z1[u] = fma2(swp2_np(Yk[u]), f2b(kdl_s[u]), Z1[u]);
#pragma unroll
                        for (int u = 0; u < ST; ++u) {
                            const float A1 = f2lo(Z1[u]);
                            Bs[u] = fmaf(A1, sd, B1[u] * cd);
                            Z[u] = f2(fmaf(-B1[u], sd, A1 * cd), f2hi(Z1[u]));
                        }
                        v[2 * e] = gp;
                        v[2 * e + 1] = gt;
                    }
                    int base = 0;
                    LaneReduce<float, NV, 1>::run(v, lane, base);
                    constexpr int NF = reduce_final_count(NV, 1);
                    constexpr int DUP = reduce_dup_mask(NV, 1);
                    if ((lane & DUP) == 0) {
                        float* dst = acc + ((size_t)warp * C + (size_t)jb * NB) * 2 + base;
                        UQOC_ASSERT(base >= 0 && ((size_t)warp * C + (size_t)jb * NB) * 2 + base + NF <= (size_t)acc_len);
#pragma unroll
                        for (int m = 0; m < NF; ++m) dst[m] += v[m];
                    }
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
        float tot = 0.0f;
#pragma unroll
        for (int q = 0; q < VB; ++q) {
#pragma unroll
            for (int w = 0; w < kWarps; ++w) tot += vb_base[(size_t)q * vb_len + w];
        }
        p.Fsum_part[(size_t)cblk * p.B + b] = tot;
    }
    if constexpr (BWD) {
        const int LC2 = C * 2;
        auto col_total = [&](int i) {                      // fixed-order sum of accumulator column i over warps / virtual blocks
            float tot = 0.0f;
#pragma unroll
            for (int q = 0; q < VB; ++q) {
                const float* a_q = vb_base + (size_t)q * vb_len + 32;
                if constexpr (WPS == 1) {
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) tot += a_q[(size_t)w * LC2 + i];
                } else {
                    tot += a_q[i];                         // chunks are consecutive: flat index == pulse index
                }
            }
            return tot;
        };
        auto dphi_total = [&](int l) {                     // d/dphi of pulse l (before the factor 1/2)
            if constexpr (kBwdCore == 0) {
                return col_total(2 * l);
            } else {
                // telescoped z-torque: S_l - S_{l-1}; below the first pulse of a sweep, the sweep's exit sum
                const bool first = (WPS == 1) ? (l == 0) : (l % C == 0);
                float prev = 0.0f;
                if (first) {
#pragma unroll
                    for (int q = 0; q < VB; ++q) {
                        const float* s_q = vb_base + (size_t)q * vb_len + kExitSlot;
                        if constexpr (WPS == 1) {
#pragma unroll
                            for (int w = 0; w < kWarps; ++w) prev += s_q[w];
                        } else {
                            prev += s_q[l / C];
                        }
                    }
                } else {
                    prev = col_total(2 * l - 2);
                }
                return col_total(2 * l) - prev;
            }
        };
        if (p.head.mode == 0) {
            float* gout = p.G_part + ((size_t)cblk * p.B + b) * L * 2;
            for (int i = tid; i < 2 * L; i += NT) gout[i] = (i & 1) ? col_total(i) : dphi_total(i >> 1) * 0.5f;
        } else {
            // head backward: one thread per pulse, (d/dphi, d/dtau) -> the head's input row
            const int po = su2_grad_width(p);
            float* gout = p.G_part + ((size_t)cblk * p.B + b) * L * po;
            for (int l = tid; l < L; l += NT) {
                float o[3];
                su2_head_grad<float>(p, b, l, dphi_total(l) * 0.5f, col_total(2 * l + 1), o);
                for (int c = 0; c < po; ++c) gout[(size_t)po * l + c] = o[c];
            }
        }
        if (p.fin.ticket != nullptr) {
            __syncthreads();                               // the accumulators are dead: the epilogue reuses shared memory
            const FinParams<float> fin = p.fin;            // a copy: the kernel parameters themselves stay in the constant bank
            su2_block_finalize<float>(fin, p.G_part, p.Fsum_part, p.cps, p.B, p.L, smem_raw, su2_grad_width(p));
        }
#ifdef UQOC_LL_TIMING
        if (tid == 0 && p.cps > 1) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMax(reinterpret_cast<unsigned long long*>(p.G_part) - 16 + 7, t_); }
#endif
    }
}

}  // namespace uqoc
