// Operand-delivery cost of sm_100a FFMA2 forms as they occur in the packed SU(2) kernel (uqoc_su2_x2.cuh):
// cycles per instruction per SM sub-partition for 128-thread blocks at 4 / 5 / 8 blocks per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o operand_ubench operand_ubench.cu && ./operand_ubench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define MUL2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 bc(float s) { return pk(s, s); }
__device__ __forceinline__ u64 swp_np(u64 a) { float lo, hi; upk(a, lo, hi); return pk(-hi, lo); }
__device__ __forceinline__ u64 swp(u64 a) { float lo, hi; upk(a, lo, hi); return pk(hi, lo); }
__device__ __forceinline__ u64 sgn_pn(u64 a) { float lo, hi; upk(a, lo, hi); return pk(lo, -hi); }

constexpr int N = 8;
template <int MODE> __global__ void __launch_bounds__(128) k(int iters, float seed, u64* out) {
    u64 acc[N], x[N], y[N];
    float s[N], t[N], u[N];
    for (int i = 0; i < N; ++i) {
        acc[i] = pk(seed + threadIdx.x + i, seed - i);
        x[i] = pk(1.0f + 1e-7f * (threadIdx.x + i), 1.0f - 1e-7f * (threadIdx.x + 2 * i));
        y[i] = pk(1e-9f * (threadIdx.x + 3 * i), 1e-9f * (i + 2 * threadIdx.x));
        s[i] = 1.0f + 1e-7f * (i + seed + threadIdx.x);
        t[i] = 1e-9f * (i + seed + threadIdx.x);
        u[i] = 1.0f + 1e-7f * (i + seed);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (MODE == 0) FMA2(acc[i], x[i], bc(s[i]), acc[i]);                       // (P fresh, S, P acc)
                if (MODE == 1) FMA2(acc[i], x[i / 4], bc(s[i]), acc[i]);                   // X shared by 4 consecutive (reuse slot A)
                if (MODE == 2) {                                                          // as the forward product: modifiers on the shared X
                    const u64 X = x[i / 4];
                    const u64 xm = (i % 4 == 0) ? X : (i % 4 == 1) ? swp_np(X) : (i % 4 == 2) ? sgn_pn(X) : swp(X);
                    FMA2(acc[i], xm, bc(s[i]), acc[i]);
                }
                if (MODE == 3) FMA2(acc[i], x[i], bc(1.0000001f), acc[i]);                 // (P, imm, P)
                if (MODE == 4) FMA2(acc[i], acc[i], bc(s[i]), bc(1e-9f));                  // (P, S, imm)
                if (MODE == 5) MUL2(acc[i], acc[i], bc(s[i]));                             // FMUL2 (P, S)
                if (MODE == 6) FMA2(acc[i], x[i], y[(i + 1) % N], acc[i]);                 // (P, P, P)
                if (MODE == 7) FMA2(acc[i], swp_np(acc[i]), bc(t[i]), acc[i]);             // (P', S, same P)
                if (MODE == 8) FMA2(acc[i % 2], x[i], bc(s[i]), acc[i % 2]);               // 2 chains only: latency bound
                if (MODE == 9) FMA2(acc[i], x[i], bc(s[i / 4]), acc[i]);                   // scalar shared by 4 consecutive (reuse slot B)
                if (MODE == 10) FMA2(acc[i], x[i / 4], bc(s[i / 4 + 2 * (i % 2)]), acc[i]); // X shared, 2 scalars alternate
                if (MODE == 11) MUL2(acc[i], acc[i], x[i]);                                // FMUL2 (P, P)
                if (MODE == 12) FMA2(acc[i], acc[i], x[i], y[i]);                          // (P chain, P, P) a = a*x + y
                if (MODE == 14) FMA2(acc[i], x[i], bc(u[i]), acc[i]);                      // (P fresh, UR uniform scalar, P acc)
                if (MODE == 13) {                                                         // FMUL2 then FFMA2 into same acc: 2 terms
                    if (i % 2 == 0) MUL2(acc[i / 2], x[i / 2], bc(s[i]));
                    else FMA2(acc[i / 2], swp_np(x[i / 2]), bc(s[i]), acc[i / 2]);
                }
            }
        }
    }
    u64 z = 0;
    for (int i = 0; i < N; ++i) z ^= acc[i];
    if (z == 0x1234567) out[0] = z;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 neg2(u64 a) { float lo, hi; upk(a, lo, hi); return pk(-lo, -hi); }
__device__ __forceinline__ u64 swp_nn(u64 a) { float lo, hi; upk(a, lo, hi); return pk(-hi, -lo); }
__device__ __forceinline__ u64 sgn_np(u64 a) { float lo, hi; upk(a, lo, hi); return pk(-lo, hi); }
// exact forward running product of su2_kernel_x2 (component pairs), 4 samples, per-sample scalars fixed: 32 instr / iter
__global__ void __launch_bounds__(128) kfwd(int iters, float seed, u64* out) {
    u64 X[4], Y[4];
    float cs[4], a1[4], a2[4], a3[4];
    for (int v = 0; v < 4; ++v) {
        X[v] = pk(1.0f, 1e-3f * (threadIdx.x + v)); Y[v] = pk(1e-3f * v, 1e-4f * threadIdx.x);
        cs[v] = 1.0f - 1e-7f * (threadIdx.x + v + seed); a1[v] = 1e-4f * (threadIdx.x + v); a2[v] = 1e-4f * (threadIdx.x + 2 * v + 1); a3[v] = 1e-5f * (threadIdx.x + 3 * v + 2);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                u64 nX = mul2(X[v], bc(cs[v]));
                u64 nY = mul2(Y[v], bc(cs[v]));
                nX = fma2(swp_np(X[v]), bc(a1[v]), nX);
                nY = fma2(sgn_pn(X[v]), bc(a2[v]), nY);
                nY = fma2(swp(X[v]), bc(a3[v]), nY);
                nY = fma2(swp_np(Y[v]), bc(a1[v]), nY);
                nX = fma2(sgn_np(Y[v]), bc(a2[v]), nX);
                nX = fma2(swp_nn(Y[v]), bc(a3[v]), nX);
                X[v] = nX; Y[v] = nY;
            }
        }
    }
    u64 z = 0;
    for (int v = 0; v < 4; ++v) z ^= X[v] ^ Y[v];
    if (z == 0x1234567) out[0] = z;
}
// the same product with the instructions that read X (then Y) in the first operand slot issued back to back (.reuse)
__global__ void __launch_bounds__(128) kfwd2(int iters, float seed, u64* out) {
    u64 X[4], Y[4];
    float cs[4], a1[4], a2[4], a3[4];
    for (int v = 0; v < 4; ++v) {
        X[v] = pk(1.0f, 1e-3f * (threadIdx.x + v)); Y[v] = pk(1e-3f * v, 1e-4f * threadIdx.x);
        cs[v] = 1.0f - 1e-7f * (threadIdx.x + v + seed); a1[v] = 1e-4f * (threadIdx.x + v); a2[v] = 1e-4f * (threadIdx.x + 2 * v + 1); a3[v] = 1e-5f * (threadIdx.x + 3 * v + 2);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                u64 nX = mul2(X[v], bc(cs[v]));
                u64 nY = mul2(sgn_pn(X[v]), bc(a2[v]));
                nY = fma2(swp(X[v]), bc(a3[v]), nY);
                nX = fma2(swp_np(X[v]), bc(a1[v]), nX);
                nY = fma2(Y[v], bc(cs[v]), nY);
                nY = fma2(swp_np(Y[v]), bc(a1[v]), nY);
                nX = fma2(sgn_np(Y[v]), bc(a2[v]), nX);
                nX = fma2(swp_nn(Y[v]), bc(a3[v]), nX);
                X[v] = nX; Y[v] = nY;
            }
        }
    }
    u64 z = 0;
    for (int v = 0; v < 4; ++v) z ^= X[v] ^ Y[v];
    if (z == 0x1234567) out[0] = z;
}
// exact backward core of su2_kernel_x2 (two samples per pair, NP = 2), sin/cos and per-pulse values fixed: 38 FFMA2-class + 4 FADD / iter
__global__ void __launch_bounds__(128) kbwd(int iters, float seed, u64* out) {
    u64 A[2], Bq[2], W3[2], kdl[2], kr[2], kr2[2], kae[2], s2[2], C2[2];
    for (int u = 0; u < 2; ++u) {
        A[u] = pk(1e-3f * (threadIdx.x + u), 2e-3f * (threadIdx.x + u)); Bq[u] = pk(1e-3f * u + 0.1f, 1e-4f * threadIdx.x); W3[u] = pk(0.3f + 1e-4f * threadIdx.x, 0.2f);
        kdl[u] = pk(1e-3f * threadIdx.x, 2e-3f * threadIdx.x + u); kr[u] = pk(0.9f - 1e-4f * threadIdx.x, 0.8f); kr2[u] = mul2(kr[u], kr[u]); kae[u] = pk(0.5f + 1e-4f * threadIdx.x, 0.6f);
        s2[u] = pk(1e-3f * (threadIdx.x + u + seed), 2e-3f * (threadIdx.x + 1)); C2[u] = pk(1.0f - 1e-6f * threadIdx.x, 1.0f - 2e-6f * (threadIdx.x + u));
    }
    const float cdf = 1.0f - 1e-6f * seed, sdf = 1e-3f * seed;
    float acc0 = 0, acc1 = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const u64 cd = bc(cdf), sd = bc(sdf);
            u64 gp = bc(0.0f), gt = bc(0.0f);
            u64 Sr[2], k1_[2], t[2], uu[2], K[2], BS[2], A1[2], B1[2], Wz[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) { t[u] = fma2(kdl[u], W3[u], A[u]); uu[u] = fma2(kdl[u], A[u], neg2(W3[u])); }
#pragma unroll
            for (int u = 0; u < 2; ++u) { Sr[u] = mul2(s2[u], kr[u]); gt = fma2(kae[u], t[u], gt); }
#pragma unroll
            for (int u = 0; u < 2; ++u) { k1_[u] = fma2(neg2(C2[u]), kr2[u], kr2[u]); BS[u] = mul2(Bq[u], Sr[u]); B1[u] = mul2(Bq[u], C2[u]); gp = fma2(Sr[u], Bq[u], gp); }
#pragma unroll
            for (int u = 0; u < 2; ++u) { K[u] = mul2(k1_[u], t[u]); gp = fma2(neg2(k1_[u]), uu[u], gp); B1[u] = fma2(neg2(uu[u]), Sr[u], B1[u]); Wz[u] = fma2(W3[u], C2[u], neg2(BS[u])); }
#pragma unroll
            for (int u = 0; u < 2; ++u) { A1[u] = fma2(A[u], C2[u], K[u]); W3[u] = fma2(kdl[u], K[u], Wz[u]); }
#pragma unroll
            for (int u = 0; u < 2; ++u) A1[u] = fma2(kdl[u], BS[u], A1[u]);
#pragma unroll
            for (int u = 0; u < 2; ++u) { A[u] = fma2(neg2(B1[u]), sd, mul2(A1[u], cd)); Bq[u] = fma2(B1[u], cd, mul2(A1[u], sd)); }
            float lo, hi; upk(gp, lo, hi); acc0 += lo + hi; upk(gt, lo, hi); acc1 += lo + hi;
        }
    }
    u64 z = 0;
    for (int u = 0; u < 2; ++u) z ^= A[u] ^ Bq[u] ^ W3[u];
    if (z == 0x1234567 || acc0 + acc1 == 1.2345f) out[0] = z;
}

// backward core re-ordered so that consecutive instructions share a register in the same operand slot (reuse cache)
#define VFMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define VMUL2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
template <int VOL>
__global__ void __launch_bounds__(128) kbwd2(int iters, float seed, u64* out) {
    u64 A[2], Bq[2], W3[2], kdl[2], kr[2], kr2[2], kae[2], s2[2], C2[2];
    for (int u = 0; u < 2; ++u) {
        A[u] = pk(1e-3f * (threadIdx.x + u), 2e-3f * (threadIdx.x + u)); Bq[u] = pk(1e-3f * u + 0.1f, 1e-4f * threadIdx.x); W3[u] = pk(0.3f + 1e-4f * threadIdx.x, 0.2f);
        kdl[u] = pk(1e-3f * threadIdx.x, 2e-3f * threadIdx.x + u); kr[u] = pk(0.9f - 1e-4f * threadIdx.x, 0.8f); kr2[u] = mul2(kr[u], kr[u]); kae[u] = pk(0.5f + 1e-4f * threadIdx.x, 0.6f);
        s2[u] = pk(1e-3f * (threadIdx.x + u + seed), 2e-3f * (threadIdx.x + 1)); C2[u] = pk(1.0f - 1e-6f * threadIdx.x, 1.0f - 2e-6f * (threadIdx.x + u));
    }
    const float cdf = 1.0f - 1e-6f * seed, sdf = 1e-3f * seed;
    float acc0 = 0, acc1 = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const u64 cd = bc(cdf), sd = bc(sdf), nsd = bc(-sdf);
            u64 gp = bc(0.0f), gt = bc(0.0f);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                u64 t, uu, Sr, k1, B1, BS, K, Wz, A1, tA, tB;
                if (VOL) {
                    VFMA2(t, kdl[u], W3[u], A[u]);
                    VFMA2(uu, kdl[u], A[u], neg2(W3[u]));
                    VMUL2(Sr, kr[u], s2[u]);
                    VFMA2(k1, kr2[u], neg2(C2[u]), kr2[u]);
                    VMUL2(B1, Bq[u], C2[u]);
                    VMUL2(BS, Bq[u], Sr);
                    VFMA2(gp, Bq[u], Sr, gp);
                    VFMA2(gt, t, kae[u], gt);
                    VMUL2(K, t, k1);
                    VFMA2(gp, uu, neg2(k1), gp);
                    VFMA2(B1, uu, neg2(Sr), B1);
                    VFMA2(Wz, C2[u], W3[u], neg2(BS));
                    VFMA2(A1, C2[u], A[u], K);
                    VFMA2(W3[u], kdl[u], K, Wz);
                    VFMA2(A1, kdl[u], BS, A1);
                    VMUL2(tA, A1, cd);
                    VMUL2(tB, A1, sd);
                    VFMA2(A[u], B1, nsd, tA);
                    VFMA2(Bq[u], B1, cd, tB);
                } else {
                    t = fma2(kdl[u], W3[u], A[u]);
                    uu = fma2(kdl[u], A[u], neg2(W3[u]));
                    Sr = mul2(kr[u], s2[u]);
                    k1 = fma2(kr2[u], neg2(C2[u]), kr2[u]);
                    B1 = mul2(Bq[u], C2[u]);
                    BS = mul2(Bq[u], Sr);
                    gp = fma2(Bq[u], Sr, gp);
                    gt = fma2(t, kae[u], gt);
                    K = mul2(t, k1);
                    gp = fma2(uu, neg2(k1), gp);
                    B1 = fma2(uu, neg2(Sr), B1);
                    Wz = fma2(C2[u], W3[u], neg2(BS));
                    A1 = fma2(C2[u], A[u], K);
                    W3[u] = fma2(kdl[u], K, Wz);
                    A1 = fma2(kdl[u], BS, A1);
                    tA = mul2(A1, cd);
                    tB = mul2(A1, sd);
                    A[u] = fma2(B1, nsd, tA);
                    Bq[u] = fma2(B1, cd, tB);
                }
            }
            float lo, hi; upk(gp, lo, hi); acc0 += lo + hi; upk(gt, lo, hi); acc1 += lo + hi;
        }
    }
    u64 z = 0;
    for (int u = 0; u < 2; ++u) z ^= A[u] ^ Bq[u] ^ W3[u];
    if (z == 0x1234567 || acc0 + acc1 == 1.2345f) out[0] = z;
}
// scalar FFMA, 3 fresh registers, 16 chains
__global__ void __launch_bounds__(128) ks(int iters, float seed, float* out) {
    float a[16], b[16], c[16];
    for (int i = 0; i < 16; ++i) { a[i] = seed + threadIdx.x + i; b[i] = 1.0f + 1e-7f * i; c[i] = 1e-9f * (i + threadIdx.x); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[i]) : "f"(a[i]), "f"(b[i]), "f"(c[i]));
        }
    }
    float z = 0;
    for (int i = 0; i < 16; ++i) z += a[i];
    if (z == 1234.5f) out[0] = z;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    return best;
}
int main() {
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    void* out; cudaMalloc(&out, 64);
    const int iters = 4096;
    const char* names[] = {"(P,S,Pacc) fresh X", "X shared x4 (reuse A)", "X shared x4 + swap/sign modifiers", "(P,imm,Pacc)", "(Pchain,S,imm)",
                           "FMUL2 (P,S)", "(P,P,Pacc)", "(P',S,sameP)", "2 chains (latency)", "S shared x4", "X shared x4, S alternating",
                           "FMUL2 (P,P)", "(Pchain,P,P)", "FMUL2+FFMA2 term pairs", "(P,UR,Pacc) uniform scalar"};
    for (int bps : {4, 5, 8}) {
        const int blocks = sms * bps;
        const double winst = (double)bps * 4 /*warps*/ * iters * 4.0 * N / 4.0;      // warp instructions per SM sub-partition
#define RUN(M) { float t = timeit([&] { k<M><<<blocks, 128>>>(iters, 3.f, (u64*)out); }); \
        printf("blocks/SM=%d mode %2d %-36s %.3f ms  %.2f cycles/instr/SMSP (at %d MHz)\n", bps, M, names[M], t, t * 1e-3 * khz * 1e3 / winst, khz / 1000); }
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14)
        { float t;
          const double wi = (double)bps * 4 * iters * 2.0 * 16 / 4.0;
          t = timeit([&] { kfwd<<<blocks, 128>>>(iters, 3.f, (u64*)out); });
          printf("blocks/SM=%d forward product replica (32 instr/iter x2)   %.3f ms  %.2f cycles/instr/SMSP\n", bps, t, t * 1e-3 * khz * 1e3 / ((double)bps * iters * 64.0));
          t = timeit([&] { kfwd2<<<blocks, 128>>>(iters, 3.f, (u64*)out); });
          printf("blocks/SM=%d forward product, X-then-Y operand order          %.3f ms  %.2f cycles/instr/SMSP\n", bps, t, t * 1e-3 * khz * 1e3 / ((double)bps * iters * 64.0));
          t = timeit([&] { kbwd<<<blocks, 128>>>(iters, 3.f, (u64*)out); });
          printf("blocks/SM=%d backward core replica (38 packed + 4 FADD /iter x2) %.3f ms  %.2f cycles/packed-instr/SMSP\n", bps, t, t * 1e-3 * khz * 1e3 / ((double)bps * iters * 76.0));
          t = timeit([&] { kbwd2<1><<<blocks, 128>>>(iters, 3.f, (u64*)out); });
          printf("blocks/SM=%d backward core, reuse-ordered (volatile asm)          %.3f ms  %.2f cycles/packed-instr/SMSP\n", bps, t, t * 1e-3 * khz * 1e3 / ((double)bps * iters * 76.0));
          t = timeit([&] { kbwd2<0><<<blocks, 128>>>(iters, 3.f, (u64*)out); });
          printf("blocks/SM=%d backward core, reuse-ordered (helpers)               %.3f ms  %.2f cycles/packed-instr/SMSP\n", bps, t, t * 1e-3 * khz * 1e3 / ((double)bps * iters * 76.0));
          t = timeit([&] { ks<<<blocks, 128>>>(iters, 3.f, (float*)out); });
          printf("blocks/SM=%d scalar FFMA (R,R,R) 16 chains              %.3f ms  %.2f cycles/instr/SMSP\n", bps, t, t * 1e-3 * khz * 1e3 / wi); }
    }
    return 0;
}
