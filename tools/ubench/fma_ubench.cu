// Micro-benchmarks of the sm_100a FP32 FMA pipe: scalar FFMA vs packed FFMA2 with 1, 2 or 3
// distinct register operands per instruction (register-file read bandwidth / bank effects).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_ubench fma_ubench.cu && ./fma_ubench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int N = 8;   // independent chains per thread
// mode 0: a = a*x + y (x,y loop-invariant)    mode 1: a_i = a_i*b_i + y     mode 2: a_i = a_i*b_i + c_i   mode 3: a_i = b_i*c_i + a_j (rotating)
template <int MODE> __global__ void __launch_bounds__(256) k_ffma2(int iters, u64 seed, u64* out) {
    u64 a[N], b[N], c[N];
    for (int i = 0; i < N; ++i) { a[i] = seed + threadIdx.x + i; b[i] = seed * 3 + i; c[i] = seed * 7 + i * 5; }
    u64 x = seed * 11, y = seed * 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (MODE == 0) a[i] = ffma2(a[i], x, y);
                if (MODE == 1) a[i] = ffma2(a[i], b[i], y);
                if (MODE == 2) a[i] = ffma2(a[i], b[i], c[i]);
                if (MODE == 3) a[i] = ffma2(b[i], c[(i + 1) % N], a[i]);
                if (MODE == 4) a[i] = fmul2(a[i], b[i]);
                if (MODE == 5) { unsigned lo = (unsigned)b[i]; u64 bb; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "r"(lo)); a[i] = ffma2(a[i], bb, c[i]); }
                if (MODE == 6) { unsigned lo = (unsigned)b[i]; u64 bb; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "r"(lo)); a[i] = ffma2(a[i], bb, y); }
            }
        }
    }
    u64 s = 0; for (int i = 0; i < N; ++i) s ^= a[i];
    if (s == 0x1234567) out[0] = s;
}
template <int MODE> __global__ void __launch_bounds__(256) k_ffma(int iters, float seed, float* out) {
    float a[2 * N], b[2 * N], c[2 * N];
    for (int i = 0; i < 2 * N; ++i) { a[i] = seed + threadIdx.x + i; b[i] = seed * 3 + i; c[i] = seed * 7 + i * 5; }
    float x = seed * 11, y = seed * 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 2 * N; ++i) {
                if (MODE == 0) a[i] = ffma(a[i], x, y);
                if (MODE == 1) a[i] = ffma(a[i], b[i], y);
                if (MODE == 2) a[i] = ffma(a[i], b[i], c[i]);
                if (MODE == 3) a[i] = ffma(b[i], c[(i + 1) % (2 * N)], a[i]);
            }
        }
    }
    float s = 0; for (int i = 0; i < 2 * N; ++i) s += a[i];
    if (s == 1234.5f) out[0] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    return best;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void* out; cudaMalloc(&out, 64);
    const int iters = 4096;
    for (int wpb = 1; wpb <= 8; wpb *= 2) {       // blocks per SM (x 8 warps)
        const int blocks = sms * wpb, threads = 256;
        const double flops2 = (double)blocks * threads * iters * 4.0 * N * 4.0;       // 2 lanes x 2 flop
        const double flops1 = (double)blocks * threads * iters * 4.0 * 2 * N * 2.0;
        float t;
#define RUN2(M) t = timeit([&] { k_ffma2<M><<<blocks, threads>>>(iters, 3, (u64*)out); }); printf("FFMA2 mode%d warps/SM=%2d: %.2f TFLOP/s\n", M, wpb * 8, flops2 / t / 1e9);
#define RUN1(M) t = timeit([&] { k_ffma<M><<<blocks, threads>>>(iters, 3.f, (float*)out); }); printf("FFMA  mode%d warps/SM=%2d: %.2f TFLOP/s\n", M, wpb * 8, flops1 / t / 1e9);
        RUN2(0) RUN2(1) RUN2(2) RUN2(3) RUN2(5) RUN2(6)
        t = timeit([&] { k_ffma2<4><<<blocks, threads>>>(iters, 3, (u64*)out); }); printf("FMUL2       warps/SM=%2d: %.2f Tmul-lanes x2/s (as FLOP: %.2f)\n", wpb * 8, flops2 / t / 1e9, flops2 / t / 2e9);
        RUN1(0) RUN1(1) RUN1(2) RUN1(3)
    }
    return 0;
}
