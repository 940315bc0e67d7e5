// How accurate is MUFU.SIN/COS (__sinf/__cosf) when the argument sits on a coarse grid of the circle?
// For m = 10..24: h_q = fl(k * 2pi / 2^m), |h_q| <= hmax; report max |__sinf(h_q) - sin(h_q)| (double truth).
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void probe(int m, float hmax, double* out /* [4]: max es, max ec, max angle err, sum sq */) {
    const double TWO_PI = 6.283185307179586476925;
    const long long n = 1LL << m;
    const float step = (float)(TWO_PI / (double)n);
    double mes = 0, mec = 0, mang = 0, ssq = 0;
    long long cnt = 0;
    const long long kmax = (long long)(hmax / step);
    for (long long k = -kmax + threadIdx.x + (long long)blockIdx.x * blockDim.x; k <= kmax; k += (long long)blockDim.x * gridDim.x) {
        const float hq = (float)k * step;
        const float s = __sinf(hq), c = __cosf(hq);
        const double ts = sin((double)hq), tc = cos((double)hq);
        const double es = (double)s - ts, ec = (double)c - tc;
        const double ang = tc * es - ts * ec;
        mes = fmax(mes, fabs(es)); mec = fmax(mec, fabs(ec)); mang = fmax(mang, fabs(ang));
        ssq += ang * ang; ++cnt;
    }
    // crude block/grid max via atomics on doubles-as-ull (values are non-negative)
    atomicMax((unsigned long long*)&out[0], __double_as_longlong(mes));
    atomicMax((unsigned long long*)&out[1], __double_as_longlong(mec));
    atomicMax((unsigned long long*)&out[2], __double_as_longlong(mang));
    atomicAdd(&out[3], ssq);
}
__global__ void probe_random(float hmax, int n, double* out) {
    double mang = 0, ssq = 0;
    for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += blockDim.x * gridDim.x) {
        unsigned x = i * 2654435761u; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        const float h = hmax * ((float)x * 2.3283064e-10f * 2.f - 1.f);
        const float s = __sinf(h), c = __cosf(h);
        const double ts = sin((double)h), tc = cos((double)h);
        const double ang = tc * ((double)s - ts) - ts * ((double)c - tc);
        mang = fmax(mang, fabs(ang)); ssq += ang * ang;
    }
    atomicMax((unsigned long long*)&out[2], __double_as_longlong(mang));
    atomicAdd(&out[3], ssq);
}
int main() {
    double* d; cudaMalloc(&d, 32); double h[4];
    for (float hmax : {0.5f, 1.57f, 3.14f}) {
        cudaMemset(d, 0, 32);
        probe_random<<<256, 256>>>(hmax, 1 << 22, d);
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("random |h|<=%.2f: max angle err %.3e rms %.3e\n", hmax, h[2], sqrt(h[3] / (1 << 22)));
        for (int m = 10; m <= 24; m += 1) {
            cudaMemset(d, 0, 32);
            probe<<<256, 256>>>(m, hmax, d);
            cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
            printf("  grid 2pi/2^%2d |h|<=%.2f: max|es| %.3e max|ec| %.3e max angle err %.3e\n", m, hmax, h[0], h[1], h[2]);
        }
    }
    return 0;
}
