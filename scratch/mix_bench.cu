// do DFMA and DMMA share an issue port / execution unit on B200? time DMMA-only, DFMA-only and both interleaved
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NM, int NF>
__global__ void k(double* out, double a, double b, int iters) {
    double c[8][2], f[16];
    for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int i = 0; i < 16; i++) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i < NM) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int j = 0; j < 2; j++) if (2 * i + j < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[2 * i + j]) : "d"(a), "d"(b));
        }
    }
    double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    for (int i = 0; i < 16; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * sizeof(double));
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int warps : {4, 8, 16}) {
        float m = timeit([&] { k<8, 0><<<148, 32 * warps>>>(out, 1.0000001, 1e-9, iters); });
        float f = timeit([&] { k<0, 16><<<148, 32 * warps>>>(out, 1.0000001, 1e-9, iters); });
        float b = timeit([&] { k<8, 16><<<148, 32 * warps>>>(out, 1.0000001, 1e-9, iters); });
        float b2 = timeit([&] { k<8, 8><<<148, 32 * warps>>>(out, 1.0000001, 1e-9, iters); });
        double cyc = clk * 1e3 * 1e-3 / iters;   // cycles per iteration per ms
        printf("warps/SM %2d: 8 DMMA %.1f cyc/iter | 16 DFMA %.1f cyc/iter | both %.1f | 8 DMMA + 8 DFMA %.1f  (per SMSP-warp set)\n", warps,
               m * cyc, f * cyc, b * cyc, b2 * cyc);
    }
    return 0;
}
