// which warps share a scheduler (SMSP)? 16 warps per CTA, only the warps in `mask` issue DMMAs
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(512, 1) k(double* out, double a, double b, int iters, unsigned mask) {
    const int warp = threadIdx.x >> 5;
    double c[16][2];
    for (int i = 0; i < 16; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    if ((mask >> warp) & 1) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) dmma(c[i][0], c[i][1], a, b);
        }
    }
    double s = 0; for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// chain-like DFMA latency probe on warp `cw` while the warps in `mask` hammer DMMAs
__global__ void __launch_bounds__(512, 1) k2(double* out, long long* cyc, double a, double b, int iters, unsigned mask, int cw) {
    const int warp = threadIdx.x >> 5;
    double c[16][2];
    for (int i = 0; i < 16; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double x = a * threadIdx.x;
    if (warp == cw) {
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) x = fma(x, b, a);
        }
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    } else if ((mask >> warp) & 1) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) dmma(c[i][0], c[i][1], a, b);
        }
    }
    double s = x; for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 512 * sizeof(double));
    long long* cyc; cudaMalloc(&cyc, 64);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    unsigned masks[] = {0x000F, 0x1111, 0x2222, 0x00F0, 0x0011, 0x0003, 0xEEEE, 0xFFFF};
    for (unsigned m : masks) {
        k<<<148, 512>>>(out, 1.0000001, 1e-9, iters, m); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<148, 512>>>(out, 1.0000001, 1e-9, iters, m); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("mask %04x (%2d warps): %.3f ms\n", m, __builtin_popcount(m), ms);
    }
    struct { unsigned m; int cw; } probes[] = {{0, 0}, {0xEEEE, 0}, {0x1110, 0}, {0x000E, 0}, {0xFFFE, 0}};
    for (auto p : probes) {
        k2<<<148, 512>>>(out, cyc, 1.0000001, 0.999, iters, p.m, p.cw); cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA chain on warp %d with DMMA mask %04x: %.1f cycles per dependent DFMA\n", p.cw, p.m, (double)h / (iters * 16.0));
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
