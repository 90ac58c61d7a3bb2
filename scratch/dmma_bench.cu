// DMMA (mma.sync m8n8k4 f64) throughput and latency on B200
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k(double* out, double a, double b, int iters) {
    double c[ILP][2];
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0; for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
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
    for (int warps : {4, 8, 16, 32}) {
        float ms = timeit([&] { k<16><<<148, 32 * warps>>>(out, 1.0000001, 1e-9, iters); });
        double n = 148.0 * warps * 16.0 * iters;   // warp-level DMMAs
        printf("ILP16 warps/SM %2d: %.3f ms  %.3f DMMA/clk/SM  = %.1f FMA/clk/SM  %.1f TFLOPS\n", warps, ms, n / (ms * 1e-3) / 148 / (clk * 1e3),
               256 * n / (ms * 1e-3) / 148 / (clk * 1e3), 512 * n / (ms * 1e-3) / 1e12);
        ms = timeit([&] { k<1><<<148, 32 * warps>>>(out, 1.0000001, 1e-9, iters); });
        n = 148.0 * warps * 1.0 * iters;
        printf("ILP1  warps/SM %2d: %.3f ms  latency-bound: %.1f cycles per dependent DMMA (if 1 warp/SMSP)\n", warps, ms, ms * 1e-3 * clk * 1e3 / iters);
    }
    return 0;
}
