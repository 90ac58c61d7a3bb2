// can a DFMA stream take its warp-uniform operand from the constant bank at full rate, and up to which footprint?
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cV[8000];
template <int F>
__global__ void __launch_bounds__(384, 1) k(double* out, double d0, int iters) {
    double acc[48];
#pragma unroll
    for (int i = 0; i < 48; i++) acc[i] = threadIdx.x + i;
    double d = d0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < F; i++) acc[i % 48] = fma(d, cV[i], acc[i % 48]);
        d += 1e-9;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 48; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename Fn> float timeit(Fn f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
template <int F> void run(double* out, int clk) {
    const int iters = 400000 / F * 10;
    for (int warps : {4, 8, 12}) {
        float ms = timeit([&] { k<F><<<148, 32 * warps>>>(out, 1.0000001, iters); });
        double per = ms * 1e-3 * clk * 1e3 / ((double)iters * F * warps / 4);   // cycles per warp-DFMA per SMSP
        printf("footprint %5d doubles (%5.1f KB) warps/SM %2d: %.3f ms  %.2f cycles per DFMA per SMSP (2.0 = full rate)\n", F, F * 8 / 1024.0, warps, ms, per);
    }
}
int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * sizeof(double));
    double h[8000]; for (int i = 0; i < 8000; i++) h[i] = 1e-3 * i;
    cudaMemcpyToSymbol(cV, h, sizeof h);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    run<256>(out, clk); run<1024>(out, clk); run<2560>(out, clk); run<4096>(out, clk); run<8000>(out, clk);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
