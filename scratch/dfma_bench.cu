// microbenchmark: DFMA throughput per SM vs warps per scheduler and independent chains per thread
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, double a, double b, int iters) {
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0; for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// shared-broadcast operand variant: x[i] = fma(d, g[i], x[i]) with g from shared memory (LDS.128 broadcast)
template <int ILP>
__global__ void ks(double* out, double d, int iters) {
    __shared__ double g[64 * 64];
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) g[i] = 1e-9 * i;
    __syncthreads();
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
        const double2* gc = reinterpret_cast<const double2*>(g + 64 * (it & 63));
#pragma unroll
        for (int i = 0; i < ILP / 2; i++) { double2 v = gc[i]; x[2*i] = fma(d, v.x, x[2*i]); x[2*i+1] = fma(d, v.y, x[2*i+1]); }
    }
    double s = 0; for (int i = 0; i < ILP; i++) s += x[i];
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
    for (int warps : {4, 8, 12, 16, 32}) {
        float ms = timeit([&] { k<52><<<148, 32 * warps>>>(out, 1.0000001, 1e-9, iters); });
        double dfma = 148.0 * 32 * warps * 52.0 * iters;
        printf("reg-only  ILP52 warps/SM %2d: %.3f ms  %.1f DFMA/clk/SM (at %d MHz nominal)  %.2f TFLOPS\n", warps, ms, dfma / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000, 2 * dfma / (ms * 1e-3) / 1e12);
        ms = timeit([&] { k<8><<<148, 32 * warps>>>(out, 1.0000001, 1e-9, iters); });
        dfma = 148.0 * 32 * warps * 8.0 * iters;
        printf("reg-only  ILP8  warps/SM %2d: %.3f ms  %.1f DFMA/clk/SM\n", warps, ms, dfma / (ms * 1e-3) / 148 / (clk * 1e3));
        ms = timeit([&] { ks<52><<<148, 32 * warps>>>(out, 1e-9, iters); });
        dfma = 148.0 * 32 * warps * 52.0 * iters;
        printf("smem-bcast ILP52 warps/SM %2d: %.3f ms  %.1f DFMA/clk/SM\n", warps, ms, dfma / (ms * 1e-3) / 148 / (clk * 1e3));
    }
    return 0;
}
