// microbenchmark 2: DFMA forms used by the SCD rank-1 update
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 26;
// B: x[i] = fma(d, g[i], x[i]), g in registers (3 distinct 64-bit register operands)
__global__ void kB(double* out, const double* gin, double d, int iters) {
    double x[2 * N], g[2 * N];
    for (int i = 0; i < 2 * N; i++) { x[i] = threadIdx.x * 1e-3 + i; g[i] = gin[i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 2 * N; i++) x[i] = fma(d, g[i], x[i]);
        d += 1e-30;
    }
    double s = 0; for (int i = 0; i < 2 * N; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// C: two columns share g: x0[i] = fma(d0, g[i], x0[i]); x1[i] = fma(d1, g[i], x1[i])
__global__ void kC(double* out, const double* gin, double d0, double d1, int iters) {
    double x0[N], x1[N], g[N];
    for (int i = 0; i < N; i++) { x0[i] = threadIdx.x * 1e-3 + i; x1[i] = x0[i] * 2; g[i] = gin[i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) { x0[i] = fma(d0, g[i], x0[i]); x1[i] = fma(d1, g[i], x1[i]); }
        d0 += 1e-30; d1 += 1e-30;
    }
    double s = 0; for (int i = 0; i < N; i++) s += x0[i] + x1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// D: like C but g streamed from shared memory every iteration (13 LDS.128 per 52 DFMA)
__global__ void kD(double* out, double d0, double d1, int iters) {
    __shared__ double gs[64 * 28];
    for (int i = threadIdx.x; i < 64 * 28; i += blockDim.x) gs[i] = 1e-9 * i;
    __syncthreads();
    double x0[N], x1[N];
    for (int i = 0; i < N; i++) { x0[i] = threadIdx.x * 1e-3 + i; x1[i] = x0[i] * 2; }
    for (int it = 0; it < iters; it++) {
        const double2* gc = reinterpret_cast<const double2*>(gs + 28 * (it & 63));
#pragma unroll
        for (int i = 0; i < N / 2; i++) {
            const double2 v = gc[i];
            x0[2*i] = fma(d0, v.x, x0[2*i]); x1[2*i] = fma(d1, v.x, x1[2*i]);
            x0[2*i+1] = fma(d0, v.y, x0[2*i+1]); x1[2*i+1] = fma(d1, v.y, x1[2*i+1]);
        }
        d0 += 1e-30; d1 += 1e-30;
    }
    double s = 0; for (int i = 0; i < N; i++) s += x0[i] + x1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// E: like D with 4 columns per thread sharing g (13 rows): 7 LDS.128 per 52 DFMA
__global__ void kE(double* out, double d0, double d1, int iters) {
    __shared__ double gs[64 * 28];
    for (int i = threadIdx.x; i < 64 * 28; i += blockDim.x) gs[i] = 1e-9 * i;
    __syncthreads();
    double x[4][14];
    for (int j = 0; j < 4; j++) for (int i = 0; i < 14; i++) x[j][i] = threadIdx.x * 1e-3 + i + j;
    double d[4] = {d0, d1, d0 * 2, d1 * 2};
    for (int it = 0; it < iters; it++) {
        const double2* gc = reinterpret_cast<const double2*>(gs + 28 * (it & 63));
#pragma unroll
        for (int i = 0; i < 7; i++) {
            const double2 v = gc[i];
#pragma unroll
            for (int j = 0; j < 4; j++) { x[j][2*i] = fma(d[j], v.x, x[j][2*i]); x[j][2*i+1] = fma(d[j], v.y, x[j][2*i+1]); }
        }
        d[0] += 1e-30;
    }
    double s = 0; for (int j = 0; j < 4; j++) for (int i = 0; i < 14; i++) s += x[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    double *out, *gin; cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&gin, 64 * sizeof(double)); cudaMemset(gin, 0, 64 * 8);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int warps : {4, 8, 12, 16}) {
        auto rate = [&](float ms, double per_thread) { return 148.0 * 32 * warps * per_thread * iters / (ms * 1e-3) / 148 / (clk * 1e3); };
        float ms = timeit([&] { kB<<<148, 32 * warps>>>(out, gin, 1e-9, iters); });
        printf("warps/SM %2d  B 3-reg-operand      : %.1f DFMA/clk/SM\n", warps, rate(ms, 52));
        ms = timeit([&] { kC<<<148, 32 * warps>>>(out, gin, 1e-9, 2e-9, iters); });
        printf("warps/SM %2d  C 2 cols share g regs : %.1f DFMA/clk/SM\n", warps, rate(ms, 52));
        ms = timeit([&] { kD<<<148, 32 * warps>>>(out, 1e-9, 2e-9, iters); });
        printf("warps/SM %2d  D 2 cols, g from smem : %.1f DFMA/clk/SM\n", warps, rate(ms, 52));
        ms = timeit([&] { kE<<<148, 32 * warps>>>(out, 1e-9, 2e-9, iters); });
        printf("warps/SM %2d  E 4 cols, g from smem : %.1f DFMA/clk/SM\n", warps, rate(ms, 56));
    }
    return 0;
}
