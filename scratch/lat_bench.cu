// dependent-issue latency of the ops on the SCD coordinate chain (single warp, clock64)
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
#define TICK(t, x) asm volatile("mov.u64 %0, %%clock64;" : "=l"(t), "+d"(x) :: "memory")
__global__ void k(double* out, long long* cyc, double a, double b, int lane_src) {
    __shared__ double sm[64];
    sm[threadIdx.x % 64] = a;
    __syncthreads();
    double x = a * (threadIdx.x + 1);
    long long t0, t1;
    // 1) DFMA chain
    TICK(t0, x);
#pragma unroll
    for (int i = 0; i < N; i++) x = fma(x, b, a);
    TICK(t1, x); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2) DADD chain
    TICK(t0, x);
#pragma unroll
    for (int i = 0; i < N; i++) x = x + a;
    TICK(t1, x); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // 3) fmax chain (DSETP+FSEL or DMNMX)
    TICK(t0, x);
#pragma unroll
    for (int i = 0; i < N; i++) x = fmax(x * 1.0000001, b);
    TICK(t1, x); if (threadIdx.x == 0) cyc[2] = t1 - t0;   // includes a DMUL per step
    // 4) DMUL chain
    TICK(t0, x);
#pragma unroll
    for (int i = 0; i < N; i++) x = x * b;
    TICK(t1, x); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 5) shuffle chain (double = 2 SHFL)
    TICK(t0, x);
#pragma unroll
    for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, lane_src);
    TICK(t1, x); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 6) LDS chain (pointer chase through shared memory)
    int idx = threadIdx.x % 64;
    __shared__ int nxt[64];
    nxt[threadIdx.x % 64] = (threadIdx.x + 1) % 64;
    __syncthreads();
    TICK(t0, x);
#pragma unroll
    for (int i = 0; i < N; i++) idx = nxt[idx];
    x += idx;
    TICK(t1, x); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 7) select on compare: x = (x < b) ? a : x + a   (DSETP + FSEL + DADD)
    TICK(t0, x);
#pragma unroll
    for (int i = 0; i < N; i++) x = (x < b) ? a : x + a;
    TICK(t1, x); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // 8) vote chain
    unsigned v = threadIdx.x;
    TICK(t0, x);
#pragma unroll
    for (int i = 0; i < N; i++) v = __ballot_sync(0xffffffffu, v & 1) + i;
    x += v;
    TICK(t1, x); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // 9) FFMA chain for reference
    float f = (float)a;
    TICK(t0, x);
#pragma unroll
    for (int i = 0; i < N; i++) f = fmaf(f, 1.0001f, 0.5f);
    x += f;
    TICK(t1, x); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    out[threadIdx.x] = x + idx + v + f;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
    k<<<1, 32>>>(out, cyc, 1.0, 0.999999, 3); cudaDeviceSynchronize();
    k<<<1, 32>>>(out, cyc, 1.0, 0.999999, 3); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA", "DADD", "DMUL+fmax", "DMUL", "SHFL(double)", "LDS", "DSETP+FSEL+DADD", "VOTE(+IADD)", "FFMA"};
    for (int i = 0; i < 9; i++) printf("%-18s %.1f cycles per dependent step\n", names[i], (double)h[i] / N);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
