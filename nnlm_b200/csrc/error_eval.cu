// error_eval.cu — a8/a9: the loss bookkeeping of the outer loop (src/nnmf.cpp:121-149, 164-177, 224-240).
// The reference materialises Ahat = W.t()*H (n x m) and makes two more passes over it; here a CTA forms a 64x64 tile
// of Ahat in registers (fp64 FMAs), folds it against the matching tile of A, and emits one partial record, so Ahat
// never exists in memory. Partials are reduced in a fixed order (launch_reduce_partials) -> reproducible totals.
// Runs only on `trace` iterations and at exit, so it is kept in fp64 for parity of the reported mse/mkl/target vectors.
#include "kernels.cuh"

namespace nnlm {

namespace {

constexpr int ET = 64;   // tile edge
constexpr int ELD = ET + 1;   // odd leading dimension: conflict-free transposing stores
constexpr int EKC = 64;  // coordinates staged per pass

template <typename TA>
__global__ void __launch_bounds__(256)
k_error(const TA* __restrict__ A, const double* __restrict__ W, const double* __restrict__ H, int k, int64_t n, int64_t m,
        double* __restrict__ part)
{
    extern __shared__ double sm[];
    double* ws = sm;                   // [EKC][ELD]
    double* hs = sm + (size_t)EKC * ELD;  // [EKC][ELD]
    __shared__ double red[2][8];
    const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
    const int64_t i0 = (int64_t)blockIdx.x * ET, j0 = (int64_t)blockIdx.y * ET;

    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
    // the rank is walked in chunks of EKC coordinates so the staging buffers stay 2 x 33 KB for every k
    for (int c0 = 0; c0 < k; c0 += EKC) {
        const int kc = min(EKC, k - c0);
        if (c0 > 0) __syncthreads();
        for (int e = threadIdx.x; e < kc * ET; e += 256) {
            const int c = e % kc, x = e / kc;                     // contiguous along c in global memory
            ws[c * ELD + x] = (i0 + x < n) ? W[c0 + c + (int64_t)k * (i0 + x)] : 0.0;
            hs[c * ELD + x] = (j0 + x < m) ? H[c0 + c + (int64_t)k * (j0 + x)] : 0.0;
        }
        __syncthreads();
        for (int c = 0; c < kc; c++) {
            double wv[4], hv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { wv[u] = ws[c * ELD + ti + 16 * u]; hv[u] = hs[c * ELD + tj + 16 * u]; }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int v = 0; v < 4; v++) acc[u][v] = fma(wv[u], hv[v], acc[u][v]);
        }
    }
    double s2 = 0.0, skl = 0.0;
#pragma unroll
    for (int v = 0; v < 4; v++) {
        const int64_t j = j0 + tj + 16 * v;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t i = i0 + ti + 16 * u;
            if (i < n && j < m) {
                const double a = static_cast<double>(A[i + n * j]);
                if (!is_missing(a)) {                         // non_missing only: src/nnmf.cpp:124-125
                    const double ah = acc[u][v];
                    const double r = a - ah;
                    s2 = fma(r, r, s2);
                    skl += -(a + TINY_NUM) * log(ah + TINY_NUM) + ah;
                }
            }
        }
    }
    s2 = warp_sum(s2);
    skl = warp_sum(skl);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s2; red[1][threadIdx.x >> 5] = skl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0, a1 = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { a0 += red[0][w]; a1 += red[1][w]; }
        const int64_t b = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        part[2 * b] = a0;
        part[2 * b + 1] = a1;
    }
}

constexpr int STATS_BLOCKS = 296;

// <X, Q> with Q[a, j] = sum over the split-K slots of Qp[slot][j][a]: the cross term of the Gram-identity MSE.
// Fixed work assignment and fixed-order reductions -> reproducible.
__global__ void __launch_bounds__(256)
k_dot_factor_cross(const double* __restrict__ X, const double* __restrict__ Qp, int splits, int64_t total /* k * ncol */,
                   double* __restrict__ part)
{
    __shared__ double red[8];
    double s = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        double q = 0.0;
        for (int sp = 0; sp < splits; sp++) q += Qp[(int64_t)sp * total + e];
        s = fma(X[e], q, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) a += red[w];
        part[blockIdx.x] = a;
    }
}

__global__ void __launch_bounds__(1024)
k_dot_small(const double* __restrict__ a, const double* __restrict__ b, int count, double* __restrict__ out)
{
    __shared__ double sm[32];
    double s = 0.0;
    for (int e = threadIdx.x; e < count; e += 1024) s = fma(a[e], b[e], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = sm[threadIdx.x];
        t = warp_sum(t);
        if (threadIdx.x == 0) out[0] = t;
    }
}

__global__ void __launch_bounds__(256)
k_factor_stats(const double* __restrict__ X, int k, int64_t cols, double* __restrict__ part)
{
    __shared__ double red[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double sq = 0.0, sm = 0.0, cs2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 8 + warp; i < cols; i += (int64_t)gridDim.x * 8) {
        double cs = 0.0;
        for (int a = lane; a < k; a += 32) {
            const double x = X[a + (int64_t)k * i];
            sq = fma(x, x, sq);
            cs += x;
        }
        cs = warp_sum(cs);
        sm += (lane == 0) ? cs : 0.0;
        cs2 += (lane == 0) ? cs * cs : 0.0;
    }
    sq = warp_sum(sq); sm = warp_sum(sm); cs2 = warp_sum(cs2);
    if (lane == 0) { red[0][warp] = sq; red[1][warp] = sm; red[2][warp] = cs2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { a0 += red[0][w]; a1 += red[1][w]; a2 += red[2][w]; }
        part[3 * blockIdx.x] = a0; part[3 * blockIdx.x + 1] = a1; part[3 * blockIdx.x + 2] = a2;
    }
}

}  // namespace

int64_t error_part_count(int64_t n, int64_t m) { return ceil_div(n, ET) * ceil_div(m, ET); }

template <typename TA>
void launch_error(const TA* A, const double* W, const double* H, int k, int64_t n, int64_t m, double* part, double* out,
                  cudaStream_t st)
{
    NNLM_REQUIRE(ceil_div(m, ET) <= 65535, "too many columns for the error kernel grid");
    const size_t smem = sizeof(double) * 2 * (size_t)EKC * ELD;
    auto kern = k_error<TA>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(n, ET), (unsigned)ceil_div(m, ET));
    kern<<<grid, 256, smem, st>>>(A, W, H, k, n, m, part);
    NNLM_LAUNCHED();
    launch_reduce_partials(part, error_part_count(n, m), 2, out, st);
}
template void launch_error<double>(const double*, const double*, const double*, int, int64_t, int64_t, double*, double*, cudaStream_t);
template void launch_error<float>(const float*, const double*, const double*, int, int64_t, int64_t, double*, double*, cudaStream_t);

void launch_dot_factor_cross(const double* X, const double* Qp, int splits, int k, int64_t ncol, double* part, double* out,
                             cudaStream_t st)
{
    k_dot_factor_cross<<<STATS_BLOCKS, 256, 0, st>>>(X, Qp, splits, (int64_t)k * ncol, part);
    NNLM_LAUNCHED();
    launch_reduce_partials(part, STATS_BLOCKS, 1, out, st);
}

void launch_dot_small(const double* a, const double* b, int count, double* out, cudaStream_t st)
{
    k_dot_small<<<1, 1024, 0, st>>>(a, b, count, out);
    NNLM_LAUNCHED();
}

int64_t stats_part_count(int64_t cols) { (void)cols; return STATS_BLOCKS; }

void launch_factor_stats(const double* X, int k, int64_t cols, double* part, double* out, cudaStream_t st)
{
    k_factor_stats<<<STATS_BLOCKS, 256, 0, st>>>(X, k, cols, part);
    NNLM_LAUNCHED();
    launch_reduce_partials(part, STATS_BLOCKS, 3, out, st);
}

}  // namespace nnlm
