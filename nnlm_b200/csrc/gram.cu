// gram.cu — K1/K1r/K6: the shared Gram matrix WtW = Wt * Wt' of a half-iteration with the reference's
// regularisation (src/update_with_missing.cpp:19-24), and sumW = rowSums(Wt) for the KL methods (:27).
// fp64 on CUDA cores: 2*len*k^2 flop (2.5e8 at n=50000, k=50) is noise next to the cross-product, but it is the
// quantity the whole half-iteration is conditioned on, so it is kept in double and reduced in a fixed order.
#include "kernels.cuh"

namespace nnlm {

namespace {

constexpr int GRAM_TI = 16;          // contraction indices staged per step
constexpr int GRAM_MAX_SPLITS = 296; // 2 CTAs per SM

// Each CTA owns a contiguous slice of i and one (16*KT)x(16*KT) output block (a0, b0) of the Gram, register-tiled
// KT x KT per thread. k <= 128 is a single block; larger k tiles the output over gridDim.y/z.
template <int KT>
__global__ void __launch_bounds__(256)
k_gram_partial(const double* __restrict__ Y, int k, int64_t len, int64_t per_split, double* __restrict__ part)
{
    constexpr int KP = 16 * KT;
    __shared__ double ysa[GRAM_TI][KP];
    __shared__ double ysb[GRAM_TI][KP];
    const int ta = threadIdx.x & 15, tb = threadIdx.x >> 4;
    const int a0 = blockIdx.y * KP, b0 = blockIdx.z * KP;
    const int64_t i_beg = (int64_t)blockIdx.x * per_split;
    const int64_t i_end = min(len, i_beg + per_split);
    double acc[KT][KT];
#pragma unroll
    for (int u = 0; u < KT; u++)
#pragma unroll
        for (int v = 0; v < KT; v++) acc[u][v] = 0.0;

    for (int64_t i0 = i_beg; i0 < i_end; i0 += GRAM_TI) {
        const int cnt = (int)min((int64_t)GRAM_TI, i_end - i0);
        for (int e = threadIdx.x; e < GRAM_TI * KP; e += 256) {
            const int ii = e / KP, a = e % KP;
            const bool in = ii < cnt;
            ysa[ii][a] = (in && a0 + a < k) ? Y[a0 + a + (int64_t)k * (i0 + ii)] : 0.0;
            ysb[ii][a] = (in && b0 + a < k) ? Y[b0 + a + (int64_t)k * (i0 + ii)] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int ii = 0; ii < GRAM_TI; ii++) {
            double ya[KT], yb[KT];
#pragma unroll
            for (int u = 0; u < KT; u++) { ya[u] = ysa[ii][ta + 16 * u]; yb[u] = ysb[ii][tb + 16 * u]; }
#pragma unroll
            for (int u = 0; u < KT; u++)
#pragma unroll
                for (int v = 0; v < KT; v++) acc[u][v] = fma(ya[u], yb[v], acc[u][v]);
        }
        __syncthreads();
    }
    double* out = part + (int64_t)blockIdx.x * k * k;
#pragma unroll
    for (int u = 0; u < KT; u++)
#pragma unroll
        for (int v = 0; v < KT; v++) {
            const int a = a0 + ta + 16 * u, b = b0 + tb + 16 * v;
            if (a < k && b < k) out[a + k * b] = acc[u][v];
        }
}

// Fixed-order sum of the split partials + the reference's regularisation:
//   diag += beta0 - beta1 (if they differ); all += beta1 (if non-zero); diag += TINY_NUM   (:20-24)
// One warp per Gram element: lanes take the split partials in a strided, fixed order and a butterfly combines them
// (bit-reproducible; the first version walked all splits in one thread and cost 59 us per half-iteration).
__global__ void __launch_bounds__(256)
k_gram_finish(const double* __restrict__ part, int splits, int k, double b0, double b1, int raw, double* __restrict__ G)
{
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= k * k) return;
    double s = 0.0;
    for (int sp = lane; sp < splits; sp += 32) s += part[(int64_t)sp * k * k + e];
    s = warp_sum(s);
    if (lane != 0) return;
    const bool diag = (e / k) == (e % k);
    if (raw) { G[e] = s; return; }
    if (b0 != b1 && diag) s += b0 - b1;
    if (b1 != 0.0) s += b1;
    if (diag) s += TINY_NUM;
    G[e] = s;
}

__global__ void __launch_bounds__(256)
k_rowsum_partial(const double* __restrict__ Y, int k, int64_t len, int64_t per_split, double* __restrict__ part)
{
    // thread t owns row a = t % k and every (256/k)-th index of the slice; combined through shared memory in fixed order
    extern __shared__ double sm[];
    const int groups = 256 / k;            // k <= 256
    const int a = threadIdx.x % k, g = threadIdx.x / k;
    const int64_t i_beg = (int64_t)blockIdx.x * per_split;
    const int64_t i_end = min(len, i_beg + per_split);
    double s = 0.0;
    if (g < groups)
        for (int64_t i = i_beg + g; i < i_end; i += groups) s += Y[a + (int64_t)k * i];
    sm[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < k) {
        double t = 0.0;
        for (int gg = 0; gg < groups; gg++) t += sm[threadIdx.x + gg * k];
        part[(int64_t)blockIdx.x * k + threadIdx.x] = t;
    }
}

__global__ void k_rowsum_finish(const double* __restrict__ part, int splits, int k, double* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // one warp per row
    if (a >= k) return;
    double s = 0.0;
    for (int sp = lane; sp < splits; sp += 32) s += part[(int64_t)sp * k + a];
    s = warp_sum(s);
    if (lane == 0) out[a] = s;
}

}  // namespace

int gram_splits(int64_t len)
{
    int64_t s = ceil_div(len, 4 * GRAM_TI);
    if (s < 1) s = 1;
    if (s > GRAM_MAX_SPLITS) s = GRAM_MAX_SPLITS;
    return (int)s;
}

void launch_gram(const double* Y, int k, int64_t len, const double* pen, double* part, double* G, cudaStream_t st)
{
    NNLM_REQUIRE(k >= 1 && k <= 256, "rank k must be in [1, 256]");
    const int splits = gram_splits(len);
    const int64_t per_split = ceil_div(ceil_div(len, splits), GRAM_TI) * GRAM_TI;
    const int kt = (k + 15) / 16;
    if (kt <= 1)      k_gram_partial<1><<<splits, 256, 0, st>>>(Y, k, len, per_split, part);
    else if (kt <= 2) k_gram_partial<2><<<splits, 256, 0, st>>>(Y, k, len, per_split, part);
    else if (kt <= 4) k_gram_partial<4><<<splits, 256, 0, st>>>(Y, k, len, per_split, part);
    else {
        const unsigned nb = (unsigned)((k + 127) / 128);
        k_gram_partial<8><<<dim3(splits, nb, nb), 256, 0, st>>>(Y, k, len, per_split, part);
    }
    NNLM_LAUNCHED();
    k_gram_finish<<<(k * k + 7) / 8, 256, 0, st>>>(part, splits, k, pen ? pen[0] : 0.0, pen ? pen[1] : 0.0, pen ? 0 : 1, G);
    NNLM_LAUNCHED();
}

__global__ void k_gram_regularise(const double* __restrict__ Graw, int k, double b0, double b1, double* __restrict__ G)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= k * k) return;
    double s = Graw[e];
    const bool diag = (e / k) == (e % k);
    if (b0 != b1 && diag) s += b0 - b1;
    if (b1 != 0.0) s += b1;
    if (diag) s += TINY_NUM;
    G[e] = s;
}

void launch_gram_regularise(const double* Graw, int k, const double* pen, double* G, cudaStream_t st)
{
    k_gram_regularise<<<(k * k + 255) / 256, 256, 0, st>>>(Graw, k, pen[0], pen[1], G);
    NNLM_LAUNCHED();
}

void launch_rowsum(const double* Y, int k, int64_t len, double* part, double* out, cudaStream_t st)
{
    NNLM_REQUIRE(k >= 1 && k <= 256, "rank k must be in [1, 256]");
    const int splits = gram_splits(len);
    const int64_t per_split = ceil_div(len, splits);
    k_rowsum_partial<<<splits, 256, 256 * sizeof(double), st>>>(Y, k, len, per_split, part);
    NNLM_LAUNCHED();
    k_rowsum_finish<<<(k + 7) / 8, 256, 0, st>>>(part, splits, k, out);
    NNLM_LAUNCHED();
}

}  // namespace nnlm
