// ingest.cu — upload-time passes over A: precision conversion, the transposed copy that replaces the reference's
// per-iteration `A.t()` (src/nnmf.cpp:117,131), missing-entry detection (src/nnmf.cpp:64-69) and the constant part
// of the KL distance (src/nnmf.cpp:70-73). All HBM-bound streaming kernels: coalesced 32x32 tiles through shared memory.
#include "kernels.cuh"

namespace nnlm {

namespace {

// conversion of a stored element: finite doubles beyond the float range saturate instead of becoming infinite, so the
// set of non-finite (= missing, src/update_with_missing.cpp:80-83) entries is identical in every storage type
template <typename TOut> __device__ __forceinline__ TOut store_cast(double a);
template <> __device__ __forceinline__ double store_cast<double>(double a) { return a; }
template <> __device__ __forceinline__ float store_cast<float>(double a) {
    if (!is_missing(a) && fabs(a) > 3.4028234663852886e38) return copysignf(3.4028234663852886e38f, (float)a);
    return static_cast<float>(a);
}

constexpr int TILE = 32;
constexpr int ROWS = 8;

template <typename TOut>
__global__ void __launch_bounds__(TILE * ROWS)
k_ingest(const double* __restrict__ src, int64_t len, int64_t ncol, int64_t j0, int64_t jc,
         TOut* __restrict__ dst_cm, TOut* __restrict__ dst_rm, double* __restrict__ part)
{
    __shared__ double tile[TILE][TILE + 1];
    __shared__ double red[3][ROWS];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * TILE, jj0 = (int64_t)blockIdx.y * TILE;
    double kl = 0.0, miss = 0.0, sq = 0.0;
    const int64_t i = i0 + tx;
#pragma unroll
    for (int r = ty; r < TILE; r += ROWS) {
        const int64_t jj = jj0 + r;
        double a = 0.0;
        if (i < len && jj < jc) {
            a = src[i + len * jj];
            if (is_missing(a)) miss += 1.0;
            else { kl += (a + TINY_NUM) * log(a + TINY_NUM) - a; sq = fma(a, a, sq); }
            if (dst_cm) dst_cm[i + len * (j0 + jj)] = store_cast<TOut>(a);
        }
        tile[r][tx] = a;
    }
    __syncthreads();
    if (dst_rm) {
        const int64_t jj = jj0 + tx;
#pragma unroll
        for (int r = ty; r < TILE; r += ROWS) {
            const int64_t ii = i0 + r;
            if (ii < len && jj < jc) dst_rm[(j0 + jj) + ncol * ii] = store_cast<TOut>(tile[tx][r]);
        }
    }
    kl = warp_sum(kl);
    miss = warp_sum(miss);
    sq = warp_sum(sq);
    if (tx == 0) { red[0][ty] = kl; red[1][ty] = miss; red[2][ty] = sq; }
    __syncthreads();
    if (tx == 0 && ty == 0) {
        double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int r = 0; r < ROWS; r++) { s0 += red[0][r]; s1 += red[1][r]; s2 += red[2][r]; }
        const int64_t b = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        part[3 * b] = s0;
        part[3 * b + 1] = s1;
        part[3 * b + 2] = s2;
    }
}

// One block, fixed association order -> bit-reproducible totals regardless of how the producers were scheduled.
__global__ void __launch_bounds__(1024)
k_reduce_partials(const double* __restrict__ part, int64_t count, int width, double* __restrict__ out)
{
    __shared__ double sm[32];
    for (int w = 0; w < width; w++) {
        double s = 0.0;
        for (int64_t e = threadIdx.x; e < count; e += blockDim.x) s += part[e * width + w];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            double t = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
            t = warp_sum(t);
            if (threadIdx.x == 0) out[w] = t;
        }
        __syncthreads();
    }
}

// bit(i + len*j) = !isfinite(A[i + len*j]); one warp per 32 consecutive linear indices via ballot.
__global__ void k_na_bits(const double* __restrict__ A, int64_t total, uint32_t* __restrict__ bits)
{
    const int64_t words = (total + 31) / 32;
    const int lane = threadIdx.x & 31;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < words; w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t e = w * 32 + lane;
        const bool miss = (e < total) && is_missing(A[e]);
        const unsigned b = __ballot_sync(0xffffffffu, miss);
        if (lane == 0) bits[w] = b;
    }
}

__global__ void k_col_missing(const double* __restrict__ A, int64_t len, int64_t ncol, int64_t* __restrict__ cnt)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < ncol; j += nwarp) {
        int c = 0;
        for (int64_t i = lane; i < len; i += 32) c += is_missing(A[i + len * j]) ? 1 : 0;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
        if (lane == 0) cnt[j] = c;
    }
}

__global__ void __launch_bounds__(TILE * ROWS)
k_transpose_d(const double* __restrict__ in, int64_t rows, int64_t cols, double* __restrict__ out)
{
    __shared__ double tile[TILE][TILE + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t r0 = (int64_t)blockIdx.x * TILE, c0 = (int64_t)blockIdx.y * TILE;
    for (int r = ty; r < TILE; r += ROWS) {
        const int64_t rr = r0 + tx, cc = c0 + r;
        tile[r][tx] = (rr < rows && cc < cols) ? in[rr + rows * cc] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < TILE; r += ROWS) {
        const int64_t cc = c0 + tx, rr = r0 + r;
        if (rr < rows && cc < cols) out[cc + cols * rr] = tile[tx][r];
    }
}

__global__ void k_mask_u8_t(const int32_t* __restrict__ in, int64_t rows, int64_t cols, uint8_t* __restrict__ out)
{
    // small (n x K) logical matrix -> K x n bytes; not on the hot path
    const int64_t total = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = e % cols, r = e / cols;       // e indexes the output (c fastest)
        out[e] = in[r + rows * c] != 0 ? 1 : 0;
    }
}

__global__ void k_mask_u8(const int32_t* __restrict__ in, int64_t total, uint8_t* __restrict__ out)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = in[e] != 0 ? 1 : 0;
}

inline int grid_for(int64_t work, int per_block, int cap = 148 * 16)
{
    int64_t g = ceil_div(work, per_block);
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

}  // namespace

int64_t ingest_part_count(int64_t len, int64_t jc) { return ceil_div(len, TILE) * ceil_div(jc, TILE); }

template <typename TOut>
void launch_ingest(const double* src, int64_t len, int64_t ncol, int64_t j0, int64_t jc, TOut* dst_cm, TOut* dst_rm,
                   double* part, cudaStream_t st)
{
    if (len <= 0 || jc <= 0) return;
    NNLM_REQUIRE(ceil_div(jc, TILE) <= 65535, "ingest chunk has too many columns");
    dim3 grid((unsigned)ceil_div(len, TILE), (unsigned)ceil_div(jc, TILE)), block(TILE, ROWS);
    k_ingest<TOut><<<grid, block, 0, st>>>(src, len, ncol, j0, jc, dst_cm, dst_rm, part);
    NNLM_LAUNCHED();
}
template void launch_ingest<double>(const double*, int64_t, int64_t, int64_t, int64_t, double*, double*, double*, cudaStream_t);
template void launch_ingest<float>(const double*, int64_t, int64_t, int64_t, int64_t, float*, float*, double*, cudaStream_t);

void launch_reduce_partials(const double* part, int64_t count, int width, double* out, cudaStream_t st)
{
    k_reduce_partials<<<1, 1024, 0, st>>>(part, count, width, out);
    NNLM_LAUNCHED();
}

void launch_na_bits(const double* A, int64_t len, int64_t ncol, uint32_t* bits, int64_t* col_missing, cudaStream_t st)
{
    const int64_t total = len * ncol;
    if (total <= 0) return;
    k_na_bits<<<grid_for(total, 256), 256, 0, st>>>(A, total, bits);
    NNLM_LAUNCHED();
    if (col_missing) {
        k_col_missing<<<grid_for(ncol * 32, 256), 256, 0, st>>>(A, len, ncol, col_missing);
        NNLM_LAUNCHED();
    }
}

void launch_transpose_d(const double* in, int64_t rows, int64_t cols, double* out, cudaStream_t st)
{
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((unsigned)ceil_div(rows, TILE), (unsigned)ceil_div(cols, TILE)), block(TILE, ROWS);
    k_transpose_d<<<grid, block, 0, st>>>(in, rows, cols, out);
    NNLM_LAUNCHED();
}

void launch_mask_to_u8_t(const int32_t* in, int64_t rows, int64_t cols, uint8_t* out_t, cudaStream_t st)
{
    if (rows * cols <= 0) return;
    k_mask_u8_t<<<grid_for(rows * cols, 256), 256, 0, st>>>(in, rows, cols, out_t);
    NNLM_LAUNCHED();
}

void launch_mask_to_u8(const int32_t* in, int64_t count, uint8_t* out, cudaStream_t st)
{
    if (count <= 0) return;
    k_mask_u8<<<grid_for(count, 256), 256, 0, st>>>(in, count, out);
    NNLM_LAUNCHED();
}

}  // namespace nnlm
