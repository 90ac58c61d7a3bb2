// solve_kl_fast.cu — host side of the cluster KL solver (kernel and design notes: solve_kl_fast.cuh)
#include "solve_kl_fast.cuh"

namespace nnlm {

using namespace klf;

namespace {
// fp64 k x len (column-major) -> fp32 [k][len] row-major copy of the fixed factor
__global__ void __launch_bounds__(256)
k_factor_rows_f32(const double* __restrict__ Y, int k, int64_t len, float* __restrict__ out)
{
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    for (int c0 = 0; c0 < k; c0 += 32) {
        for (int r = ty; r < 32; r += 8) {
            const int64_t i = i0 + r;
            const int c = c0 + tx;
            tile[r][tx] = (i < len && c < k) ? (float)Y[c + (int64_t)k * i] : 0.0f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int c = c0 + r;
            const int64_t i = i0 + tx;
            if (c < k && i < len) out[(int64_t)c * len + i] = tile[tx][r];
        }
        __syncthreads();
    }
}

}  // namespace

bool solve_kl_fast_supported(int k, int64_t len)
{
    KlfShape sh;
    return k >= 1 && k <= 256 && klf_shape(len, &sh) && klf_smem(k, sh.E) <= 220 * 1024;
}

void launch_factor_rows_f32(const double* Y, int k, int64_t len, float* out, cudaStream_t st)
{
    if (len <= 0) return;
    k_factor_rows_f32<<<(unsigned)ceil_div(len, 32), 256, 0, st>>>(Y, k, len, out);
    NNLM_LAUNCHED();
}

void launch_solve_kl_fast(int method, double* X, const float* Y32, const float* A, const float* WH0, const double* sumY,
                          const uint8_t* mask, int k,
                          int64_t len, int64_t ncol, const double* pen, unsigned max_iter, double rel_tol,
                          unsigned long long* sweeps, cudaStream_t st)
{
    NNLM_REQUIRE(method == 3 || method == 4, "solve_kl_fast handles methods 3 and 4");
    KlfShape sh;
    NNLM_REQUIRE(solve_kl_fast_supported(k, len) && klf_shape(len, &sh), "solve_kl_fast: shape not supported");
    if (ncol <= 0) return;
    if (method == 3) (sh.E <= 8 ? launch_m3_lo : launch_m3_hi)(NNLM_KLF_PASS);
    else             (sh.E <= 8 ? launch_m4_lo : launch_m4_hi)(NNLM_KLF_PASS);
}

}  // namespace nnlm
