// na_gram.cu — K9 on the tensor cores: the per-column masked Grams of the NA path (reference
// src/update_with_missing.cpp:88-96: `WtW = Wt.cols(nm) * Wt.cols(nm).t()` for every column j) as ONE dense contraction.
//
// With M[i,j] = 1 where A[i,j] is missing (not finite, :80-83) and y_i the i-th column of the fixed factor,
//     G_j = sum_{i present} y_i y_i' = G_full - sum_i M[i,j] y_i y_i'
// so the correction of all columns is  S = M' Z  with the Khatri-Rao self-product Z[i,(a,b)] = y_i[a] y_i[b], a >= b
// (k(k+1)/2 columns), extended by the k columns Z[i,a] = y_i[a] whose contraction gives the masked row sums
// sum_{i missing} y_i[a] (used to correct the centring term of the masked cross-product). SURVEY.md §2.1 K9 calls this
// "the only truly tensor-bound kernel" of the package.
//
// Precision. The solver is conditioned on G_j, so the contraction must be exact to far below 1e-5: each column of Z is
// written as FOUR 11-bit fixed-point slices against a power-of-two bound of that column (Ozaki-style splitting),
//     z = sign * (s0 2^33 + s1 2^22 + s2 2^11 + s3) * 2^(e - 44),  0 <= s < 2048,   2^e > max_i |y[a,i]| * max_i |y[b,i]|,
// stored as fp16 (integers up to 2048 are exact). The mask is 0/1. Every product is then exact, and an fp32 TMEM
// accumulator holds the exact sum of up to 2^13 of them; the cross-product kernel drains it into fp64 every 4096 indices
// (cross_tc.cu MODE 1). The only error is the 2^-44 truncation of Z relative to the column bound: measured against the
// fp64 per-column Gram in tests/test_gpu_na_path.py.
// Cost per half-iteration at config 4: ceil(1325 / 128) = 11 tiles of 128 Z-columns x 2 slice pairs = 22 passes of the
// HBM-bound cross-product kernel over the fp16 mask plane (two slices share one pass: cross_tc.cu MODE 2).
#include <algorithm>

#include "kernels.cuh"

namespace nnlm {

namespace {

constexpr int ZS = NA_SLICES;      // slices per Z column
static_assert(NA_SLICES % 2 == 0, "slices go through the contraction in pairs");
constexpr int ZT = NA_TILE;        // Z columns (factor-plane rows) per contraction

__device__ __forceinline__ int pair_a_of(int p)      // p = a (a + 1) / 2 + b, b <= a
{
    int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while ((a + 1) * (a + 2) / 2 <= p) a++;
    while (a * (a + 1) / 2 > p) a--;
    return a;
}

// Z planes of one tile of ZT columns: plane[s][row][i], row-major along i with pitch ld. rowmax: bit patterns of max_i |Y[a,i]|.
__global__ void __launch_bounds__(256)
k_z_slices(const double* __restrict__ Y, int k, int64_t len, int64_t ld, const unsigned long long* __restrict__ rowmax, int p0,
           int npairs, __half* __restrict__ planes, double* __restrict__ unscale /* [ZS][ZT] */)
{
    extern __shared__ double ys[];          // [64][k + 1]
    __shared__ int s_a[ZT], s_b[ZT];
    __shared__ double s_scale[ZT];          // 2^(44 - e)
    const int64_t i0 = (int64_t)blockIdx.x * 64;
    const int cnt = (int)min((int64_t)64, len - i0);
    const int kp = k + 1;
    for (int e = threadIdx.x; e < cnt * k; e += 256) ys[(e / k) * kp + (e % k)] = Y[(int64_t)k * i0 + e];
    const int kk2 = k * (k + 1) / 2;
    for (int r = threadIdx.x; r < ZT; r += 256) {
        const int p = p0 + r;
        int a = -1, b = -1;
        double bound = 0.0;
        if (p < kk2) {
            a = pair_a_of(p); b = p - a * (a + 1) / 2;
            bound = __longlong_as_double((long long)rowmax[a]) * __longlong_as_double((long long)rowmax[b]);
        } else if (p < npairs) {
            a = p - kk2;                                                // linear column: z = y[a]
            bound = __longlong_as_double((long long)rowmax[a]);
        }
        int ex = 0;
        if (bound > 0.0) { frexp(bound, &ex); ex += 1; }               // 2^ex > bound (one extra bit: the product bound is rounded)
        s_a[r] = a; s_b[r] = b;
        s_scale[r] = bound > 0.0 ? ldexp(1.0, 44 - ex) : 0.0;
        if (blockIdx.x == 0)
            for (int s = 0; s < ZS; s++) unscale[s * ZT + r] = bound > 0.0 ? ldexp(1.0, ex - 11 * (s + 1)) : 0.0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ZT * 64; e += 256) {
        const int r = e / 64, ii = e % 64;
        if (i0 + ii >= ld) continue;
        double sl[ZS] = {0.0, 0.0, 0.0, 0.0};
        const int a = s_a[r];
        if (a >= 0 && ii < cnt) {
            double z = ys[ii * kp + a];
            if (s_b[r] >= 0) z *= ys[ii * kp + s_b[r]];
            const double sg = z < 0 ? -1.0 : 1.0;
            double t = floor(fabs(z) * s_scale[r]);                    // integer below 2^44, exact in fp64
            sl[0] = floor(t * (1.0 / 8589934592.0)); t -= sl[0] * 8589934592.0;      // 2^33
            sl[1] = floor(t * (1.0 / 4194304.0));    t -= sl[1] * 4194304.0;         // 2^22
            sl[2] = floor(t * (1.0 / 2048.0));       t -= sl[2] * 2048.0;            // 2^11
            sl[3] = t;
#pragma unroll
            for (int s = 0; s < ZS; s++) sl[s] *= sg;
        }
#pragma unroll
        for (int s = 0; s < ZS; s++) planes[((int64_t)s * ZT + r) * ld + i0 + ii] = __float2half_rn((float)sl[s]);
    }
}

// S[j][p0 + a] += sum over slots of Qp[slot][j][a]
__global__ void __launch_bounds__(256)
k_fold(const double* __restrict__ Qp, int slots, int64_t ncol, int64_t pt, int p0, double* __restrict__ S)
{
    const int64_t total = ncol * ZT;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const int64_t j = e / ZT;
        const int a = (int)(e % ZT);
        double s = 0.0;
        for (int sp = 0; sp < slots; sp++) s += Qp[(int64_t)sp * total + e];
        S[j * pt + p0 + a] += s;
    }
}

// fp16 0/1 planes of the missing indicator: a_plane[j][i] (pitch ld_a) and t_plane[i][j] (pitch ld_t)
template <typename TA>
__global__ void __launch_bounds__(256)
k_mask_planes(const TA* __restrict__ A, int64_t len, int64_t ncol, __half* __restrict__ a_plane, int64_t ld_a,
              __half* __restrict__ t_plane, int64_t ld_t)
{
    __shared__ unsigned char tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
    const __half one = __float2half_rn(1.0f), zero = __float2half_rn(0.0f);
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int64_t i = i0 + tx, j = j0 + r;
        unsigned char miss = 0;
        if (i < len && j < ncol) {
            const double a = static_cast<double>(A[i + len * j]);
            miss = is_missing(a) ? 1 : 0;
            if (a_plane) a_plane[i + ld_a * j] = miss ? one : zero;
        }
        tile[r][tx] = miss;
    }
    __syncthreads();
    if (t_plane) {
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t j = j0 + tx, i = i0 + r;
            if (i < len && j < ncol) t_plane[j + ld_t * i] = tile[tx][r] ? one : zero;
        }
    }
}

}  // namespace

int na_pair_count(int k) { return k * (k + 1) / 2 + k; }
int64_t na_packed_width(int k) { return ceil_div(na_pair_count(k), ZT) * ZT; }

template <typename TA>
void launch_mask_planes(const TA* A, int64_t len, int64_t ncol, __half* a_plane, int64_t ld_a, __half* t_plane, int64_t ld_t,
                        cudaStream_t st)
{
    NNLM_REQUIRE(ceil_div(ncol, 32) <= 65535, "too many columns for the mask plane grid");
    dim3 grid((unsigned)ceil_div(len, 32), (unsigned)ceil_div(ncol, 32));
    k_mask_planes<TA><<<grid, 256, 0, st>>>(A, len, ncol, a_plane, ld_a, t_plane, ld_t);
    NNLM_LAUNCHED();
}
template void launch_mask_planes<double>(const double*, int64_t, int64_t, __half*, int64_t, __half*, int64_t, cudaStream_t);
template void launch_mask_planes<float>(const float*, int64_t, int64_t, __half*, int64_t, __half*, int64_t, cudaStream_t);

// S (ncol x na_packed_width(k), fp64, zeroed here) = for every column j the packed lower triangle of sum_{i missing} y_i y_i'
// followed by the k masked row sums sum_{i missing} y_i. mask_plane: [ncol][ld] fp16 0/1 over the contraction index.
// zplanes: scratch of NA_SLICES * NA_TILE * ld halves; zunscale: NA_SLICES * NA_TILE doubles; Qp: plan.slots * ncol * NA_TILE.
void launch_na_gram_tc(const CrossPlan& plan, const double* Y, int k, const __half* mask_plane, unsigned long long* rowmax,
                       __half* zplanes, double* zunscale, double* Qp, double* S, cudaStream_t st)
{
    const int64_t len = plan.len, ncol = plan.ncol, ld = plan.ld_f, pt = na_packed_width(k);
    const int npairs = na_pair_count(k);
    NNLM_CUDA_CHECK(cudaMemsetAsync(S, 0, sizeof(double) * (size_t)ncol * pt, st));
    launch_rowmax(Y, k, len, rowmax, st);
    const size_t smem = sizeof(double) * 64 * (k + 1);
    if (smem > 48 * 1024) NNLM_CUDA_CHECK(cudaFuncSetAttribute(k_z_slices, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int p0 = 0; p0 < npairs; p0 += ZT) {
        k_z_slices<<<(unsigned)ceil_div(ld, 64), 256, smem, st>>>(Y, k, len, ld, rowmax, p0, npairs, zplanes, zunscale);
        NNLM_LAUNCHED();
        for (int s = 0; s < ZS; s += 2) {              // two slices per pass over the mask plane
            if (plan.pairs)
                launch_mask_tc2(plan, mask_plane, zplanes + (size_t)s * ZT * ld, zplanes + (size_t)(s + 1) * ZT * ld, zunscale + s * ZT, Qp, st);
            else
                launch_cross_tc_exact2(plan, mask_plane, zplanes + (size_t)s * ZT * ld, zplanes + (size_t)(s + 1) * ZT * ld,
                                       zunscale + s * ZT, Qp, st);
            k_fold<<<(int)std::min<int64_t>(ceil_div(ncol * ZT, 256), 148 * 16), 256, 0, st>>>(Qp, plan.slots, ncol, pt, p0, S);
            NNLM_LAUNCHED();
        }
    }
}

}  // namespace nnlm
