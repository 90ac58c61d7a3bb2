// cross_simt.cu — K2 on CUDA cores in fp64: the cross-product  Q = Wt * A  (k x ncol) that the reference forms one
// column at a time as `Wt * A.col(j)` (src/update_with_missing.cpp:39,45). This is the "exact" path: every product and
// every accumulation is an IEEE double FMA, so the result differs from the reference's BLAS dgemv only by summation
// order. It is used for small problems (where the run is launch-latency-bound anyway) and as the full-size
// cross-check of the tcgen05 path.
//
// Mapping: a CTA owns TJ = 32*CJ*NWJ columns of A and a slice of the contraction index (split-K, partials summed in a
// fixed order by the solver). Warp (wa, wj): rows a in [16*wa, 16*wa+16) x columns {lane + 32*(CJ*wj + c)}.
// The Wt tile is read from shared memory as warp-wide broadcasts (all lanes of a warp share the same rows), the A
// tile as one conflict-free element per lane per column slot, so the inner loop is bound by the DFMA pipe.
#include <algorithm>

#include "kernels.cuh"

namespace nnlm {

namespace {

constexpr int RA = 16;     // rows of the output per thread
constexpr int CJ = 2;      // column slots per thread
constexpr int NWJ = 2;     // warps along columns
constexpr int TJ = 32 * CJ * NWJ;   // 128 columns per CTA
constexpr int TI = 32;     // contraction indices per stage

template <typename TA, int NWA>
__global__ void __launch_bounds__(32 * NWJ * NWA)
k_cross_simt(const double* __restrict__ Y, const TA* __restrict__ A, int k, int64_t len, int64_t ncol,
             int64_t per_split, double* __restrict__ Qp)
{
    constexpr int KP = RA * NWA;
    constexpr int NT = 32 * NWJ * NWA;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* ys = reinterpret_cast<double*>(smem_raw);          // [TI][KP]
    double* as = ys + TI * KP;                                  // [TI][TJ + 1]: odd stride -> conflict-free column stores
    constexpr int AS_LD = TJ + 1;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wj = warp % NWJ, wa = warp / NWJ;
    const int64_t j0 = (int64_t)blockIdx.x * TJ;
    const int64_t i_beg = (int64_t)blockIdx.y * per_split;
    const int64_t i_end = min(len, i_beg + per_split);

    double acc[RA][CJ];
#pragma unroll
    for (int r = 0; r < RA; r++)
#pragma unroll
        for (int c = 0; c < CJ; c++) acc[r][c] = 0.0;

    for (int64_t i0 = i_beg; i0 < i_end; i0 += TI) {
        const int cnt = (int)min((int64_t)TI, i_end - i0);
        // Wt tile: k*cnt contiguous doubles starting at Y + k*i0
        for (int e = threadIdx.x; e < TI * KP; e += NT) {
            const int ii = e / KP, a = e % KP;
            ys[e] = (ii < cnt && a < k) ? Y[a + (int64_t)k * (i0 + ii)] : 0.0;
        }
        // A tile: lane <-> contraction index (coalesced along the contiguous column), warps stride over columns
        for (int jj = warp; jj < TJ; jj += NT / 32) {
            const int64_t j = j0 + jj;
            double v = 0.0;
            if (lane < cnt && j < ncol) {
                v = static_cast<double>(A[(i0 + lane) + len * j]);
                if (is_missing(v)) v = 0.0;      // NA path: the masked cross-product of src/update_with_missing.cpp:91
            }
            as[lane * AS_LD + jj] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int ii = 0; ii < TI; ii++) {
            double yv[RA], av[CJ];
            const double2* yrow = reinterpret_cast<const double2*>(ys + ii * KP + RA * wa);
#pragma unroll
            for (int r = 0; r < RA / 2; r++) { const double2 t = yrow[r]; yv[2 * r] = t.x; yv[2 * r + 1] = t.y; }
#pragma unroll
            for (int c = 0; c < CJ; c++) av[c] = as[ii * AS_LD + 32 * (CJ * wj + c) + lane];
#pragma unroll
            for (int r = 0; r < RA; r++)
#pragma unroll
                for (int c = 0; c < CJ; c++) acc[r][c] = fma(yv[r], av[c], acc[r][c]);
        }
        __syncthreads();
    }
    double* out = Qp + (int64_t)blockIdx.y * k * ncol;
#pragma unroll
    for (int c = 0; c < CJ; c++) {
        const int64_t j = j0 + 32 * (CJ * wj + c) + lane;
        if (j < ncol) {
#pragma unroll
            for (int r = 0; r < RA; r++) {
                const int a = RA * wa + r;
                if (a < k) out[a + (int64_t)k * j] = acc[r][c];
            }
        }
    }
}

template <typename TA, int NWA>
void launch_one(const double* Y, const TA* A, int k, int64_t len, int64_t ncol, int splits, double* Qp, cudaStream_t st)
{
    constexpr int KP = RA * NWA;
    const size_t smem = sizeof(double) * (TI * KP + TI * (TJ + 1));
    auto kern = k_cross_simt<TA, NWA>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t per_split = ceil_div(ceil_div(len, splits), TI) * TI;
    dim3 grid((unsigned)ceil_div(ncol, TJ), (unsigned)splits);
    kern<<<grid, 32 * NWJ * NWA, smem, st>>>(Y, A, k, len, ncol, per_split, Qp);
    NNLM_LAUNCHED();
}

}  // namespace

int cross_simt_splits(int k, int64_t len, int64_t ncol)
{
    (void)k;
    const int64_t tiles = ceil_div(ncol, TJ);
    int64_t s = ceil_div(2 * 148 * 2, tiles);            // aim at >= 2 waves of 2 CTAs/SM
    const int64_t max_s = std::max<int64_t>(1, len / (8 * TI));
    if (s > max_s) s = max_s;
    if (s > 64) s = 64;
    if (s < 1) s = 1;
    return (int)s;
}

template <typename TA>
void launch_cross_simt(const double* Y, const TA* A, int k, int64_t len, int64_t ncol, int splits, double* Qp, cudaStream_t st)
{
    NNLM_REQUIRE(k >= 1 && k <= 256, "rank k must be in [1, 256]");
    if (ncol <= 0) return;
    const int nwa = (k + RA - 1) / RA;
    switch (nwa) {
        case 1: launch_one<TA, 1>(Y, A, k, len, ncol, splits, Qp, st); break;
        case 2: launch_one<TA, 2>(Y, A, k, len, ncol, splits, Qp, st); break;
        case 3: launch_one<TA, 3>(Y, A, k, len, ncol, splits, Qp, st); break;
        case 4: launch_one<TA, 4>(Y, A, k, len, ncol, splits, Qp, st); break;
        case 5: case 6: launch_one<TA, 6>(Y, A, k, len, ncol, splits, Qp, st); break;
        case 7: case 8: launch_one<TA, 8>(Y, A, k, len, ncol, splits, Qp, st); break;
        default: {
            // k in (128, 256]: two passes over row blocks of 128 (rare: nnlm with many predictors)
            NNLM_REQUIRE(false, "exact cross-product supports k <= 128");
        }
    }
}
template void launch_cross_simt<double>(const double*, const double*, int, int64_t, int64_t, int, double*, cudaStream_t);
template void launch_cross_simt<float>(const double*, const float*, int, int64_t, int64_t, int, double*, cudaStream_t);

}  // namespace nnlm
