// solve_ls.cu — K3/K4/K5: the per-column non-negative least-squares solvers for the square loss, dense A (one shared
// regularised Gram for all columns, src/update_with_missing.cpp:17-25). One warp per column (columns are independent
// given Wt: src/update_with_missing.cpp:29-30); the solver itself is warp_solve_ls in solve_core.cuh.
#include <algorithm>

#include "kernels.cuh"
#include "solve_core.cuh"

namespace nnlm {

namespace {

constexpr int WARPS = 8;

template <int RPL, int METHOD>
__global__ void __launch_bounds__(32 * WARPS)
k_solve_ls(double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ Qp, int splits,
           const uint8_t* __restrict__ mask, int k, int64_t ncol, double l1, unsigned max_iter, double rel_tol,
           unsigned long long* __restrict__ sweeps)
{
    constexpr int KR = 32 * RPL;
    extern __shared__ double gs[];   // [k][KR], gs[r + KR*c] = V[r, c], rows >= k are zero
    for (int e = threadIdx.x; e < k * KR; e += 32 * WARPS) {
        const int c = e / KR, r = e % KR;
        gs[e] = (r < k) ? G[r + k * c] : 0.0;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int64_t nwarp = (int64_t)gridDim.x * WARPS;
    unsigned long long my_sweeps = 0;

    for (int64_t col = warp0; col < ncol; col += nwarp) {
        double h[RPL], q[RPL];
        unsigned mk[RPL];
        int n_masked = 0;
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int r = lane + 32 * s;
            const bool valid = r < k;
            h[s] = valid ? X[r + (int64_t)k * col] : 0.0;
            double acc = 0.0;
            if (valid)
                for (int sp = 0; sp < splits; sp++) acc += Qp[((int64_t)sp * ncol + col) * k + r];
            q[s] = acc;
            const bool mb = valid && mask != nullptr && mask[r + (int64_t)k * col] != 0;
            mk[s] = __ballot_sync(0xffffffffu, mb);
            n_masked += __popc(mk[s]);
        }
        if (n_masked == k) continue;                        // src/update_with_missing.cpp:33-34
        my_sweeps += warp_solve_ls<RPL, METHOD>(h, q, mk, gs, k, l1, max_iter, rel_tol);
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int r = lane + 32 * s;
            if (r < k) X[r + (int64_t)k * col] = h[s];
        }
    }
    if (lane == 0 && my_sweeps) atomicAdd(sweeps, my_sweeps);
}

template <int RPL>
void launch_rpl(int method, double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k,
                int64_t ncol, double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, cudaStream_t st)
{
    const size_t smem = sizeof(double) * (size_t)k * 32 * RPL;
    int64_t blocks = ceil_div(ncol, WARPS);
    const int64_t per_sm = std::max<int64_t>(1, std::min<int64_t>(8, (int64_t)(200 * 1024) / (int64_t)std::max<size_t>(smem, 1)));
    blocks = std::min<int64_t>(blocks, 148 * per_sm);
    if (method == 1) {
        auto kern = k_solve_ls<RPL, 1>;
        NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, 32 * WARPS, smem, st>>>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps);
    } else {
        auto kern = k_solve_ls<RPL, 2>;
        NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, 32 * WARPS, smem, st>>>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps);
    }
    NNLM_LAUNCHED();
}

}  // namespace

void launch_solve_ls(int method, double* X, const double* G, const double* Qp, int splits, const uint8_t* mask,
                     int k, int64_t ncol, double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps,
                     cudaStream_t st)
{
    NNLM_REQUIRE(method == 1 || method == 2, "solve_ls handles methods 1 and 2");
    NNLM_REQUIRE(k >= 1 && k <= 128, "rank k must be in [1, 128]");
    if (ncol <= 0) return;
    const int rpl = (k + 31) / 32;
    switch (rpl) {
        case 1: launch_rpl<1>(method, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
        case 2: launch_rpl<2>(method, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
        case 3: launch_rpl<3>(method, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
        default: launch_rpl<4>(method, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
    }
}

}  // namespace nnlm
