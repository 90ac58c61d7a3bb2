// solve_ls_missing.cu — K9 + K4/K5 for the NA path of the square loss: reference src/update_with_missing.cpp:58-139.
// For every column j the reference forms its own Gram over the rows where A[:,j] is finite,
//   WtW_j = Wt[:, nm_j] * Wt[:, nm_j]'   (:90; complete columns recompute the full product, :95-96)
// regularises it (:98-103) and runs the same coordinate solver (:107-117). The masked cross-product Wt[:,nm_j]*A[nm_j,j]
// (:91) is the ordinary cross-product of A with its non-finite entries read as zero and comes from the cross kernels.
//
// One CTA per column. The per-column Gram is built in shared memory, in fp64, from whichever index set is smaller:
//   complement  G_j = G_full - sum_{i missing} y_i y_i'      (G_full = unregularised Gram of the whole factor)
//   direct      G_j =          sum_{i present} y_i y_i'
// The index set is compacted in ascending order (ballot + prefix), so the summation order is fixed. Rows y_i are staged
// 64 at a time in shared memory and the symmetric rank-64 update runs as fp64 tensor-core MMAs (DMMA.8x8x4, the full
// fp64 rate with two 8-byte operands per thread): only the 8x8 tiles on and below the diagonal are accumulated
// (nt(nt+1)/2 of them, nt = ceil(k/8), dealt round-robin to the 8 warps as C fragments) and mirrored when the Gram is
// written out. A first version with 4x4 DFMA register tiles fed by LDS.128 ran the per-column Gram at ~1/4 of the fp64
// rate over the full square (config 4: 268 ms per ANLS iteration). Warp 0 then runs warp_solve_ls on the finished Gram.
// "Missing" is bit-exactly the reference's predicate: the entry is not finite (find_finite, :80-83), evaluated on the
// stored value of A (fp64, or fp32 whose non-finite set is identical by construction of the conversion).
#include <algorithm>

#include "kernels.cuh"
#include "solve_core.cuh"

namespace nnlm {

namespace {

constexpr int NT = 256;
constexpr int CH = 64;     // staged rows per step

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <typename TA> __device__ __forceinline__ bool missing_v(TA v);
template <> __device__ __forceinline__ bool missing_v<double>(double v) { return is_missing(v); }
template <> __device__ __forceinline__ bool missing_v<float>(float v) { return ((__float_as_uint(v) >> 23) & 0xffu) == 0xffu; }

template <int RPL, int METHOD, typename TA>
__global__ void __launch_bounds__(NT)
k_solve_ls_missing(double* __restrict__ X, const double* __restrict__ Y, const TA* __restrict__ A,
                   const double* __restrict__ Gfull, const double* __restrict__ Qp, int splits,
                   const uint8_t* __restrict__ mask, int k, int64_t len, int64_t ncol, double p0, double p1, double l1,
                   unsigned max_iter, double rel_tol, unsigned long long* __restrict__ sweeps)
{
    constexpr int KR = 32 * RPL;
    constexpr int KP = KR + 4;                 // pitch of the staged rows: A/B fragment loads hit 32 distinct banks
    constexpr int NTMAX = KR / 8;              // 8x8 tiles per dimension
    constexpr int TPW = (NTMAX * (NTMAX + 1) / 2 + NT / 32 - 1) / (NT / 32);   // lower-triangle tiles per warp
    extern __shared__ __align__(32) double smd[];
    double* gs = smd;                          // [KR][KR] (column-major, leading dimension KR)
    double* ys = gs + KR * KR;                 // [CH][KP] staged rows
    __shared__ int64_t s_idx[CH];
    __shared__ int s_wcnt[NT / 32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    // the tiles of this warp: tile t of the lower triangle is (ta, tb) with t = ta (ta + 1) / 2 + tb, tb <= ta
    const int ntk = (k + 7) / 8, ntiles = ntk * (ntk + 1) / 2;
    int ta_[TPW], tb_[TPW];
#pragma unroll
    for (int j = 0; j < TPW; j++) {
        const int t = warp + (NT / 32) * j;
        int ta = 0;
        while ((ta + 1) * (ta + 2) / 2 <= t) ta++;
        ta_[j] = ta; tb_[j] = t - ta * (ta + 1) / 2;
    }

    for (int64_t col = blockIdx.x; col < ncol; col += gridDim.x) {
        const TA* Aj = A + len * col;
        const uint8_t* mcol = mask ? mask + (int64_t)k * col : nullptr;
        if (mcol) {                                          // src/update_with_missing.cpp:77-78
            int nm = 0;
            for (int c = 0; c < k; c++) nm += mcol[c] != 0;
            if (nm == k) continue;
        }
        // ---- count the missing entries of the column ----
        int cnt = 0;
        for (int64_t i = threadIdx.x; i < len; i += NT) cnt += missing_v<TA>(Aj[i]) ? 1 : 0;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
        __syncthreads();
        if (lane == 0) s_wcnt[warp] = cnt;
        __syncthreads();
        int64_t n_missing = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) n_missing += s_wcnt[w];
        const bool complement = 2 * n_missing <= len;        // subtract the missing rows, or add the present ones
        const double sgn = complement ? -1.0 : 1.0;

        // ---- accumulate the lower-triangle 8x8 tiles as DMMA C fragments ----
        double acc[TPW][2];
#pragma unroll
        for (int j = 0; j < TPW; j++) acc[j][0] = acc[j][1] = 0.0;

        __syncthreads();
        auto flush = [&](int fill) {
            // stage rows y_i for the collected indices (zero beyond `fill` and beyond k), then the rank-`fill` update
            for (int e = threadIdx.x; e < CH * KR; e += NT) {
                const int s = e / KR, r = e % KR;
                ys[s * KP + r] = (s < fill && r < k) ? Y[r + (int64_t)k * s_idx[s]] : 0.0;
            }
            __syncthreads();
            const int steps = (fill + 3) >> 2;
            for (int q = 0; q < steps; q++) {
                const double* row = ys + (4 * q + tig) * KP + gid;
#pragma unroll
                for (int j = 0; j < TPW; j++) {
                    if (warp + (NT / 32) * j < ntiles)
                        dmma(acc[j][0], acc[j][1], row[8 * ta_[j]], row[8 * tb_[j]]);
                }
            }
            __syncthreads();
        };

        if (n_missing > 0) {
            int fill = 0;                                     // entries waiting in s_idx (uniform across the CTA)
            for (int64_t i0 = 0; i0 < len; i0 += NT) {
                const int64_t i = i0 + threadIdx.x;
                const bool miss = (i < len) && missing_v<TA>(Aj[i]);
                const bool take = (i < len) && (miss == complement);
                const unsigned bal = __ballot_sync(0xffffffffu, take);
                if (lane == 0) s_wcnt[warp] = __popc(bal);
                __syncthreads();
                int base = 0, total = 0;
#pragma unroll
                for (int w = 0; w < NT / 32; w++) { if (w < warp) base += s_wcnt[w]; total += s_wcnt[w]; }
                const int vpos = fill + base + __popc(bal & ((1u << lane) - 1u));   // position in the virtual list
                int consumed = 0;
                while (true) {                                // windows of CH entries, ascending index order
                    if (take && vpos >= consumed && vpos < consumed + CH) s_idx[vpos - consumed] = i;
                    __syncthreads();
                    if (fill + total - consumed >= CH) { flush(CH); consumed += CH; }
                    else break;
                }
                fill = fill + total - consumed;
            }
            if (fill > 0) flush(fill);
        }

        // ---- G_j = (complement ? G_full : 0) + sgn * acc, then the reference's regularisation (:98-103) ----
        __syncthreads();
        for (int e = threadIdx.x; e < KR * KR; e += NT) gs[e] = 0.0;      // rows / columns >= 8 * ntk stay zero
        __syncthreads();
#pragma unroll
        for (int j = 0; j < TPW; j++) {
            if (warp + (NT / 32) * j < ntiles) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int a = 8 * ta_[j] + gid, b = 8 * tb_[j] + 2 * tig + e;      // C fragment: row gid, columns 2 tig + e
                    double g = 0.0;
                    if (a < k && b < k) {
                        g = (complement ? Gfull[a + k * b] : 0.0) + sgn * acc[j][e];
                        if (p0 != p1 && a == b) g += p0 - p1;
                        if (p1 != 0.0) g += p1;
                        if (a == b) g += TINY_NUM;
                    }
                    if (ta_[j] != tb_[j] || a >= b) gs[a + KR * b] = g;           // diagonal tiles: keep their lower half ...
                    if (ta_[j] != tb_[j] || a > b) gs[b + KR * a] = g;            // ... and mirror it (the Gram is symmetric)
                }
            }
        }
        __syncthreads();

        // ---- solve (warp 0) ----
        if (warp == 0) {
            double h[RPL], q[RPL];
            unsigned mk[RPL];
#pragma unroll
            for (int s = 0; s < RPL; s++) {
                const int r = lane + 32 * s;
                const bool valid = r < k;
                h[s] = valid ? X[r + (int64_t)k * col] : 0.0;
                double a = 0.0;
                if (valid)
                    for (int sp = 0; sp < splits; sp++) a += Qp[((int64_t)sp * ncol + col) * k + r];
                q[s] = a;
                const bool mb = valid && mcol != nullptr && mcol[r] != 0;
                mk[s] = __ballot_sync(0xffffffffu, mb);
            }
            const unsigned t = warp_solve_ls<RPL, METHOD>(h, q, mk, gs, k, l1, max_iter, rel_tol);
#pragma unroll
            for (int s = 0; s < RPL; s++) {
                const int r = lane + 32 * s;
                if (r < k) X[r + (int64_t)k * col] = h[s];
            }
            if (lane == 0 && t) atomicAdd(sweeps, (unsigned long long)t);
        }
        __syncthreads();
    }
}

template <int RPL, typename TA>
void launch_rpl(int method, double* X, const double* Y, const TA* A, const double* Gfull, const double* Qp, int splits,
                const uint8_t* mask, int k, int64_t len, int64_t ncol, const double* pen, unsigned max_iter, double rel_tol,
                unsigned long long* sweeps, cudaStream_t st)
{
    constexpr int KR = 32 * RPL;
    const size_t smem = sizeof(double) * ((size_t)KR * KR + (size_t)CH * (KR + 4));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ncol, 148 * 8));
    if (method == 1) {
        auto kern = k_solve_ls_missing<RPL, 1, TA>;
        NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NT, smem, st>>>(X, Y, A, Gfull, Qp, splits, mask, k, len, ncol, pen[0], pen[1], pen[2], max_iter, rel_tol, sweeps);
    } else {
        auto kern = k_solve_ls_missing<RPL, 2, TA>;
        NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NT, smem, st>>>(X, Y, A, Gfull, Qp, splits, mask, k, len, ncol, pen[0], pen[1], pen[2], max_iter, rel_tol, sweeps);
    }
    NNLM_LAUNCHED();
}

}  // namespace

template <typename TA>
void launch_solve_ls_missing(int method, double* X, const double* Y, const TA* A, const double* Gfull, const double* Qp,
                             int splits, const uint8_t* mask, int k, int64_t len, int64_t ncol, const double* pen,
                             unsigned max_iter, double rel_tol, unsigned long long* sweeps, cudaStream_t st)
{
    NNLM_REQUIRE(method == 1 || method == 2, "solve_ls_missing handles methods 1 and 2");
    NNLM_REQUIRE(k >= 1 && k <= 128, "rank k must be in [1, 128]");
    if (ncol <= 0) return;
    const int rpl = (k + 31) / 32;
    switch (rpl) {
        case 1: launch_rpl<1, TA>(method, X, Y, A, Gfull, Qp, splits, mask, k, len, ncol, pen, max_iter, rel_tol, sweeps, st); break;
        case 2: launch_rpl<2, TA>(method, X, Y, A, Gfull, Qp, splits, mask, k, len, ncol, pen, max_iter, rel_tol, sweeps, st); break;
        case 3: launch_rpl<3, TA>(method, X, Y, A, Gfull, Qp, splits, mask, k, len, ncol, pen, max_iter, rel_tol, sweeps, st); break;
        default: launch_rpl<4, TA>(method, X, Y, A, Gfull, Qp, splits, mask, k, len, ncol, pen, max_iter, rel_tol, sweeps, st); break;
    }
}
template void launch_solve_ls_missing<double>(int, double*, const double*, const double*, const double*, const double*, int,
                                              const uint8_t*, int, int64_t, int64_t, const double*, unsigned, double,
                                              unsigned long long*, cudaStream_t);
template void launch_solve_ls_missing<float>(int, double*, const double*, const float*, const double*, const double*, int,
                                             const uint8_t*, int, int64_t, int64_t, const double*, unsigned, double,
                                             unsigned long long*, cudaStream_t);

}  // namespace nnlm
