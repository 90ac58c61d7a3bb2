// solve_ls_missing.cu — K9 + K4/K5 for the NA path of the square loss: reference src/update_with_missing.cpp:58-139.
// For every column j the reference forms its own Gram over the rows where A[:,j] is finite,
//   WtW_j = Wt[:, nm_j] * Wt[:, nm_j]'   (:90; complete columns recompute the full product, :95-96)
// regularises it (:98-103) and runs the same coordinate solver (:107-117). The masked cross-product Wt[:,nm_j]*A[nm_j,j]
// (:91) is the ordinary cross-product of A with its non-finite entries read as zero and comes from the cross kernels.
//
// One CTA per column. The per-column Gram is built in fp64 from whichever index set is smaller:
//   complement  G_j = G_full - sum_{i missing} y_i y_i'      (G_full = unregularised Gram of the whole factor)
//   direct      G_j =          sum_{i present} y_i y_i'
// The index set is compacted in ascending order (ballot + prefix), so the summation order is fixed. Rows y_i are staged
// 64 at a time in shared memory and the symmetric rank-64 update runs as fp64 tensor-core MMAs (DMMA.8x8x4, the full
// fp64 rate with two 8-byte operands per thread): only the 8x8 tiles on and below the diagonal are accumulated
// (nt(nt+1)/2 of them, nt = ceil(k/8), dealt round-robin to the 8 warps as C fragments) and mirrored when the Gram is
// written out. A first version with 4x4 DFMA register tiles fed by LDS.128 ran the per-column Gram at ~1/4 of the fp64
// rate over the full square, and solved each column on warp 0 while the other seven warps of the CTA waited (config 4: 268 ms
// per ANLS iteration). The work is now two kernels: A builds the Grams with all warps and writes them to a scratch
// buffer (k*k doubles per column, chunked to ~1 GB), B solves one column per warp, 7 warps per CTA.
// "Missing" is bit-exactly the reference's predicate: the entry is not finite (find_finite, :80-83), evaluated on the
// stored value of A (fp64, or fp32 whose non-finite set is identical by construction of the conversion).
#include <algorithm>
#include <string>
#include <cstdlib>

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

// Kernel A: the regularised per-column Grams of columns [col0, col0 + ncol) -> Gout[j][k*k] (column-major k x k each).
template <int RPL, typename TA>
__global__ void __launch_bounds__(NT, RPL <= 2 ? 3 : 1)
k_gram_missing(const double* __restrict__ Y, const TA* __restrict__ A, const double* __restrict__ Gfull,
               const uint8_t* __restrict__ mask, int k, int64_t len, int64_t col0, int64_t ncol, double p0, double p1,
               double* __restrict__ Gout)
{
    constexpr int KR = 32 * RPL;
    constexpr int KP = KR + 4;                 // pitch of the staged rows: A/B fragment loads hit 32 distinct banks
    constexpr int NTMAX = KR / 8;              // 8x8 tiles per dimension
    constexpr int TPW = (NTMAX * (NTMAX + 1) / 2 + NT / 32 - 1) / (NT / 32);   // lower-triangle tiles per warp
    extern __shared__ __align__(32) double smd[];
    double* ys = smd;                          // [CH][KP] staged rows
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

    for (int64_t cl = blockIdx.x; cl < ncol; cl += gridDim.x) {
        const int64_t col = col0 + cl;
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
            // (all loads of a thread are issued before its first store: a rolled load -> store loop paid the L2 latency
            //  sixteen times per flush and was 28 % of the kernel's samples)
            constexpr int PER = CH * KR / NT;
            double v[PER];
#pragma unroll
            for (int i = 0; i < PER; i++) {
                const int e = threadIdx.x + i * NT, s = e / KR, r = e % KR;
                v[i] = (s < fill && r < k) ? Y[r + (int64_t)k * s_idx[s]] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < PER; i++) {
                const int e = threadIdx.x + i * NT;
                ys[(e / KR) * KP + (e % KR)] = v[i];
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
        double* Gj = Gout + (int64_t)k * k * cl;
#pragma unroll
        for (int j = 0; j < TPW; j++) {
            if (warp + (NT / 32) * j < ntiles) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int a = 8 * ta_[j] + gid, b = 8 * tb_[j] + 2 * tig + e;      // C fragment: row gid, columns 2 tig + e
                    if (a < k && b < k) {
                        double g = (complement ? Gfull[a + k * b] : 0.0) + sgn * acc[j][e];
                        if (p0 != p1 && a == b) g += p0 - p1;
                        if (p1 != 0.0) g += p1;
                        if (a == b) g += TINY_NUM;
                        if (ta_[j] != tb_[j] || a >= b) Gj[a + k * b] = g;            // diagonal tiles: keep their lower half ...
                        if (ta_[j] != tb_[j] || a > b) Gj[b + k * a] = g;             // ... and mirror it (the Gram is symmetric)
                    }
                }
            }
        }
        __syncthreads();
    }
}

// Kernel B: one WARP per column. The warp copies its column's Gram into its own slice of shared memory and runs
// warp_solve_ls (solve_core.cuh); 7 warps per CTA at k <= 64, so 7 columns are in flight per SM instead of the one the
// fused version had while the other seven warps of its CTA waited at a barrier (64 % of all warp samples, ncu).
template <int RPL, int METHOD>
__global__ void __launch_bounds__(256)
k_solve_batch(double* __restrict__ X, const double* __restrict__ Gin, const double* __restrict__ Qp, int splits,
              const uint8_t* __restrict__ mask, int k, int64_t col0, int64_t ncol, int64_t ncol_total, double l1,
              unsigned max_iter, double rel_tol, unsigned long long* __restrict__ sweeps)
{
    constexpr int KR = 32 * RPL;
    extern __shared__ __align__(32) double smd[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    double* gs = smd + (size_t)warp * KR * KR;      // [KR][KR] column-major, rows / columns >= k zero
    unsigned long long my_sweeps = 0;
    for (int e = lane; e < KR * KR; e += 32) gs[e] = 0.0;
    for (int64_t cl = (int64_t)blockIdx.x * wpc + warp; cl < ncol; cl += (int64_t)gridDim.x * wpc) {
        const int64_t col = col0 + cl;
        const uint8_t* mcol = mask ? mask + (int64_t)k * col : nullptr;
        if (mcol) {                                          // src/update_with_missing.cpp:77-78
            int nm = 0;
            for (int c = 0; c < k; c++) nm += mcol[c] != 0;
            if (nm == k) continue;
        }
        __syncwarp();
        const double* Gj = Gin + (int64_t)k * k * cl;
        for (int e = lane; e < k * k; e += 32) gs[(e % k) + KR * (e / k)] = Gj[e];
        __syncwarp();
        double h[RPL], q[RPL];
        unsigned mk[RPL];
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int r = lane + 32 * s;
            const bool valid = r < k;
            h[s] = valid ? X[r + (int64_t)k * col] : 0.0;
            double a = 0.0;
            if (valid)
                for (int sp = 0; sp < splits; sp++) a += Qp[((int64_t)sp * ncol_total + col) * k + r];
            q[s] = a;
            const bool mb = valid && mcol != nullptr && mcol[r] != 0;
            mk[s] = __ballot_sync(0xffffffffu, mb);
        }
        my_sweeps += warp_solve_ls<RPL, METHOD>(h, q, mk, gs, k, l1, max_iter, rel_tol);
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int r = lane + 32 * s;
            if (r < k) X[r + (int64_t)k * col] = h[s];
        }
    }
    if (lane == 0 && my_sweeps) atomicAdd(sweeps, my_sweeps);
}

// Kernel B': as kernel B, with the per-column Gram assembled on the fly from the shared raw Gram and the packed tensor-core
// corrections of na_gram.cu: G_j[a,b] = Gfull[a,b] - S[j][pair(max, min)] + regularisation (src/update_with_missing.cpp:98-103).
template <int RPL, int METHOD, bool TRI>
__global__ void __launch_bounds__(512)
k_solve_batch_packed(double* __restrict__ X, const double* __restrict__ Gfull, const double* __restrict__ S, int64_t pt,
                     const double* __restrict__ Qp, int splits, const double* __restrict__ center, const uint8_t* __restrict__ mask,
                     int k, int64_t ncol, double p0, double p1, double l1, unsigned max_iter, double rel_tol,
                     unsigned long long* __restrict__ sweeps)
{
    extern __shared__ __align__(32) double smd[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int kk2 = k * (k + 1) / 2;
    // TRI: the lower triangle only (solve_core.cuh); else columns packed at pitch k + slack for the lanes beyond row k
    const int per_warp = TRI ? kk2 : k * k + 32 * RPL;
    double* gs = smd + (size_t)warp * per_warp;
    unsigned long long my_sweeps = 0;
    for (int e = lane; e < per_warp; e += 32) gs[e] = 0.0;
    for (int64_t col = (int64_t)blockIdx.x * wpc + warp; col < ncol; col += (int64_t)gridDim.x * wpc) {
        const uint8_t* mcol = mask ? mask + (int64_t)k * col : nullptr;
        if (mcol) {                                          // src/update_with_missing.cpp:77-78
            int nm = 0;
            for (int c = 0; c < k; c++) nm += mcol[c] != 0;
            if (nm == k) continue;
        }
        __syncwarp();
        const double* Sj = S + pt * col;
        // lower triangle from the packed corrections (contiguous reads), mirrored: G_j = Gfull - S_j + regularisation (:98-103)
        for (int p = lane; p < kk2; p += 32) {
            int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
            while ((a + 1) * (a + 2) / 2 <= p) a++;
            while (a * (a + 1) / 2 > p) a--;
            const int b = p - a * (a + 1) / 2;
            double g = Gfull[a + k * b] - Sj[p];
            if (p0 != p1 && a == b) g += p0 - p1;
            if (p1 != 0.0) g += p1;
            if (a == b) g += TINY_NUM;
            if (TRI) gs[p] = g;
            else { gs[a + k * b] = g; gs[b + k * a] = g; }
        }
        __syncwarp();
        double h[RPL], q[RPL];
        unsigned mk[RPL];
        const double cj = center ? center[col] : 0.0;
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int r = lane + 32 * s;
            const bool valid = r < k;
            h[s] = valid ? X[r + (int64_t)k * col] : 0.0;
            double a = 0.0;
            if (valid) {
                for (int sp = 0; sp < splits; sp++) a += Qp[((int64_t)sp * ncol + col) * k + r];
                if (center) a = fma(-cj, Sj[kk2 + r], a);   // the planes held A - c_j with missing = 0: remove c_j * (masked row sum)
            }
            q[s] = a;
            const bool mb = valid && mcol != nullptr && mcol[r] != 0;
            mk[s] = __ballot_sync(0xffffffffu, mb);
        }
        my_sweeps += warp_solve_ls<RPL, METHOD, TRI>(h, q, mk, gs, k, l1, max_iter, rel_tol, k);
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int r = lane + 32 * s;
            if (r < k) X[r + (int64_t)k * col] = h[s];
        }
    }
    if (lane == 0 && my_sweeps) atomicAdd(sweeps, my_sweeps);
}

template <int RPL>
void launch_packed_rpl(int method, double* X, const double* Gfull, const double* S, int64_t pt, const double* Qp, int splits,
                       const double* center, const uint8_t* mask, int k, int64_t ncol, const double* pen, unsigned max_iter,
                       double rel_tol, unsigned long long* sweeps, cudaStream_t st)
{
    // the per-column Gram is kept as its lower triangle (16 instead of 10 columns in flight per SM at k = 50);
    // NNLM_NA_SOLVER=square selects the full square at pitch k (experiments)
    static const bool square = [] { const char* e = getenv("NNLM_NA_SOLVER"); return e && std::string(e) == "square"; }();
    const size_t per_warp = sizeof(double) * (square ? (size_t)k * k + 32 * RPL : (size_t)k * (k + 1) / 2);
    const int wpc = std::max(1, std::min(16, (int)(220 * 1024 / per_warp)));      // columns in flight per SM
    const size_t smem = (size_t)wpc * per_warp;
    auto kb = square ? (method == 1 ? k_solve_batch_packed<RPL, 1, false> : k_solve_batch_packed<RPL, 2, false>)
                     : (method == 1 ? k_solve_batch_packed<RPL, 1, true> : k_solve_batch_packed<RPL, 2, true>);
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(ncol, wpc), 148));
    kb<<<grid, 32 * wpc, smem, st>>>(X, Gfull, S, pt, Qp, splits, center, mask, k, ncol, pen[0], pen[1], pen[2], max_iter, rel_tol, sweeps);
    NNLM_LAUNCHED();
}

template <int RPL, typename TA>
void launch_rpl(int method, double* X, const double* Y, const TA* A, const double* Gfull, const double* Qp, int splits,
                const uint8_t* mask, int k, int64_t len, int64_t ncol, const double* pen, unsigned max_iter, double rel_tol,
                unsigned long long* sweeps, cudaStream_t st)
{
    constexpr int KR = 32 * RPL;
    // per-column Grams go through a stream-ordered scratch buffer of at most ~1 GB: columns are processed in chunks
    static const int64_t budget = [] { const char* e = getenv("NNLM_NA_SCRATCH_DOUBLES"); return e ? atoll(e) : ((int64_t)1 << 27); }();
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(ncol, budget / ((int64_t)k * k)));
    double* Gout = nullptr;
    pool_setup_once();
    NNLM_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&Gout), sizeof(double) * (size_t)chunk * k * k, st));
    const size_t smem_a = sizeof(double) * ((size_t)CH * (KR + 4));
    const int wpc = std::max(1, std::min(7, (int)(220 * 1024 / (sizeof(double) * KR * KR))));   // solver warps per CTA
    const size_t smem_b = sizeof(double) * (size_t)wpc * KR * KR;
    auto ka = k_gram_missing<RPL, TA>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    auto kb = method == 1 ? k_solve_batch<RPL, 1> : k_solve_batch<RPL, 2>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    for (int64_t c0 = 0; c0 < ncol; c0 += chunk) {
        const int64_t nc = std::min<int64_t>(chunk, ncol - c0);
        const int grid_a = (int)std::max<int64_t>(1, std::min<int64_t>(nc, 148 * 8));
        ka<<<grid_a, NT, smem_a, st>>>(Y, A, Gfull, mask, k, len, c0, nc, pen[0], pen[1], Gout);
        NNLM_LAUNCHED();
        const int grid_b = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nc, wpc), 148 * 4));
        kb<<<grid_b, 32 * wpc, smem_b, st>>>(X, Gout, Qp, splits, mask, k, c0, nc, ncol, pen[2], max_iter, rel_tol, sweeps);
        NNLM_LAUNCHED();
    }
    NNLM_CUDA_CHECK(cudaFreeAsync(Gout, st));
}

}  // namespace

void launch_solve_ls_missing_packed(int method, double* X, const double* Gfull, const double* S, int64_t pt, const double* Qp,
                                    int splits, const double* center, const uint8_t* mask, int k, int64_t ncol, const double* pen,
                                    unsigned max_iter, double rel_tol, unsigned long long* sweeps, cudaStream_t st)
{
    NNLM_REQUIRE(method == 1 || method == 2, "solve_ls_missing handles methods 1 and 2");
    NNLM_REQUIRE(k >= 1 && k <= 128, "rank k must be in [1, 128]");
    if (ncol <= 0) return;
    switch ((k + 31) / 32) {
        case 1: launch_packed_rpl<1>(method, X, Gfull, S, pt, Qp, splits, center, mask, k, ncol, pen, max_iter, rel_tol, sweeps, st); break;
        case 2: launch_packed_rpl<2>(method, X, Gfull, S, pt, Qp, splits, center, mask, k, ncol, pen, max_iter, rel_tol, sweeps, st); break;
        case 3: launch_packed_rpl<3>(method, X, Gfull, S, pt, Qp, splits, center, mask, k, ncol, pen, max_iter, rel_tol, sweeps, st); break;
        default: launch_packed_rpl<4>(method, X, Gfull, S, pt, Qp, splits, center, mask, k, ncol, pen, max_iter, rel_tol, sweeps, st); break;
    }
}

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
