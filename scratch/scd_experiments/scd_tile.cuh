// scd_tile.cuh — K3/K4, the throughput-oriented sequential-coordinate-descent solver for the square loss
// (reference src/base_algorithms.cpp:3-37 preceded by mu = WtW*h - WtA (+beta2), src/update_with_missing.cpp:39-41).
//
// The coordinate loop is strictly sequential inside one column (each step sees the mu left by the previous one), but
// columns are independent (src/update_with_missing.cpp:29-30), and every column applies the SAME Gram column at step c:
//     d_j = clamp(h_cj - mu_cj / V_cc) - h_cj ;   mu_:j += d_j * V_:c          for all columns j
// i.e. one step is a rank-1 update of the k x ncol matrix mu. Measured on B200 (scratch/dfma_bench.cu): DFMA issues at
// 64 lanes/clk/SM from registers, but an operand fetched by a warp-wide shared-memory broadcast costs 8 B x 32 lanes of
// register-file return bandwidth, which caps a "one thread = one whole column" layout at ~31 DFMA/clk/SM (every DFMA needs
// its own V value; ncu: fp64 pipe 40 % busy, short-scoreboard stalls on top). The layout here register-blocks the rank-1
// update: a warp owns a tile of mu, T row groups x S column slots (T*S = 32 lanes), each thread holding KQ rows x C columns,
// so one loaded V value feeds C DFMAs.
//   * mu[KQ][C] lives in registers with compile-time indices (the sweep over c is fully unrolled; the per-step code is
//     kept small because instruction fetch becomes the bound beyond ~100 KB of loop body),
//   * h lives in a per-warp shared-memory slab hs[c][column]; only the row group that owns row c touches it,
//   * step c: every lane forms candidate d values from its own row, the owning row group's are the real ones and one
//     shuffle per column hands them to the other row groups; every thread then does KQ x C independent DFMAs,
//   * the division by V_cc is a multiplication with the reciprocal computed once per half-iteration (<= 1 ulp per step),
//   * the reference's `tmp != Hj(k)` branch becomes d = 0 (adding 0 * V leaves mu bit-identical): no divergence.
// Control flow per column is the reference's: a column stops changing when its max relative change drops to rel_tol or at
// max_iter (finished columns are frozen with d = 0 while the rest of the tile keeps sweeping); its sweep count is summed
// into total_raw_iter.
#pragma once
#include <algorithm>

#include "kernels.cuh"

namespace nnlm {
namespace scd_tile {

// resident warps per CTA (1 CTA / SM), bounded by the register file: mu + one Gram column slice per thread
template <int T, int KQ> struct Cfg { static constexpr int WARPS = (T == 2) ? (KQ <= 20 ? 12 : 8) : 16; };

template <int KQ, int T, int C>
__global__ void __launch_bounds__(32 * Cfg<T, KQ>::WARPS, 1)
k_scd_tile(double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ Qp, int splits,
           const uint8_t* __restrict__ mask, int k, int64_t ncol, double l1, unsigned max_iter, double rel_tol,
           unsigned long long* __restrict__ sweeps, unsigned int* __restrict__ next_group)
{
    static_assert(C == 2, "two columns per thread");
    constexpr int WARPS = Cfg<T, KQ>::WARPS;
    constexpr int S = 32 / T;               // column slots per warp
    constexpr int CW = S * C;               // columns per warp tile
    constexpr int KB = KQ * T;              // padded rank (<= 64)
    constexpr int GQ = (KQ + 1) & ~1;       // rows per thread padded to a 16-byte multiple
    extern __shared__ __align__(16) double sm[];
    double* gp = sm;                        // [KB][T][GQ]: gp[(c*T + t)*GQ + i] = V[T*i + t, c] (zero padded)
    double* rinv = gp + KB * T * GQ;        // [KB]
    double* hs = rinv + KB + (threadIdx.x >> 5) * (KB * CW);    // this warp's slab [KB][CW]

    for (int e = threadIdx.x; e < KB * T * GQ; e += 32 * WARPS) {
        const int i = e % GQ, t = (e / GQ) % T, c = e / (GQ * T);
        const int r = T * i + t;
        gp[e] = (i < KQ && r < k && c < k) ? G[r + k * c] : 0.0;
    }
    for (int c = threadIdx.x; c < KB; c += 32 * WARPS) rinv[c] = (c < k) ? 1.0 / G[c + k * c] : 0.0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int s = lane % S, t = lane / S;
    const int64_t ngroups = (ncol + CW - 1) / CW;
    unsigned long long my_sweeps = 0;

    while (true) {
        // column tiles are handed out dynamically: a warp that finishes early takes the next one
        unsigned int g32 = 0;
        if (lane == 0) g32 = atomicAdd(next_group, 1u);
        g32 = __shfl_sync(0xffffffffu, g32, 0);
        const int64_t grp = g32;
        if (grp >= ngroups) break;
        const int64_t col0 = grp * CW;
        const int cnt = (int)min((int64_t)CW, ncol - col0);
        const int total = cnt * k;                 // contiguous doubles of this tile in X / Qp

        // ---- q = sum of split-K partials, staged coalesced; mu = -q (+ l1) ----
        for (int e = lane; e < KB * CW; e += 32) hs[e] = 0.0;
        __syncwarp();
        for (int e = lane; e < total; e += 32) {
            double acc = 0.0;
            for (int sp = 0; sp < splits; sp++) acc += Qp[((int64_t)sp * ncol + col0) * k + e];
            hs[(e % k) * CW + e / k] = acc;
        }
        __syncwarp();
        double mu[KQ][C];
#pragma unroll
        for (int i = 0; i < KQ; i++) {
            const int r = T * i + t;
#pragma unroll
            for (int jj = 0; jj < C; jj++) {
                const bool valid = r < k && (C * s + jj) < cnt;
                mu[i][jj] = valid ? (l1 - hs[r * CW + C * s + jj]) : 0.0;
            }
        }
        __syncwarp();
        // ---- h staged the same way (columns beyond cnt and rows beyond k stay zero) ----
        for (int e = lane; e < KB * CW; e += 32) hs[e] = 0.0;
        __syncwarp();
        for (int e = lane; e < total; e += 32) hs[(e % k) * CW + e / k] = X[col0 * k + e];
        // mbits: coordinates of the column that are masked (never updated); a fully masked column is skipped entirely
        // (src/update_with_missing.cpp:33-34), a column beyond the matrix edge likewise
        unsigned long long mbits[C];
        bool cont[C];
        const unsigned long long kmask = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
#pragma unroll
        for (int jj = 0; jj < C; jj++) {
            mbits[jj] = 0;
            const int cj = C * s + jj;
            if (mask != nullptr && cj < cnt) {
                const uint8_t* mc = mask + (col0 + cj) * k;
                for (int r = 0; r < k; r++) mbits[jj] |= (unsigned long long)(mc[r] != 0) << r;
            }
            cont[jj] = cj < cnt && (mbits[jj] & kmask) != kmask;       // rel_err starts at 1 + rel_tol
        }
        __syncwarp();

        // ---- mu += V h ----
#pragma unroll
        for (int c = 0; c < KB; c++) {
            if (c < k) {
                const double2* gsrc = reinterpret_cast<const double2*>(gp + (c * T + t) * GQ);
                double g[GQ];
#pragma unroll
                for (int i2 = 0; i2 < GQ / 2; i2++) { const double2 v = gsrc[i2]; g[2 * i2] = v.x; g[2 * i2 + 1] = v.y; }
                const double2 h01 = *reinterpret_cast<const double2*>(hs + c * CW + C * s);
#pragma unroll
                for (int i = 0; i < KQ; i++) {
                    mu[i][0] = fma(g[i], h01.x, mu[i][0]);
                    mu[i][1] = fma(g[i], h01.y, mu[i][1]);
                }
            }
        }

        // ---- sweeps ----
        unsigned tcount[C] = {0, 0};
        for (unsigned it = 0; it < max_iter; it++) {
            if (!__any_sync(0xffffffffu, cont[0] || cont[1])) break;
            // coordinates this sweep must leave alone: masked ones, or all of them once the column has converged
            const unsigned long long fz0 = cont[0] ? mbits[0] : ~0ull, fz1 = cont[1] ? mbits[1] : ~0ull;
            bool flag0 = false, flag1 = false;
            // h and 1/V_cc of the coming step are loaded one step ahead (they do not depend on the coordinate chain)
            double2 hnx = *reinterpret_cast<const double2*>(hs + C * s);
            double rnx = rinv[0];
#pragma unroll
            for (int c = 0; c < KB; c++) {
                if (c < k) {
                    constexpr int dummy = 0; (void)dummy;
                    const int to = c % T, io = c / T;                       // owner row group / its local row (compile time)
                    const int ion = ((c + 1) % KB) / T;                     // local row the NEXT step's chain reads
                    const double2* gsrc = reinterpret_cast<const double2*>(gp + (c * T + t) * GQ);
                    double g[GQ];
#pragma unroll
                    for (int i2 = 0; i2 < GQ / 2; i2++) { const double2 v = gsrc[i2]; g[2 * i2] = v.x; g[2 * i2 + 1] = v.y; }
                    double2* hptr = reinterpret_cast<double2*>(hs + c * CW + C * s);
                    const double2 h01 = hnx;
                    const double rc = rnx;
                    const int cn = (c + 1 < k) ? c + 1 : 0;
                    hnx = *reinterpret_cast<const double2*>(hs + cn * CW + C * s);
                    rnx = rinv[cn];
                    // every lane forms candidates from its own row io; only the owner row group's are meaningful
                    const double c0 = fmax(fma(-mu[io][0], rc, h01.x), 0.0);
                    const double c1 = fmax(fma(-mu[io][1], rc, h01.y), 0.0);
                    double d0 = ((fz0 >> c) & 1ull) ? 0.0 : c0 - h01.x;
                    double d1 = ((fz1 >> c) & 1ull) ? 0.0 : c1 - h01.y;
                    const bool own = (t == to);                             // the owner keeps h and the convergence flags
                    if (own) *hptr = make_double2(d0 != 0.0 ? c0 : h01.x, d1 != 0.0 ? c1 : h01.y);
                    flag0 |= own & (2.0 * fabs(d0) > rel_tol * (c0 + h01.x + TINY_NUM));
                    flag1 |= own & (2.0 * fabs(d1) > rel_tol * (c1 + h01.y + TINY_NUM));
                    d0 = __shfl_sync(0xffffffffu, d0, s + S * to);
                    d1 = __shfl_sync(0xffffffffu, d1, s + S * to);
                    // the row the next step's chain depends on is updated first
                    mu[ion][0] = fma(d0, g[ion], mu[ion][0]);
                    mu[ion][1] = fma(d1, g[ion], mu[ion][1]);
#pragma unroll
                    for (int i = 0; i < KQ; i++) {
                        if (i != ion) {
                            mu[i][0] = fma(d0, g[i], mu[i][0]);
                            mu[i][1] = fma(d1, g[i], mu[i][1]);
                        }
                    }
                }
            }
            // a column's flags were raised by whichever row group owned the coordinate: combine across row groups
            unsigned f0 = flag0 ? 1u : 0u, f1 = flag1 ? 1u : 0u;
#pragma unroll
            for (int x = S; x < 32; x <<= 1) { f0 |= __shfl_xor_sync(0xffffffffu, f0, x); f1 |= __shfl_xor_sync(0xffffffffu, f1, x); }
            if (cont[0]) tcount[0]++;
            if (cont[1]) tcount[1]++;
            cont[0] = cont[0] && (f0 != 0u || (0.0 > rel_tol));
            cont[1] = cont[1] && (f1 != 0u || (0.0 > rel_tol));
        }
        if (t == 0) my_sweeps += (unsigned long long)tcount[0] + tcount[1];
        __syncwarp();
        for (int e = lane; e < total; e += 32) X[col0 * k + e] = hs[(e % k) * CW + e / k];
        __syncwarp();
    }
#pragma unroll
    for (int x = 16; x > 0; x >>= 1) my_sweeps += __shfl_xor_sync(0xffffffffu, my_sweeps, x);
    if (lane == 0 && my_sweeps) atomicAdd(sweeps, my_sweeps);
}

template <int KQ, int T, int C>
void launch(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, double l1,
            unsigned max_iter, double rel_tol, unsigned long long* sweeps, unsigned int* counter, cudaStream_t st)
{
    constexpr int WARPS = Cfg<T, KQ>::WARPS;
    constexpr int S = 32 / T, CW = S * C, KB = KQ * T, GQ = (KQ + 1) & ~1;
    const size_t smem = sizeof(double) * ((size_t)KB * T * GQ + KB + (size_t)WARPS * KB * CW);
    auto kern = k_scd_tile<KQ, T, C>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = ceil_div(ncol, CW);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(groups, 148));
    NNLM_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    kern<<<grid, 32 * WARPS, smem, st>>>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter);
    NNLM_LAUNCHED();
}

#define NNLM_SCD_TILE_ARGS double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, \
    double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, unsigned int* counter, cudaStream_t st
#define NNLM_SCD_TILE_PASS X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st

// explicit-instantiation entry points, spread over several translation units to keep the build parallel.
// kq4 = ceil(k / 4): the padded rank is 4*kq4 in both layouts.
void launch_wide_lo(int kq4, NNLM_SCD_TILE_ARGS);    // T = 2 (32-column tiles), 4*kq4 <= 32
void launch_wide_mid(int kq4, NNLM_SCD_TILE_ARGS);   // T = 2, 36..48
void launch_wide_hi(int kq4, NNLM_SCD_TILE_ARGS);    // T = 2, 52..64
void launch_narrow_lo(int kq4, NNLM_SCD_TILE_ARGS);  // T = 4 (16-column tiles), 4*kq4 <= 40
void launch_narrow_hi(int kq4, NNLM_SCD_TILE_ARGS);  // T = 4, 44..64

}  // namespace scd_tile
}  // namespace nnlm
