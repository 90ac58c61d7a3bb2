// scd_dmma.cuh — K3/K4, the sequential-coordinate-descent solver for the square loss, blocked so that the bulk of the
// arithmetic runs as fp64 tensor-core MMAs (reference src/base_algorithms.cpp:3-37 preceded by mu = WtW*h - WtA (+beta2),
// src/update_with_missing.cpp:39-41).
//
// One coordinate step applies, to every column j of the tile:   d_j = clamp(h_cj - mu_cj / V_cc) - h_cj ;  mu_:j += d_j V_:c
// The step is sequential in c, but the value it reads, mu_cj, only depends on the d of EARLIER coordinates. Coordinates are
// therefore processed in blocks of 8 (= one 8-row tile of mu):
//   (1) the 8 x 8 diagonal part runs exactly as in the reference — 8 sequential steps per column, one thread per column,
//       all operands in registers, no branches, no stores: the dependent chain per step is
//       DFMA (candidate) -> integer clamp -> DADD (d) -> integer freeze-mask -> DFMA (next row of the tile), ~40 cycles;
//   (2) the other k-8 rows of mu receive the block's eight rank-1 updates at once, mu_rest += V[rest, blk] * D[blk, :],
//       as DMMA.8x8x4 instructions on the C fragments that hold mu. Only the NEXT diagonal tile needs them right away;
//       the rest is deferred and interleaved with the next block's sequential steps, where it fills the bubbles of the
//       dependent chain (a first version that ran (1) and (2) back to back was latency-bound: 2576 cycles per block per
//       scheduler against 768 cycles of MMA pipe time).
// Measured on B200 (scratch/dmma_bench.cu, dfma_bench*.cu, lat_bench.cu): DMMA.8x8x4 sustains 64 FMA/clk/SM — the full
// fp64 rate — with one 8-byte A, one 8-byte B and a 16-byte C operand per thread, whereas a DFMA stream fed by shared-memory
// broadcasts is capped near 31 FMA/clk/SM by register-file return bandwidth and needs 8x the instructions; and a
// one-thread-per-column solver whose every step issues k DFMAs after a vote/branch runs at ~600 cycles per step (ncu:
// fixed-latency "wait" and short-scoreboard stalls on top, fp64 pipe 40 % busy).
// Arithmetic differences from the reference, all <= 1 ulp per operation: 1/V_cc is a reciprocal computed once per
// half-iteration; the `tmp != Hj(k)` branch becomes d = 0 (adding 0 * V leaves mu bit-identical); rows outside the block
// see the block's updates summed 4 at a time inside the MMA instead of one FMA after another.
// Control flow per column is the reference's: a column stops changing when its max relative change drops to rel_tol or at
// max_iter (finished columns are frozen while the rest of the tile keeps sweeping); sweep counts are summed into total_raw_iter.
#pragma once
#include <algorithm>

#include "kernels.cuh"

namespace nnlm {
namespace scd_dmma {

// resident warps per CTA (one CTA per SM): narrower tiles hold fewer accumulators per thread
template <int NB, int CT> struct Cfg { static constexpr int WARPS = (NB > 8) ? 8 : ((CT == 4) ? 8 : 12); };

// D(8x8) += A(8x4, row-major) * B(4x8, col-major): a = A[lane>>2][lane&3], b = B[lane&3][lane>>2], c = C[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// max(x, +0) on the bit pattern: negative values (and -0) become +0; cheaper on the dependent chain than DSETP + FSEL
__device__ __forceinline__ double clamp0(double x)
{
    int hi = __double2hiint(x), lo = __double2loint(x);
    const int keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, lo & keep);
}
__device__ __forceinline__ double and_mask(double x, int m)
{
    return __hiloint2double(__double2hiint(x) & m, __double2loint(x) & m);
}

template <int NB, int CT>   // NB blocks of 8 coordinates (padded rank KB = 8*NB), CT column tiles of 8 (NC = 8*CT columns per warp)
__global__ void __launch_bounds__(32 * Cfg<NB, CT>::WARPS, 1)
k_scd_dmma(double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ Qp, int splits,
           const uint8_t* __restrict__ mask, int k, int64_t ncol, double l1, unsigned max_iter, double rel_tol,
           unsigned long long* __restrict__ sweeps, unsigned int* __restrict__ next_group)
{
    constexpr int KB = 8 * NB, NC = 8 * CT, KS = KB + 4, WARPS = Cfg<NB, CT>::WARPS;
    extern __shared__ __align__(16) double sm[];
    double* gc = sm;                                   // [KB][KS]: gc[c*KS + r] = V[r, c] (symmetric; zero padded)
    double* rinv = gc + KB * KS;                       // [KB] 1 / V[c,c] (0 for padding)
    double* wbase = rinv + KB + (threadIdx.x >> 5) * (KB * NC + 24 * NC);
    double* hs = wbase;                                // [KB][NC] current h of the tile's columns
    double* dsm = hs + KB * NC;                        // [2][8][NC] d of the current and of the previous block
    double* tsm = dsm + 16 * NC;                       // [8][NC]  transposition buffer (fragment <-> thread-per-column)

    for (int e = threadIdx.x; e < KB * KS; e += 32 * WARPS) {
        const int c = e / KS, r = e % KS;
        gc[e] = (r < k && c < k) ? G[r + k * c] : 0.0;
    }
    for (int c = threadIdx.x; c < KB; c += 32 * WARPS) rinv[c] = (c < k) ? 1.0 / G[c + k * c] : 0.0;
    __syncthreads();

    const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const int colx = lane < NC ? lane : NC - 1;        // the column this lane owns in the sequential part
    const int64_t ngroups = (ncol + NC - 1) / NC;
    unsigned long long my_sweeps = 0;

    while (true) {
        unsigned int g32 = 0;
        if (lane == 0) g32 = atomicAdd(next_group, 1u);     // tiles are handed out dynamically
        g32 = __shfl_sync(0xffffffffu, g32, 0);
        const int64_t grp = g32;
        if (grp >= ngroups) break;
        const int64_t col0 = grp * NC;
        const int cnt = (int)min((int64_t)NC, ncol - col0);
        const int total = cnt * k;

        // ---- stage h (rows >= k and columns >= cnt read as zero) ----
        for (int e = lane; e < KB * NC; e += 32) hs[e] = 0.0;
        __syncwarp();
        for (int e = lane; e < total; e += 32) hs[(e % k) * NC + e / k] = X[col0 * k + e];
        // ---- mu = l1 - q in C-fragment layout ----
        double mu[NB][CT][2];
#pragma unroll
        for (int rt = 0; rt < NB; rt++)
#pragma unroll
            for (int ct = 0; ct < CT; ct++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int r = 8 * rt + gid, c = 8 * ct + 2 * tig + e;
                    double q = 0.0;
                    if (r < k && c < cnt)
                        for (int sp = 0; sp < splits; sp++) q += Qp[((int64_t)sp * ncol + col0 + c) * k + r];
                    mu[rt][ct][e] = (r < k && c < cnt) ? l1 - q : 0.0;
                }
        // ---- per-column state of the sequential part (lane = column) ----
        constexpr int MW = (KB + 63) / 64;                           // 64-bit words of the per-column coordinate mask
        unsigned long long mbits[MW];
        const bool have = lane < cnt;
        bool any_free = false;
#pragma unroll
        for (int w = 0; w < MW; w++) {
            mbits[w] = 0;
            if (mask != nullptr && have) {
                const uint8_t* mc = mask + (col0 + lane) * k;
                for (int r = 64 * w; r < k && r < 64 * w + 64; r++) mbits[w] |= (unsigned long long)(mc[r] != 0) << (r - 64 * w);
            }
            const int kw = k - 64 * w;                               // coordinates of this word that exist
            const unsigned long long kmask = kw >= 64 ? ~0ull : (kw <= 0 ? 0ull : ((1ull << kw) - 1ull));
            any_free = any_free || (mbits[w] & kmask) != kmask;
            mbits[w] |= ~kmask;                                      // padding coordinates are never updated
        }
        bool cont = have && any_free;                                // fully masked column: src/update_with_missing.cpp:33-34
        __syncwarp();

        // ---- mu += V h : the same block update with D := h, including the diagonal tile ----
#pragma unroll
        for (int b = 0; b < NB; b++) {
#pragma unroll
            for (int kh = 0; kh < 2; kh++) {
                double bf[CT];
#pragma unroll
                for (int ct = 0; ct < CT; ct++) bf[ct] = hs[(8 * b + 4 * kh + tig) * NC + 8 * ct + gid];
#pragma unroll
                for (int rt = 0; rt < NB; rt++) {
                    const double a = gc[(8 * b + 4 * kh + tig) * KS + 8 * rt + gid];
#pragma unroll
                    for (int ct = 0; ct < CT; ct++) dmma(mu[rt][ct][0], mu[rt][ct][1], a, bf[ct]);
                }
            }
        }

        // ---- sweeps ----
        // Pipeline per block b:  P1 diagonal tile -> one thread per column;  P2 the 8 sequential steps, interleaved with the
        // DEFERRED MMAs of the previous block (rows outside both diagonal tiles: independent registers, they fill the bubbles
        // of the dependent chain);  P3 publish d, return the tile, and apply block b's update to the NEXT diagonal tile at
        // once (it is the only part the next chain waits for).
        for (int e = lane; e < 16 * NC; e += 32) dsm[e] = 0.0;       // "previous block" of the very first block: d = 0
        __syncwarp();
        unsigned t = 0;
        int cur = 0;                                                 // dsm buffer of the block being processed
        for (unsigned it = 0; it < max_iter; it++) {
            if (!__any_sync(0xffffffffu, cont)) break;
            unsigned long long fz[MW];                                // coordinates this sweep must leave alone
#pragma unroll
            for (int w = 0; w < MW; w++) fz[w] = cont ? mbits[w] : ~0ull;
            bool flag = false;
#pragma unroll
            for (int b = 0; b < NB; b++) {
                constexpr int dummy = 0; (void)dummy;
                const int pb = (b + NB - 1) % NB, nb = (b + 1) % NB;  // previous / next block (compile time)
                const double* dprev = dsm + (cur ^ 1) * 8 * NC;
                double* dcur = dsm + cur * 8 * NC;
                // P1: diagonal tile: fragments -> one thread per column
#pragma unroll
                for (int ct = 0; ct < CT; ct++)
                    *reinterpret_cast<double2*>(tsm + gid * NC + 8 * ct + 2 * tig) = make_double2(mu[b][ct][0], mu[b][ct][1]);
                __syncwarp();
                double m8[8], h8[8], dd[8];
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    m8[r] = tsm[r * NC + colx];
                    h8[r] = hs[(8 * b + r) * NC + colx];
                }
                double bfp[2][CT];                                   // B fragments of the previous block's d
#pragma unroll
                for (int kh = 0; kh < 2; kh++)
#pragma unroll
                    for (int ct = 0; ct < CT; ct++) bfp[kh][ct] = dprev[(4 * kh + tig) * NC + 8 * ct + gid];
                // P2: eight sequential coordinate steps (registers only) + deferred MMAs of block pb
                int slot = 0;                                        // deferred (kh, rt) pairs are dealt round-robin to the steps
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int cc = 8 * b + c;
                    const double hc = h8[c];
                    const double cand = clamp0(fma(-m8[c], rinv[cc], hc));
                    const int live = ((fz[cc >> 6] >> (cc & 63)) & 1ull) ? 0 : -1;
                    const double d = and_mask(cand - hc, live);
                    dd[c] = d;
                    h8[c] = live ? cand : hc;
                    const double2* gcol = reinterpret_cast<const double2*>(gc + cc * KS + 8 * b);
                    const double2 g01 = gcol[0], g23 = gcol[1], g45 = gcol[2], g67 = gcol[3];
                    const double gv[8] = {g01.x, g01.y, g23.x, g23.y, g45.x, g45.y, g67.x, g67.y};
                    if (c + 1 < 8) m8[c + 1] = fma(d, gv[c + 1], m8[c + 1]);      // the row the next step reads goes first
#pragma unroll
                    for (int r = 0; r < 8; r++)
                        if (r != c + 1) m8[r] = fma(d, gv[r], m8[r]);
                    flag = flag || (2.0 * fabs(d) > rel_tol * (cand + hc + TINY_NUM));
#pragma unroll
                    for (int kh = 0; kh < 2; kh++)
#pragma unroll
                        for (int rt = 0; rt < NB; rt++) {
                            if (rt != pb && rt != b) {
                                if ((slot++ & 7) == c) {
                                    const double a = gc[(8 * pb + 4 * kh + tig) * KS + 8 * rt + gid];
#pragma unroll
                                    for (int ct = 0; ct < CT; ct++) dmma(mu[rt][ct][0], mu[rt][ct][1], a, bfp[kh][ct]);
                                }
                            }
                        }
                    slot = 0;
                }
                // P3: publish d and the new h, return the diagonal tile to the fragments
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (lane < NC) {
                        dcur[c * NC + lane] = dd[c];
                        tsm[c * NC + lane] = m8[c];
                        if (dd[c] != 0.0) hs[(8 * b + c) * NC + lane] = h8[c];
                    }
                }
                __syncwarp();
#pragma unroll
                for (int ct = 0; ct < CT; ct++) {
                    const double2 v = *reinterpret_cast<const double2*>(tsm + gid * NC + 8 * ct + 2 * tig);
                    mu[b][ct][0] = v.x; mu[b][ct][1] = v.y;
                }
                if (NB > 1) {
                    // the next diagonal tile takes block b's update now (the next chain reads it)
#pragma unroll
                    for (int kh = 0; kh < 2; kh++) {
                        const double a = gc[(8 * b + 4 * kh + tig) * KS + 8 * nb + gid];
#pragma unroll
                        for (int ct = 0; ct < CT; ct++)
                            dmma(mu[nb][ct][0], mu[nb][ct][1], a, dcur[(4 * kh + tig) * NC + 8 * ct + gid]);
                    }
                }
                __syncwarp();
                cur ^= 1;
            }
            if (cont) t++;
            cont = cont && (flag || (0.0 > rel_tol));
        }
        if (have) my_sweeps += t;
        __syncwarp();
        for (int e = lane; e < total; e += 32) X[col0 * k + e] = hs[(e % k) * NC + e / k];
        __syncwarp();
    }
#pragma unroll
    for (int x = 16; x > 0; x >>= 1) my_sweeps += __shfl_xor_sync(0xffffffffu, my_sweeps, x);
    if (lane == 0 && my_sweeps) atomicAdd(sweeps, my_sweeps);
}

template <int NB, int CT>
void launch(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, double l1,
            unsigned max_iter, double rel_tol, unsigned long long* sweeps, unsigned int* counter, cudaStream_t st)
{
    constexpr int KB = 8 * NB, NC = 8 * CT, KS = KB + 4, WARPS = Cfg<NB, CT>::WARPS;
    const size_t smem = sizeof(double) * ((size_t)KB * KS + KB + (size_t)WARPS * (KB * NC + 24 * NC));
    auto kern = k_scd_dmma<NB, CT>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = ceil_div(ncol, NC);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(groups, 148));
    NNLM_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    kern<<<grid, 32 * WARPS, smem, st>>>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter);
    NNLM_LAUNCHED();
}

#define NNLM_SCD_ARGS double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, \
    double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, unsigned int* counter, cudaStream_t st
#define NNLM_SCD_PASS X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st

// explicit-instantiation entry points (several translation units keep the build parallel); nb = ceil(k / 8)
void launch_ct1_big_a(int nb, NNLM_SCD_ARGS);  //  8-column tiles, 9 <= nb <= 12
void launch_ct1_big_b(int nb, NNLM_SCD_ARGS);  //  8-column tiles, 13 <= nb <= 16 (k <= 128)

}  // namespace scd_dmma
}  // namespace nnlm
