// scd_chain.cuh — K3/K4, second generation of the blocked sequential-coordinate-descent solver for the square loss
// (reference src/base_algorithms.cpp:3-37 preceded by mu = WtW*h - WtA (+beta2), src/update_with_missing.cpp:39-41).
//
// Same blocking as scd_dmma.cuh (coordinates in blocks of 8; mu = V h - q of a tile of columns lives in DMMA C fragments),
// restructured after an ncu source-level capture of that kernel (profiles/r1_m_scd_stalls.md): its 8-step dependent chain
// ran at ~250 cycles per step because every step waited on shared-memory broadcasts of V, a second dependent chain (the
// short-circuit convergence test) shared the in-order issue slot, and the chain's DFMAs queued behind other warps' DMMAs.
// Here:
//   * the chain never touches mu. At block entry P_r = h_r - mu'_r / V_rr are formed for the 8 coordinates, where the MMA
//     side maintains mu' = mu - L h with L the strictly lower triangle of the block's own 8x8 diagonal tile of V (the Gram
//     copy the MMAs read has those entries zeroed), so P_r already holds + (V_rc / V_rr) h_c(old) for the earlier
//     coordinates c of the block. Step c reads cand = P_c, clamps it to the new value hn_c and folds it into the later
//     candidates with ONE fma per row, P_r -= (V_rc / V_rr) hn_c (r > c, multipliers precomputed per half-iteration).
//     Dependent path per step: DFMA -> sign test -> select -> DFMA (round 1 had DFMA -> DADD -> select -> select -> DFMA:
//     two fp64-pipe instructions per step queueing behind the neighbours' DMMAs instead of one); d = hn - hc, which the
//     MMAs and the exit test need, is computed off that path;
//   * all of mu, the diagonal tile included, takes the block's eight rank-1 updates as DMMA.8x8x4 afterwards (full fp64
//     rate, operands straight from fragments): the next diagonal tile first, so its transposition to one-thread-per-column
//     overlaps the remaining MMAs;
//   * the convergence test is three fp64 operations per step whose sign bits are OR-ed: no dependence between steps.
// fp64 pipe budget per block and 32 columns: 56 DMMA x 16 cycles + ~70 DFMA-class x 2 cycles (DFMA/DMMA share one pipe at
// 64 FMA/clk/SM: scratch/mix_bench.cu). Arithmetic differences from the reference, all at rounding level: reciprocal and
// pre-multiplied V_rc/V_rr instead of a division per step; the candidate of coordinate r inside a block is accumulated
// as h_r - mu'_r/V_rr - sum_c (V_rc/V_rr) hn_c instead of through mu; `tmp != Hj(k)` becomes d = 0; the exit test
// 2|d|/(h_new+h_old+1e-16) > tol is evaluated as |d| - (tol/2)(h_new+h_old) > (tol/2)1e-16.
// Control flow per column is the reference's (stop when the max relative change <= rel_tol or at max_iter; finished
// columns are frozen while the rest of the tile keeps sweeping); sweep counts are summed into total_raw_iter.
#pragma once
#include <algorithm>
#include <type_traits>

#include "kernels.cuh"

namespace nnlm {
namespace scd_chain {

// NH = ceil(k / 4) half-blocks of 4 coordinates (the K extent of one DMMA); NB = ceil(NH / 2) blocks of 8. A trailing half-block
// that is all padding is skipped at compile time (k = 50: 13 half-blocks instead of 14).
// resident warps per CTA for 8- and 16-column tiles. Measured on config 2 (bench.py, iterations/s): 8 warps 435, 10 warps 443,
// 12 warps 473, 16 warps (128 registers, small spills) 414 — more warps per scheduler means more DMMAs in front of every
// chain instruction (profiles/r1_m_scd_stalls.md), fewer leave the pipe idle.
#ifndef NNLM_SCD_WARPS_NARROW
#define NNLM_SCD_WARPS_NARROW 12
#endif
template <int NH, int CT> struct Cfg { static constexpr int NB = (NH + 1) / 2, WARPS = (NB > 8 || CT == 4) ? 8 : NNLM_SCD_WARPS_NARROW; };

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double flip_sign(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }

__host__ __device__ constexpr int widx(int c, int r) { return c * (15 - c) / 2 + (r - c - 1); }   // dense index of the pair (c, r > c)
// offset of step c's multipliers inside a block's 32-entry record (7-c used entries, padded to an even count)
__host__ __device__ constexpr int woff(int c) { return c == 0 ? 0 : c == 1 ? 8 : c == 2 ? 14 : c == 3 ? 20 : c == 4 ? 24 : c == 5 ? 28 : 30; }

template <int NH, int CT>
__global__ void __launch_bounds__(32 * Cfg<NH, CT>::WARPS, 1)
k_scd_chain(double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ Qp, int splits,
            const uint8_t* __restrict__ mask, int k, int64_t ncol, double l1, unsigned max_iter, double rel_tol,
            unsigned long long* __restrict__ sweeps, unsigned int* __restrict__ next_group)
{
    constexpr int NB = Cfg<NH, CT>::NB, KB = 8 * NB, NC = 8 * CT, KS = KB + 4, WARPS = Cfg<NH, CT>::WARPS;
    extern __shared__ __align__(16) double sm[];
    double* gc = sm;                                   // [KB][KS]: gc[c*KS + r] = V[r, c] (symmetric; zero padded)
    double* rinv = gc + KB * KS;                       // [KB] 1 / V[c,c] (0 for padding)
    double* wl = rinv + KB;                            // [NB][32]: V[r,c] / V[r,r], r > c inside a diagonal tile, in step order
    double* wbase = wl + NB * 32 + (threadIdx.x >> 5) * (KB * NC + 16 * NC);
    double* hs = wbase;                                // [KB][NC] current h of the tile's columns
    double* dsm = hs + KB * NC;                        // [8][NC] d of the block just finished
    double* tsm = dsm + 8 * NC;                        // [8][NC] the next diagonal tile of mu, one thread per column

    for (int e = threadIdx.x; e < KB * KS; e += 32 * WARPS) {
        const int c = e / KS, r = e % KS;
        gc[e] = (r < k && c < k) ? G[r + k * c] : 0.0;
    }
    for (int c = threadIdx.x; c < KB; c += 32 * WARPS) rinv[c] = (c < k) ? 1.0 / G[c + k * c] : 0.0;
    for (int e = threadIdx.x; e < NB * 32; e += 32 * WARPS) wl[e] = 0.0;
    __syncthreads();
    for (int e = threadIdx.x; e < NB * 64; e += 32 * WARPS) {
        const int b = e >> 6, c = (e >> 3) & 7, r = e & 7;
        if (r > c) wl[b * 32 + woff(c) + (r - c - 1)] = rinv[8 * b + r] * gc[(8 * b + c) * KS + 8 * b + r];
    }
    __syncthreads();
    // From here on the MMA side works with V' = V minus the strictly lower triangle of every diagonal 8x8 block, i.e. it
    // maintains mu'_r = mu_r - sum_{c < r, same block} V_rc h_c. The candidates formed from it,
    // P_r = h_r - mu'_r / V_rr = h_r - mu_r / V_rr + sum_{c<r} (V_rc / V_rr) h_c(old), already contain the old h of the
    // earlier coordinates of the block, so a chain step folds in the NEW value with one FMA, P_r -= (V_rc / V_rr) hn_c,
    // and the difference d = hn - hc is off the dependent path (it is only published for the MMAs and the exit test).
    for (int e = threadIdx.x; e < NB * 64; e += 32 * WARPS) {
        const int b = e >> 6, c = (e >> 3) & 7, r = e & 7;
        if (r > c) gc[(8 * b + c) * KS + 8 * b + r] = 0.0;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3, warp = threadIdx.x >> 5;
    const int colx = lane < NC ? lane : NC - 1;        // the column this lane owns in the sequential part
    const int64_t ngroups = (ncol + NC - 1) / NC;
    const double tolh = 0.5 * rel_tol, c0 = tolh * TINY_NUM;
    unsigned long long my_sweeps = 0;
    // first round: groups dealt round-robin over the SMs, then over the warps (= schedulers) of an SM; later rounds on demand
    int64_t grp = (int64_t)blockIdx.x + (int64_t)gridDim.x * warp;

    while (grp < ngroups) {
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
                    if (r < k && c < cnt) {
#pragma unroll 1
                        for (int sp = 0; sp < splits; sp++) q += Qp[((int64_t)sp * ncol + col0 + c) * k + r];
                    }
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
#pragma unroll 1
                for (int r = 64 * w; r < k && r < 64 * w + 64; r++) mbits[w] |= (unsigned long long)(mc[r] != 0) << (r - 64 * w);
            }
            const int kw = k - 64 * w;                               // coordinates of this word that exist
            const unsigned long long kmask = kw >= 64 ? ~0ull : (kw <= 0 ? 0ull : ((1ull << kw) - 1ull));
            any_free = any_free || (mbits[w] & kmask) != kmask;
            mbits[w] &= kmask;                                       // padding coordinates have h = mu = 0: d = 0 by itself
        }
        bool cont = have && any_free;                                // fully masked column: src/update_with_missing.cpp:33-34
        __syncwarp();

        // ---- mu += V h : the block update with D := h (a rolled loop over the half-blocks: it runs once per group) ----
#pragma unroll 1
        for (int hb = 0; hb < NH; hb++) {
            double bf[CT];
#pragma unroll
            for (int ct = 0; ct < CT; ct++) bf[ct] = hs[(4 * hb + tig) * NC + 8 * ct + gid];
#pragma unroll
            for (int rt = 0; rt < NB; rt++) {
                const double a = gc[(4 * hb + tig) * KS + 8 * rt + gid];
#pragma unroll
                for (int ct = 0; ct < CT; ct++) dmma(mu[rt][ct][0], mu[rt][ct][1], a, bf[ct]);
            }
        }
        // the first diagonal tile scaled by 1 / V_rr (row gid of the fragment), one thread per column
        {
            const double ri = rinv[gid];
#pragma unroll
            for (int ct = 0; ct < CT; ct++)
                *reinterpret_cast<double2*>(tsm + gid * NC + 8 * ct + 2 * tig) = make_double2(mu[0][ct][0] * ri, mu[0][ct][1] * ri);
        }
        // multipliers of the first PRE steps of the next block, fetched while the MMAs of step (4) run (narrow tiles have the
        // registers for it): the chain then starts without waiting on shared-memory broadcasts
        constexpr int PRE = (NB > 8) ? 0 : (WARPS > 12 ? (CT == 1 ? 3 : 0) : (CT == 1 ? 7 : (CT == 2 ? 3 : 0)));
        double wn[PRE ? widx(PRE, PRE + 1) : 1];
#pragma unroll
        for (int c = 0; c < PRE; c++)
#pragma unroll
            for (int r = c + 1; r < 8; r++) wn[widx(c, r)] = wl[woff(c) + (r - c - 1)];

        // ---- sweeps ----
        // Software pipeline per block b:  (1) candidates of the block;  (2) the 8 sequential steps, interleaved with the
        // DEFERRED MMAs of the previous block (all row tiles except b itself: independent registers, they run in the shadow of
        // the dependent chain, and one warp alone keeps the fp64 pipe busy);  (3) publish d;  (4) block b's update of the NEXT
        // diagonal tile at once — the only part the next chain waits for — and its transposition.
        double bfp[2][CT];                                           // B fragments of the previous block's d
#pragma unroll
        for (int kh = 0; kh < 2; kh++)
#pragma unroll
            for (int ct = 0; ct < CT; ct++) bfp[kh][ct] = 0.0;
        unsigned t = 0;
        for (unsigned it = 0; it < max_iter; it++) {
            if (!__any_sync(0xffffffffu, cont)) break;
            unsigned long long fz[MW];                                // coordinates this sweep must leave alone
#pragma unroll
            for (int w = 0; w < MW; w++) fz[w] = cont ? mbits[w] : ~0ull;
            int flagbits = 0;
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const int pb = (b + NB - 1) % NB, nb = (b + 1) % NB;  // previous / next block (compile time after unrolling)
                // (1) unclamped candidates of the block's 8 coordinates
                __syncwarp();
                double P[8], h8[8], dd[8];
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    h8[r] = hs[(8 * b + r) * NC + colx];
                    P[r] = h8[r] - tsm[r * NC + colx];
                }
                // (2) the sequential steps + the deferred MMAs of block pb. The convergence test (3 fp64 operations per step)
                // is only evaluated while some running column of the tile has not exceeded the tolerance yet in this sweep.
                const double* wb = wl + b * 32;
                const bool need_flag = NB > 8 || __any_sync(0xffffffffu, cont && flagbits >= 0);
                auto steps = [&](auto with_flag) {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const int cc = 8 * b + c;
                        const double cand = P[c], hc = h8[c];
                        const bool live = !((fz[cc >> 6] >> (cc & 63)) & 1ull);
                        const double alt = live ? 0.0 : hc;                    // what h becomes when the candidate is not taken
                        const bool take = live && !(__double2hiint(cand) < 0); // dependent path: DFMA -> sign test -> select -> DFMA
                        const double hn = take ? cand : alt;
                        const double d = hn - hc;                              // off the dependent path
                        dd[c] = d;
                        h8[c] = hn;
#pragma unroll
                        for (int r = c + 1; r < 8; r++) P[r] = fma(-(c < PRE ? wn[widx(c, r)] : wb[woff(c) + (r - c - 1)]), hn, P[r]);
                        // 2|d| > tol (hn + hc + 1e-16)  <=>  (tol/2)(hn + hc) + (tol/2)1e-16 - |d| < 0 : collect the sign bits
                        if (decltype(with_flag)::value) flagbits |= __double2hiint(fma(hn + hc, tolh, c0 - fabs(d)));
                        if (NB > 1) {
                            // deferred (rt, kh) pairs are dealt round-robin to the steps: pair (rt, kh) goes to step
                            // (its ordinal among the pairs with rt != b) mod 8 — a pure function of the unrolled indices
#pragma unroll
                            for (int rt = 0; rt < NB; rt++)
#pragma unroll
                                for (int kh = 0; kh < 2; kh++) {
                                    const int ord = 2 * rt + kh - (rt > b ? 2 : 0);
                                    if (rt != b && 2 * pb + kh < NH && (ord & 7) == c) {
                                        const double a = gc[(8 * pb + 4 * kh + tig) * KS + 8 * rt + gid];
#pragma unroll
                                        for (int ct = 0; ct < CT; ct++) dmma(mu[rt][ct][0], mu[rt][ct][1], a, bfp[kh][ct]);
                                    }
                                }
                        }
                    }
                };
                if (NB > 8 || need_flag) steps(std::true_type{}); else steps(std::false_type{});   // (one copy keeps big ranks unrollable)
                // (3) publish d and the new h
                if (lane < NC) {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        dsm[c * NC + lane] = dd[c];
                        hs[(8 * b + c) * NC + lane] = h8[c];               // equals the old value whenever d = 0
                    }
                }
                __syncwarp();
                // (4) the next diagonal tile takes block b's update now; the other tiles during the next chain
#pragma unroll
                for (int kh = 0; kh < 2; kh++)
#pragma unroll
                    for (int ct = 0; ct < CT; ct++) bfp[kh][ct] = dsm[(4 * kh + tig) * NC + 8 * ct + gid];
#pragma unroll
                for (int kh = 0; kh < 2; kh++) {
                    if (2 * b + kh < NH) {
                        const double a = gc[(8 * b + 4 * kh + tig) * KS + 8 * nb + gid];
#pragma unroll
                        for (int ct = 0; ct < CT; ct++) dmma(mu[nb][ct][0], mu[nb][ct][1], a, bfp[kh][ct]);
                    }
                }
                {
                    const double ri = rinv[8 * nb + gid];
#pragma unroll
                    for (int ct = 0; ct < CT; ct++)
                        *reinterpret_cast<double2*>(tsm + gid * NC + 8 * ct + 2 * tig) = make_double2(mu[nb][ct][0] * ri, mu[nb][ct][1] * ri);
                }
#pragma unroll
                for (int c = 0; c < PRE; c++)
#pragma unroll
                    for (int r = c + 1; r < 8; r++) wn[widx(c, r)] = wl[nb * 32 + woff(c) + (r - c - 1)];
            }
            if (cont) t++;
            cont = cont && (flagbits < 0 || (0.0 > rel_tol));
        }
        if (have) my_sweeps += t;
        __syncwarp();
        for (int e = lane; e < total; e += 32) X[col0 * k + e] = hs[(e % k) * NC + e / k];
        __syncwarp();

        unsigned int g32 = 0;
        if (lane == 0) g32 = atomicAdd(next_group, 1u);
        grp = (int64_t)gridDim.x * WARPS + __shfl_sync(0xffffffffu, g32, 0);
    }
#pragma unroll
    for (int x = 16; x > 0; x >>= 1) my_sweeps += __shfl_xor_sync(0xffffffffu, my_sweeps, x);
    if (lane == 0 && my_sweeps) atomicAdd(sweeps, my_sweeps);
}

template <int NH, int CT>
void launch(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, double l1,
            unsigned max_iter, double rel_tol, unsigned long long* sweeps, unsigned int* counter, cudaStream_t st)
{
    constexpr int NB = Cfg<NH, CT>::NB, KB = 8 * NB, NC = 8 * CT, KS = KB + 4, WARPS = Cfg<NH, CT>::WARPS;
    const size_t smem = sizeof(double) * ((size_t)KB * KS + KB + NB * 32 + (size_t)WARPS * (KB * NC + 16 * NC));
    auto kern = k_scd_chain<NH, CT>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = ceil_div(ncol, NC);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(groups, 148));
    NNLM_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    kern<<<grid, 32 * WARPS, smem, st>>>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter);
    NNLM_LAUNCHED();
}

#define NNLM_SCDC_ARGS double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, \
    double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, unsigned int* counter, cudaStream_t st
#define NNLM_SCDC_PASS X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st

// explicit-instantiation entry points (several translation units keep the build parallel); nh = ceil(k / 4) <= 32
void launch_ct4_lo(int nh, NNLM_SCDC_ARGS);   // 32-column tiles, nh 1..8
void launch_ct4_hi(int nh, NNLM_SCDC_ARGS);   // 32-column tiles, nh 9..16
void launch_ct2_lo(int nh, NNLM_SCDC_ARGS);   // 16-column tiles
void launch_ct2_hi(int nh, NNLM_SCDC_ARGS);
void launch_ct1_lo(int nh, NNLM_SCDC_ARGS);   //  8-column tiles
void launch_ct1_hi(int nh, NNLM_SCDC_ARGS);
void launch_big_a(int nh, NNLM_SCDC_ARGS);    //  8-column tiles, 8 warps, nh 17..20 (k 65..80)
void launch_big_b(int nh, NNLM_SCDC_ARGS);    //  nh 21..24
void launch_big_c(int nh, NNLM_SCDC_ARGS);    //  nh 25..28
void launch_big_d(int nh, NNLM_SCDC_ARGS);    //  nh 29..32 (k <= 128)

}  // namespace scd_chain
}  // namespace nnlm
