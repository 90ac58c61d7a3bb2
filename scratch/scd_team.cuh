// scd_team.cuh — K3, third generation of the blocked SCD solver for the square loss: the work of one tile of columns is split
// over a TEAM of warps that sit on different schedulers — one chain warp and MM MMA warps (reference
// src/base_algorithms.cpp:3-37 preceded by mu = WtW*h - WtA (+beta2), src/update_with_missing.cpp:39-41).
//
// scd_chain.cuh runs the 8 dependent coordinate steps of a block and the block's DMMAs in ONE warp. Measured with one tile per
// SM (scratch/pair_bench.cu): ~920 cycles per block for an 8-column tile, ~1200 for 16 columns — the chain's DFMAs queue behind
// the warp's own 16-cycle DMMAs on the scheduler's single fp64 pipe, and nothing overlaps the chain's latency. That is the
// floor of every launch with fewer tiles than warp slots (the H-half, every shard of a multi-GPU run), and three such warps
// per scheduler keep the pipe only 60 % busy on the W-half. Here
//   * the CHAIN warp owns h and the candidates: per block it waits for the scaled diagonal tile of mu', runs the 8 steps
//     (DFMA -> sign test -> select -> DFMA; no DMMA and no exit-test arithmetic in that instruction stream: 44 fp64
//     instructions per block instead of 68 + 28 DMMAs) and publishes d;
//   * MMA warp j owns the row tiles rt = j (mod MM) of mu' as DMMA C fragments. The owner of the NEXT diagonal tile applies the
//     block's d to it first and publishes it scaled by 1/V_rr; all of them then give their other tiles the rank-8 update while
//     the chain warp is already running the next block.
// Two named barriers per team: `a` (everybody waits: h staged, d published, tile done) and `b` (MMA warps arrive, the chain
// warp waits: next diagonal tile published). Arithmetic is that of scd_chain.cuh operation by operation — every mu' entry
// receives the same DMMAs in the same order, a step computes the same expression — so results are bit-identical.
#pragma once
#include <algorithm>
#include <type_traits>

#include "scd_chain.cuh"

namespace nnlm {
namespace scd_team {

using scd_chain::dmma;
using scd_chain::widx;
using scd_chain::woff;

#ifdef NNLM_TEAM_PROF
__device__ long long g_team_prof[16];
#define TEAM_TICK(acc, t0) do { long long t1_ = clock64(); acc += t1_ - t0; t0 = t1_; } while (0)
#else
#define TEAM_TICK(acc, t0) do { } while (0)
#endif

// hn = (live && cand is not negative) ? cand : alt, with the sign test and `live` folded into ONE predicate
__device__ __forceinline__ double take_if_nonneg(double cand, double alt, int live)
{
    double hn;
    asm("{\n\t.reg .pred p, q;\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tsetp.ne.s32 q, %3, 0;\n\t"
        "setp.ge.and.s32 p, hi, 0, q;\n\tselp.f64 %0, %1, %2, p;\n\t}"
        : "=d"(hn) : "d"(cand), "d"(alt), "r"(live));
    return hn;
}

__device__ __forceinline__ void team_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// non-blocking arrival; `dep` is a value whose producing loads must have returned before the others may overwrite their source
__device__ __forceinline__ void team_arrive(int id, int nthreads, double dep)
{
    asm volatile("bar.arrive %0, %1; // %2" ::"r"(id), "r"(nthreads), "d"(dep) : "memory");
}

template <int NH, int CT, int MM> struct Cfg {
    static constexpr int NB = (NH + 1) / 2;
    static constexpr int GW = 1 + MM;                                        // warps per team
    static constexpr int TEAMS = MM == 1 ? 7 : (MM == 2 ? 5 : 3);            // two barrier ids each, ids 1..15; <= 16 warps
    static constexpr int THREADS = 32 * GW * TEAMS;
    static constexpr int NTL = (NB + MM - 1) / MM;                           // row tiles per MMA warp (upper bound)
};

template <int NH, int CT, int MM>
struct Smem {
    static constexpr int NB = Cfg<NH, CT, MM>::NB, KB = 8 * NB, NC = 8 * CT, KS = KB + 4;
    static constexpr int PER_TEAM = KB * NC + 16 * NC + 2;
    static constexpr size_t doubles = (size_t)KB * KS + KB + NB * 32 + (size_t)Cfg<NH, CT, MM>::TEAMS * PER_TEAM;
};

struct Args {
    double* X; const double* Qp; int splits; const uint8_t* mask; int k; int64_t ncol; double l1; unsigned max_iter; double rel_tol;
    unsigned long long* sweeps; unsigned int* next_group;
};

// ------------------------------------------------ MMA warp J of a team ------------------------------------------------
template <int NH, int CT, int MM, int J>
__device__ __forceinline__ void mma_role(const Args& p, const double* gc, const double* rinv, double* hs, const double* dsm,
                                         double* tsm, volatile int* ctl, int ida, int idb, int64_t grp, int lane)
{
    constexpr int NB = Cfg<NH, CT, MM>::NB, NC = 8 * CT, KS = 8 * NB + 4, GW = Cfg<NH, CT, MM>::GW, TEAMS = Cfg<NH, CT, MM>::TEAMS;
    constexpr int NTL = (NB - J + MM - 1) / MM;                     // tiles J, J + MM, ... of this warp
    const int gid = lane >> 2, tig = lane & 3, k = p.k;
    const int64_t ngroups = (p.ncol + NC - 1) / NC;
    while (grp < ngroups) {
        const int64_t col0 = grp * NC;
        const int cnt = (int)min((int64_t)NC, p.ncol - col0);
        double mu[NTL > 0 ? NTL : 1][CT][2];
#pragma unroll
        for (int li = 0; li < NTL; li++)
#pragma unroll
            for (int ct = 0; ct < CT; ct++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int r = 8 * (li * MM + J) + gid, c = 8 * ct + 2 * tig + e;
                    double q = 0.0;
                    if (r < k && c < cnt) {
#pragma unroll 1
                        for (int sp = 0; sp < p.splits; sp++) q += p.Qp[((int64_t)sp * p.ncol + col0 + c) * k + r];
                    }
                    mu[li][ct][e] = (r < k && c < cnt) ? p.l1 - q : 0.0;
                }
        team_sync(ida, 32 * GW);                                    // S0: the chain warp has staged h and its verdict
        bool go = ctl[0] != 0;
        if (go) {
#pragma unroll 1
            for (int hb = 0; hb < NH; hb++) {
                double bf[CT];
#pragma unroll
                for (int ct = 0; ct < CT; ct++) bf[ct] = hs[(4 * hb + tig) * NC + 8 * ct + gid];
#pragma unroll
                for (int li = 0; li < NTL; li++) {
                    const double a = gc[(4 * hb + tig) * KS + 8 * (li * MM + J) + gid];
#pragma unroll
                    for (int ct = 0; ct < CT; ct++) dmma(mu[li][ct][0], mu[li][ct][1], a, bf[ct]);
                }
            }
            double dep = 0.0;
            if (J == 0) {                                           // tile 0 is this warp's: the first diagonal tile, scaled
                const double ri = rinv[gid];
#pragma unroll
                for (int ct = 0; ct < CT; ct++)
                    *reinterpret_cast<double2*>(tsm + gid * NC + 8 * ct + 2 * tig) = make_double2(mu[0][ct][0] * ri, mu[0][ct][1] * ri);
            }
            team_arrive(idb, 32 * GW, dep);                         // S1
#ifdef NNLM_TEAM_PROF
            long long pt = clock64(), pm_def = 0, pm_crit = 0;
#endif
            while (go) {
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    const int nb = (b + 1) % NB;
                    const bool owner = (nb % MM) == J;              // compile time after unrolling
                    double an[2] = {0.0, 0.0}, ri = 0.0;
                    if (owner) {                                    // A fragments of the next diagonal tile: fetched before the wait
#pragma unroll
                        for (int kh = 0; kh < 2; kh++)
                            if (2 * b + kh < NH) an[kh] = gc[(8 * b + 4 * kh + tig) * KS + 8 * nb + gid];
                        ri = rinv[8 * nb + gid];
                    }
                    TEAM_TICK(pm_def, pt);
                    team_sync(ida, 32 * GW);                        // S2: d of block b is published
                    double bfp[2][CT];
#pragma unroll
                    for (int kh = 0; kh < 2; kh++)
#pragma unroll
                        for (int ct = 0; ct < CT; ct++) bfp[kh][ct] = dsm[(4 * kh + tig) * NC + 8 * ct + gid];
                    go = ctl[0] != 0;
                    if (!go) break;
                    if (owner) {
#pragma unroll
                        for (int kh = 0; kh < 2; kh++)
                            if (2 * b + kh < NH) {
#pragma unroll
                                for (int ct = 0; ct < CT; ct++) dmma(mu[nb / MM][ct][0], mu[nb / MM][ct][1], an[kh], bfp[kh][ct]);
                            }
#pragma unroll
                        for (int ct = 0; ct < CT; ct++)
                            *reinterpret_cast<double2*>(tsm + gid * NC + 8 * ct + 2 * tig) =
                                make_double2(mu[nb / MM][ct][0] * ri, mu[nb / MM][ct][1] * ri);
                    }
                    double depv = bfp[0][0];
#pragma unroll
                    for (int ct = 0; ct < CT; ct++) depv += bfp[1][ct] + bfp[0][ct];   // (all of d has arrived in registers)
                    team_arrive(idb, 32 * GW, depv);                // S1: the chain warp may start block nb
                    TEAM_TICK(pm_crit, pt);
#pragma unroll
                    for (int li = 0; li < NTL; li++) {
                        const int rt = li * MM + J;
                        if (rt != nb) {
#pragma unroll
                            for (int kh = 0; kh < 2; kh++)
                                if (2 * b + kh < NH) {
                                    const double a = gc[(8 * b + 4 * kh + tig) * KS + 8 * rt + gid];
#pragma unroll
                                    for (int ct = 0; ct < CT; ct++) dmma(mu[li][ct][0], mu[li][ct][1], a, bfp[kh][ct]);
                                }
                        }
                    }
                }
            }
#ifdef NNLM_TEAM_PROF
            if (lane == 0 && blockIdx.x == 0 && threadIdx.x < 32 * GW) { g_team_prof[2 * J] = pm_def; g_team_prof[2 * J + 1] = pm_crit; }
#endif
        }
        team_sync(ida, 32 * GW);                                    // E: the tile is written back, the next group is published
        grp = (int64_t)gridDim.x * TEAMS + (unsigned)ctl[1];
    }
}

// --------------------------------------------------- the chain warp ---------------------------------------------------
template <int NH, int CT, int MM>
__device__ __forceinline__ void chain_role(const Args& p, const double* wl, double* hs, double* dsm, const double* tsm,
                                           volatile int* ctl, int ida, int idb, int64_t grp, int lane)
{
    constexpr int NB = Cfg<NH, CT, MM>::NB, KB = 8 * NB, NC = 8 * CT, GW = Cfg<NH, CT, MM>::GW, TEAMS = Cfg<NH, CT, MM>::TEAMS;
    const int k = p.k;
    const int64_t ngroups = (p.ncol + NC - 1) / NC;
    const int colx = lane < NC ? lane : NC - 1;
    const double tolh = 0.5 * p.rel_tol, c0 = tolh * TINY_NUM;
    unsigned long long my_sweeps = 0;
    while (grp < ngroups) {
        const int64_t col0 = grp * NC;
        const int cnt = (int)min((int64_t)NC, p.ncol - col0);
        const int total = cnt * k;
        for (int e = lane; e < KB * NC; e += 32) hs[e] = 0.0;
        __syncwarp();
        for (int e = lane; e < total; e += 32) hs[(e % k) * NC + e / k] = p.X[col0 * k + e];
        constexpr int MW = (KB + 63) / 64;
        unsigned long long mbits[MW];
        const bool have = lane < cnt;
        bool any_free = false;
#pragma unroll
        for (int w = 0; w < MW; w++) {
            mbits[w] = 0;
            if (p.mask != nullptr && have) {
                const uint8_t* mc = p.mask + (col0 + lane) * k;
#pragma unroll 1
                for (int r = 64 * w; r < k && r < 64 * w + 64; r++) mbits[w] |= (unsigned long long)(mc[r] != 0) << (r - 64 * w);
            }
            const int kw = k - 64 * w;
            const unsigned long long kmask = kw >= 64 ? ~0ull : (kw <= 0 ? 0ull : ((1ull << kw) - 1ull));
            any_free = any_free || (mbits[w] & kmask) != kmask;
            mbits[w] &= kmask;                                       // padding coordinates have h = mu = 0: d = 0 by itself
        }
        bool cont = have && any_free;                                // fully masked column: src/update_with_missing.cpp:33-34
        bool go = p.max_iter > 0 && __any_sync(0xffffffffu, cont);
        if (lane == 0) ctl[0] = go ? 1 : 0;
        // the 28 multipliers of a block live in registers for the whole block (the chain must not wait on shared memory)
        double2 wn[16];
#pragma unroll
        for (int i = 0; i < 16; i++) wn[i] = reinterpret_cast<const double2*>(wl)[i];
        team_sync(ida, 32 * GW);                                     // S0
        unsigned t = 0;
        int flagbits = 0;
        bool need_flag = true;
        double hold[8];
#pragma unroll
        for (int r = 0; r < 8; r++) hold[r] = hs[r * NC + colx];
#ifdef NNLM_TEAM_PROF
        long long pt = clock64(), pc_init = 0, pc_steps = 0, pc_pub = 0, pc_pre = 0;
#endif
        if (go) team_sync(idb, 32 * GW);                             // S1: the first diagonal tile is published
        for (unsigned it = 0; go; it++) {
            unsigned long long fz[MW];
#pragma unroll
            for (int w = 0; w < MW; w++) fz[w] = cont ? mbits[w] : ~0ull;
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const int nb = (b + 1) % NB;
                double P[8];
#pragma unroll
                for (int r = 0; r < 8; r++) P[r] = hold[r] - tsm[r * NC + colx];
                TEAM_TICK(pc_init, pt);
                // the sequential steps. d and the new h go to shared memory as they appear; the exit test (three fp64
                // operations per coordinate) is evaluated AFTER the hand-over, from those stores, except in a sweep's last block
                const bool inline_flag = (b == NB - 1) && need_flag;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int cc = 8 * b + c;
                    const double cand = P[c], hc = hold[c];
                    const int live = (int)(((fz[cc >> 6] >> (cc & 63)) & 1ull) ^ 1ull);
                    const double alt = live ? 0.0 : hc;
                    const double hn = take_if_nonneg(cand, alt, live);      // dependent path: DFMA -> ISETP -> SEL -> DFMA
#pragma unroll
                    for (int r = c + 1; r < 8; r++) {
                        const int wi = woff(c) + (r - c - 1);
                        P[r] = fma(-((wi & 1) ? wn[wi >> 1].y : wn[wi >> 1].x), hn, P[r]);
                    }
                    const double d = hn - hc;
                    if (lane < NC) {
                        dsm[c * NC + lane] = d;
                        hs[cc * NC + lane] = hn;                            // equals the old value whenever d = 0
                    }
                    if (b == NB - 1) {
                        if (inline_flag) flagbits |= __double2hiint(fma(hn + hc, tolh, c0 - fabs(d)));
                    }
                }
                TEAM_TICK(pc_steps, pt);
                if (b == NB - 1) {                                          // end of the sweep: does anybody go on?
                    if (cont) t++;
                    cont = cont && (flagbits < 0 || (0.0 > p.rel_tol));
                    go = (it + 1 < p.max_iter) && __any_sync(0xffffffffu, cont);
                    if (lane == 0) ctl[0] = go ? 1 : 0;
                    flagbits = 0;
                }
                team_sync(ida, 32 * GW);                                    // S2: d (and the verdict) are published
                TEAM_TICK(pc_pub, pt);
                if (b == NB - 1 && !go) break;
                // off the critical path (the MMA warps are updating the next diagonal tile): the exit test of this block, the
                // next block's h and multipliers
                if (b != NB - 1 && need_flag) {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const double d = dsm[c * NC + colx], hn = hs[(8 * b + c) * NC + colx];
                        flagbits |= __double2hiint(fma(hn + hold[c], tolh, c0 - fabs(d)));
                    }
                }
                need_flag = NB > 8 || __any_sync(0xffffffffu, cont && flagbits >= 0);
#pragma unroll
                for (int r = 0; r < 8; r++) hold[r] = hs[(8 * nb + r) * NC + colx];
#pragma unroll
                for (int i = 0; i < 16; i++) wn[i] = reinterpret_cast<const double2*>(wl + nb * 32)[i];
                TEAM_TICK(pc_pre, pt);
                team_sync(idb, 32 * GW);                                    // S1: the next diagonal tile is published
            }
        }
        if (have) my_sweeps += t;
#ifdef NNLM_TEAM_PROF
        if (lane == 0 && blockIdx.x == 0 && threadIdx.x < 32 * GW) {
            g_team_prof[8] = t; g_team_prof[9] = pc_init; g_team_prof[10] = pc_steps; g_team_prof[11] = pc_pub; g_team_prof[12] = pc_pre;
        }
#endif
        for (int e = lane; e < total; e += 32) p.X[col0 * k + e] = hs[(e % k) * NC + e / k];
        if (lane == 0) ctl[1] = (int)atomicAdd(p.next_group, 1u);
        team_sync(ida, 32 * GW);                                            // E
        grp = (int64_t)gridDim.x * TEAMS + (unsigned)ctl[1];
    }
#pragma unroll
    for (int x = 16; x > 0; x >>= 1) my_sweeps += __shfl_xor_sync(0xffffffffu, my_sweeps, x);
    if (lane == 0 && my_sweeps) atomicAdd(p.sweeps, my_sweeps);
}

template <int NH, int CT, int MM>
__global__ void __maxnreg__(128)
k_scd_team(const double* __restrict__ G, Args p)
{
    constexpr int NB = Cfg<NH, CT, MM>::NB, KB = 8 * NB, NC = 8 * CT, KS = KB + 4, GW = Cfg<NH, CT, MM>::GW;
    constexpr int THREADS = Cfg<NH, CT, MM>::THREADS, PER_TEAM = Smem<NH, CT, MM>::PER_TEAM;
    extern __shared__ __align__(16) double sm[];
    double* gc = sm;                                   // [KB][KS]: gc[c*KS + r] = V[r, c] (symmetric; zero padded)
    double* rinv = gc + KB * KS;                       // [KB] 1 / V[c,c] (0 for padding)
    double* wl = rinv + KB;                            // [NB][32]: V[r,c] / V[r,r], r > c inside a diagonal tile, in step order
    const int k = p.k;
    for (int e = threadIdx.x; e < KB * KS; e += THREADS) {
        const int c = e / KS, r = e % KS;
        gc[e] = (r < k && c < k) ? G[r + k * c] : 0.0;
    }
    for (int c = threadIdx.x; c < KB; c += THREADS) rinv[c] = (c < k) ? 1.0 / G[c + k * c] : 0.0;
    for (int e = threadIdx.x; e < NB * 32; e += THREADS) wl[e] = 0.0;
    __syncthreads();
    for (int e = threadIdx.x; e < NB * 64; e += THREADS) {
        const int b = e >> 6, c = (e >> 3) & 7, r = e & 7;
        if (r > c) wl[b * 32 + woff(c) + (r - c - 1)] = rinv[8 * b + r] * gc[(8 * b + c) * KS + 8 * b + r];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NB * 64; e += THREADS) {         // V' = V minus the strict lower triangle of the diagonal tiles
        const int b = e >> 6, c = (e >> 3) & 7, r = e & 7;
        if (r > c) gc[(8 * b + c) * KS + 8 * b + r] = 0.0;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, team = warp / GW;
    int pos = warp % GW;
    if (GW == 2) pos ^= (team >> 1) & 1;                            // two-warp teams: alternate so every scheduler hosts both roles
    double* hs = wl + NB * 32 + team * PER_TEAM;                    // [KB][NC] current h of the tile's columns
    double* dsm = hs + KB * NC;                                     // [8][NC] d of the block just finished (chain -> MMA)
    double* tsm = dsm + 8 * NC;                                     // [8][NC] the next diagonal tile of mu' / V_rr (MMA -> chain)
    volatile int* ctl = reinterpret_cast<volatile int*>(tsm + 8 * NC);   // [0] keep sweeping, [1] next group
    const int ida = 1 + 2 * team, idb = 2 + 2 * team;
    const int64_t grp = (int64_t)blockIdx.x + (int64_t)gridDim.x * team;
    if (pos == 0) chain_role<NH, CT, MM>(p, wl, hs, dsm, tsm, ctl, ida, idb, grp, lane);
    else if (pos == 1) mma_role<NH, CT, MM, 0>(p, gc, rinv, hs, dsm, tsm, ctl, ida, idb, grp, lane);
    else if (MM > 1 && pos == 2) mma_role<NH, CT, MM, (MM > 1 ? 1 : 0)>(p, gc, rinv, hs, dsm, tsm, ctl, ida, idb, grp, lane);
    else if (MM > 2 && pos == 3) mma_role<NH, CT, MM, (MM > 2 ? 2 : 0)>(p, gc, rinv, hs, dsm, tsm, ctl, ida, idb, grp, lane);
    else if (MM > 3 && pos == 4) mma_role<NH, CT, MM, (MM > 3 ? 3 : 0)>(p, gc, rinv, hs, dsm, tsm, ctl, ida, idb, grp, lane);
}

template <int NH, int CT, int MM>
void launch(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, double l1,
            unsigned max_iter, double rel_tol, unsigned long long* sweeps, unsigned int* counter, cudaStream_t st)
{
    constexpr int NC = 8 * CT, THREADS = Cfg<NH, CT, MM>::THREADS;
    const size_t smem = sizeof(double) * Smem<NH, CT, MM>::doubles;
    auto kern = k_scd_team<NH, CT, MM>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = ceil_div(ncol, NC);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(groups, 148));
    NNLM_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    Args a{X, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter};
    kern<<<grid, THREADS, smem, st>>>(G, a);
    NNLM_LAUNCHED();
}

// explicit-instantiation entry points; nh = ceil(k / 4)
void launch_t11_a(int nh, NNLM_SCDC_ARGS);    // 8-column tiles, one MMA warp, nh 1..8
void launch_t11_b(int nh, NNLM_SCDC_ARGS);    // nh 9..16
void launch_t11_c(int nh, NNLM_SCDC_ARGS);    // nh 17..24
void launch_t11_d(int nh, NNLM_SCDC_ARGS);    // nh 25..32
void launch_t22_a(int nh, NNLM_SCDC_ARGS);    // 16-column tiles, two MMA warps, nh 1..8
void launch_t22_b(int nh, NNLM_SCDC_ARGS);    // nh 9..16

}  // namespace scd_team
}  // namespace nnlm
