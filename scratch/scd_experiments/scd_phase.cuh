// scd_phase.cuh — K3/K4, the blocked sequential-coordinate-descent solver for the square loss, rank k <= 64, with the
// chain and MMA work of a scheduler separated in time (reference src/base_algorithms.cpp:3-37 preceded by
// mu = WtW*h - WtA (+beta2), src/update_with_missing.cpp:39-41).
//
// Blocking as in scd_dmma.cuh / scd_chain.cuh: coordinates in blocks of 8; mu = V h - q of a group of 8*CT columns lives in
// DMMA C fragments of one warp; per block the warp (1) runs the 8 sequential steps with one thread per column and then
// (2) applies the block's eight rank-1 updates to all of mu as DMMA.8x8x4. What those kernels got wrong is WHEN: DFMA and
// DMMA share the fp64 pipe of a scheduler, a DMMA holds it for 16 cycles, and a dependent DFMA chain that takes 8 cycles
// per step alone takes 56 when three other warps of the scheduler are issuing DMMAs (scratch/smsp_map.cu; warp w runs on
// scheduler w % 4). With free-running warps every chain instruction queued behind the neighbours' MMAs: pipe 55-60 % busy.
// A first fix that pinned all chains to scheduler 0 (scd_ws.cuh) ran out of ISSUE slots there (12 groups x ~500
// instructions per block). Here the three warps of a scheduler alternate together, with two named barriers per block:
//   chain phase: all three run their 8 steps (no MMA in flight on this scheduler: 8-cycle DFMAs, ~26 cycles per step);
//   MMA phase:   all three issue their MMAs back to back (the pipe is the bound and stays full).
// Further changes against scd_chain.cuh:
//   * rows beyond the last full tile (k mod 8 in 1..3, e.g. k = 50) are kept one-thread-per-column next to the chain
//     (3 extra FMAs per step) and their coordinates form a short last block: at k = 50 a warp holds 6 row tiles instead
//     of 7 — 32-column groups fit 12 warps per SM at 168 registers — and a sweep costs 6 x 13 MMAs per 8 columns, the
//     unpadded count;
//   * h stays in X (L2-resident): the chain reads/writes 64 B per thread and block; the loads for the next block are
//     issued at the start of the MMA phase.
// Arithmetic differences from the reference, all at rounding level: reciprocal and pre-multiplied V_rc/V_rr instead of a
// division per step; inside a block the candidate of coordinate r is h_r - mu_r/V_rr - sum_c (V_rc/V_rr) d_c instead of
// going through mu; `tmp != Hj(k)` becomes d = 0; the exit test 2|d|/(h_new+h_old+1e-16) > tol is evaluated as
// (tol/2)(h_new+h_old) + (tol/2)1e-16 - |d| < 0; rows outside the block see its 8 updates summed 4 at a time inside the MMA.
// Control flow per column is the reference's (a column stops when its max relative change <= rel_tol or at max_iter;
// finished columns are frozen while the rest of the group keeps sweeping); sweep counts are summed into total_raw_iter.
#pragma once
#include <algorithm>

#include "kernels.cuh"

namespace nnlm {
namespace scd_phase {

constexpr int WARPS = 12;                 // 3 per scheduler

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip_sign(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }
// named barriers over the 3 warps (96 threads) of one scheduler
__device__ __forceinline__ void sched_sync(int id) { asm volatile("barrier.sync %0, 96;" ::"r"(id) : "memory"); }
__device__ __forceinline__ bool sched_or(int id, bool v)
{
    int r;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.s32 p, %2, 0;\n\t"
        "barrier.red.or.pred q, %1, 96, p;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t"
        "}" : "=r"(r) : "r"(id), "r"((int)v) : "memory");
    return r != 0;
}

__host__ __device__ constexpr int widx(int c, int r) { return c * (15 - c) / 2 + (r - c - 1); }   // dense index of pair (c, r > c)
// offset of step c's multipliers inside a block's 32-entry record (7-c used entries, padded to an even count)
__host__ __device__ constexpr int woff(int c) { return c == 0 ? 0 : c == 1 ? 8 : c == 2 ? 14 : c == 3 ? 20 : c == 4 ? 24 : c == 5 ? 28 : 30; }

// shared-memory layout in doubles
template <int NT, bool REM, int CT> struct Lay {
    static constexpr int KT = 8 * NT;                  // coordinates backed by row tiles
    static constexpr int KB = KT + (REM ? 4 : 0);      // padded coordinate count (the remainder is one half-block)
    static constexpr int KS = KT + 4;                  // pitch of gc: gc[c*KS + r] = V[r, c], r < KT, c < KB
    static constexpr int NC = 8 * CT;
    static constexpr int DP = NC + 4, TP = NC + 8;     // bank-conflict-free pitches of the d and tile buffers
    static constexpr int NBLK = NT + (REM ? 1 : 0);
    static constexpr int GC = 0;
    static constexpr int RINV = GC + KB * KS;
    static constexpr int WL = RINV + KB;               // [NBLK][32]
    static constexpr int VREM = WL + NBLK * 32;        // [KB][4]: V[KT + r, c] (REM only)
    static constexpr int WARP0 = VREM + (REM ? KB * 4 : 0);
    static constexpr int DSM = 0, TSM = 8 * DP, PERWARP = TSM + 8 * TP;
    static constexpr int TOTAL = WARP0 + WARPS * PERWARP;
};

template <int NT, bool REM, int CT>
__global__ void __launch_bounds__(32 * WARPS, 1)
k_scd_phase(double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ Qp, int splits,
            const uint8_t* __restrict__ mask, int k, int64_t ncol, double l1, unsigned max_iter, double rel_tol,
            unsigned long long* __restrict__ sweeps, unsigned int* __restrict__ next_group)
{
    using L = Lay<NT, REM, CT>;
    constexpr int KT = L::KT, KB = L::KB, KS = L::KS, NC = L::NC, NBLK = L::NBLK, DP = L::DP, TP = L::TP;
    constexpr bool PRELOAD = CT <= 2;
    extern __shared__ __align__(16) double sm[];
    double* gc = sm + L::GC;
    double* rinv = sm + L::RINV;
    double* wl = sm + L::WL;
    double* vrem = sm + L::VREM;

    for (int e = threadIdx.x; e < KB * KS; e += 32 * WARPS) {
        const int c = e / KS, r = e % KS;
        gc[e] = (r < k && r < KT && c < k) ? G[r + k * c] : 0.0;
    }
    for (int c = threadIdx.x; c < KB; c += 32 * WARPS) rinv[c] = (c < k) ? 1.0 / G[c + k * c] : 0.0;
    for (int e = threadIdx.x; e < NBLK * 32; e += 32 * WARPS) wl[e] = 0.0;
    if (REM)
        for (int e = threadIdx.x; e < KB * 4; e += 32 * WARPS) {
            const int c = e >> 2, r = KT + (e & 3);
            vrem[e] = (r < k && c < k) ? G[r + k * c] : 0.0;
        }
    __syncthreads();
    for (int e = threadIdx.x; e < NBLK * 64; e += 32 * WARPS) {
        const int b = e >> 6, c = (e >> 3) & 7, r = e & 7;
        const int rr = 8 * b + r, cc = 8 * b + c;
        if (r > c && rr < k && cc < k) wl[b * 32 + woff(c) + (r - c - 1)] = rinv[rr] * G[rr + k * cc];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const int bid = 1 + (warp & 3);                   // barrier of this warp's scheduler (warps w, w+4, w+8)
    const int colx = lane < NC ? lane : NC - 1;
    double* dsm = sm + L::WARP0 + warp * L::PERWARP + L::DSM;     // [8][DP] d of the block just finished
    double* tsm = sm + L::WARP0 + warp * L::PERWARP + L::TSM;     // [8][TP] the next diagonal tile of mu, one thread per column
    const int64_t ngroups = (ncol + NC - 1) / NC;
    const double tolh = 0.5 * rel_tol, c0 = tolh * TINY_NUM;
    unsigned long long my_sweeps = 0;
    // first round: groups dealt round-robin over the SMs, then over the warps of an SM; later rounds on demand
    int64_t grp = (int64_t)blockIdx.x + (int64_t)gridDim.x * warp;
    // schedulers 2 and 3 start half a block period late: their chain phases (shared-memory bound) then fall into the MMA
    // phases (pipe bound) of schedulers 0 and 1 instead of all twelve warps broadcasting multipliers at the same time
    if (CT == 4 && (warp & 2)) {
        const long long t0 = clock64();
        while (clock64() - t0 < 1800) { }
    }

    while (true) {
        const bool has = grp < ngroups;
        if (!sched_or(bid, has)) break;
        const int64_t col0 = has ? grp * NC : 0;
        const int cnt = has ? (int)min((int64_t)NC, ncol - col0) : 0;
        const bool have = lane < cnt;
        double* xcol = X + (col0 + (have ? lane : 0)) * k;          // this lane's column of h
        const bool xvec = (reinterpret_cast<uintptr_t>(xcol) & 15) == 0 && (k & 1) == 0;

        // ---- mu = l1 - q in C-fragment layout ----
        double mu[NT][CT][2];
#pragma unroll
        for (int rt = 0; rt < NT; rt++)
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
        // ---- mu += V h : the block update with D := h, read straight from X ----
        if (has) {
#pragma unroll 1
            for (int hb = 0; hb < KB / 4; hb++) {
                double bf[CT];
#pragma unroll
                for (int ct = 0; ct < CT; ct++) {
                    const int c = 8 * ct + gid, r = 4 * hb + tig;
                    bf[ct] = (c < cnt && r < k) ? X[(col0 + c) * k + r] : 0.0;
                }
#pragma unroll
                for (int rt = 0; rt < NT; rt++) {
                    const double a = gc[(4 * hb + tig) * KS + 8 * rt + gid];
#pragma unroll
                    for (int ct = 0; ct < CT; ct++) dmma(mu[rt][ct][0], mu[rt][ct][1], a, bf[ct]);
                }
            }
        }
        // ---- per-column state of the sequential part (lane = column) ----
        constexpr int MW = (KB + 63) / 64;                           // 64-bit words of the per-column coordinate mask
        unsigned long long mbits[MW];
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
        double mrem[3] = {0.0, 0.0, 0.0};                            // mu of the remainder rows: l1 - q + V[rem, :] h
        if (REM) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
                double q = 0.0;
                if (have && KT + r < k) {
#pragma unroll 1
                    for (int sp = 0; sp < splits; sp++) q += Qp[((int64_t)sp * ncol + col0 + lane) * k + KT + r];
                }
                mrem[r] = (have && KT + r < k) ? l1 - q : 0.0;
            }
            if (have) {
#pragma unroll 1
                for (int c = 0; c < k; c++) {
                    const double hc = xcol[c];
#pragma unroll
                    for (int r = 0; r < 3; r++) mrem[r] = fma(vrem[c * 4 + r], hc, mrem[r]);
                }
            }
        }
        // the first diagonal tile scaled by 1 / V_rr (row gid of the fragment), one thread per column; h of the first block
        {
            const double ri = rinv[gid];
#pragma unroll
            for (int ct = 0; ct < CT; ct++)
                *reinterpret_cast<double2*>(tsm + gid * TP + 8 * ct + 2 * tig) = make_double2(mu[0][ct][0] * ri, mu[0][ct][1] * ri);
        }
        double pm[3] = {0.0, 0.0, 0.0};                              // mrem / V_rr at the time the remainder block starts
        double wn[PRELOAD ? 18 : 1];                                 // multipliers of the next block, fetched ahead (narrow groups)
        if (PRELOAD) {
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int r = c + 1; r < 8; r++) wn[widx(c, r)] = wl[woff(c) + (r - c - 1)];
        }
        double h8[8];
#pragma unroll
        for (int r = 0; r < 8; r++) h8[r] = (have && r < k) ? xcol[r] : 0.0;

        // ---- sweeps ----
        unsigned t = 0;
        for (unsigned it = 0; it < max_iter; it++) {
            const bool alive = has && __any_sync(0xffffffffu, cont);
            if (!sched_or(bid, alive)) break;
            int flagbits = 0;
#pragma unroll
            for (int b = 0; b < NBLK; b++) {
                const bool tile = !REM || b < NT;                    // tile-backed block (8 steps) or the remainder block (3 steps)
                const int nb = (b + 1) % NBLK;                       // next block; nbt = its row tile if it has one
                const int nbt = (!REM || nb < NT) ? nb : -1;
                const int cb = 8 * b;
                // ======== chain phase ========
                if (alive) {
                    __syncwarp();
                    const unsigned fz8 = cont ? (unsigned)(mbits[cb >> 6] >> (cb & 63)) & 0xffu : 0xffu;
                    double P[8];
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        if (tile) P[r] = h8[r] - tsm[r * TP + colx];
                        else P[r] = (r < 3) ? h8[r] - pm[r] : 0.0;
                    }
                    const double* wb = wl + b * 32;
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        if (!tile && c >= 3) { if (lane < NC) dsm[c * DP + lane] = 0.0; continue; }
                        const double cand = P[c], hc = h8[c];
                        const bool neg = __double2hiint(cand) < 0;
                        const bool live = !((fz8 >> c) & 1u);
                        const double dpos = cand - hc;
                        double d = neg ? flip_sign(hc) : dpos;             // integer sign flip: keeps the fp64 pipe for the FMAs
                        double hn = neg ? 0.0 : cand;
                        d = live ? d : 0.0;
                        hn = live ? hn : hc;
                        if (lane < NC) dsm[c * DP + lane] = d;              // publish d
                        h8[c] = hn;
#pragma unroll
                        for (int r = c + 1; r < (tile ? 8 : 3); r++)
                            P[r] = fma(-((PRELOAD && c < 3) ? wn[widx(c, r)] : wb[woff(c) + (r - c - 1)]), d, P[r]);
                        // 2|d| > tol (hn + hc + 1e-16)  <=>  (tol/2)(hn + hc) + (tol/2)1e-16 - |d| < 0 : collect the sign bits
                        flagbits |= __double2hiint(fma(hn + hc, tolh, c0 - fabs(d)));
                    }
                    // the new h
                    if (cont) {
                        if (xvec && cb + 8 <= k) {
                            double2* hp = reinterpret_cast<double2*>(xcol + cb);
                            hp[0] = make_double2(h8[0], h8[1]); hp[1] = make_double2(h8[2], h8[3]);
                            hp[2] = make_double2(h8[4], h8[5]); hp[3] = make_double2(h8[6], h8[7]);
                        } else {
#pragma unroll
                            for (int r = 0; r < 8; r++) if (cb + r < k) xcol[cb + r] = h8[r];
                        }
                    }
                }
                sched_sync(bid);
                // ======== MMA phase ========
                if (alive) {
                    __syncwarp();
                    // h of the next block: in flight during the MMAs
                    const int cn = 8 * nb;
                    if (have) {
                        if (xvec && cn + 8 <= k) {
                            const double2* hp = reinterpret_cast<const double2*>(xcol + cn);
                            const double2 v0 = hp[0], v1 = hp[1], v2 = hp[2], v3 = hp[3];
                            h8[0] = v0.x; h8[1] = v0.y; h8[2] = v1.x; h8[3] = v1.y; h8[4] = v2.x; h8[5] = v2.y; h8[6] = v3.x; h8[7] = v3.y;
                        } else {
#pragma unroll
                            for (int r = 0; r < 8; r++) h8[r] = (cn + r < k) ? xcol[cn + r] : 0.0;
                        }
                    }
                    const int nkh = tile ? 2 : 1;                    // the remainder block is one half-block
                    if (nbt >= 0) {
                        // the next diagonal tile first: its transposition overlaps the other MMAs
#pragma unroll
                        for (int kh = 0; kh < nkh; kh++) {
                            const double a = gc[(cb + 4 * kh + tig) * KS + 8 * nbt + gid];
#pragma unroll
                            for (int ct = 0; ct < CT; ct++)
                                dmma(mu[nbt][ct][0], mu[nbt][ct][1], a, dsm[(4 * kh + tig) * DP + 8 * ct + gid]);
                        }
                        const double ri = rinv[8 * nbt + gid];
#pragma unroll
                        for (int ct = 0; ct < CT; ct++)
                            *reinterpret_cast<double2*>(tsm + gid * TP + 8 * ct + 2 * tig) = make_double2(mu[nbt][ct][0] * ri, mu[nbt][ct][1] * ri);
                    }
                    if (REM) {
                        // the remainder rows take the block's updates here, off the chain phase (each thread re-reads its column's d)
#pragma unroll
                        for (int c = 0; c < (tile ? 8 : 3); c++) {
                            const double dc = dsm[c * DP + colx];
#pragma unroll
                            for (int r = 0; r < 3; r++) mrem[r] = fma(vrem[(cb + c) * 4 + r], dc, mrem[r]);
                        }
                        if (nb == NT) {
#pragma unroll
                            for (int r = 0; r < 3; r++) pm[r] = mrem[r] * rinv[KT + r];
                        }
                    }
                    if (PRELOAD) {
#pragma unroll
                        for (int c = 0; c < 3; c++)
#pragma unroll
                            for (int r = c + 1; r < 8; r++) wn[widx(c, r)] = wl[nb * 32 + woff(c) + (r - c - 1)];
                    }
#pragma unroll
                    for (int kh = 0; kh < nkh; kh++) {
                        double bf[CT];
#pragma unroll
                        for (int ct = 0; ct < CT; ct++) bf[ct] = dsm[(4 * kh + tig) * DP + 8 * ct + gid];
#pragma unroll
                        for (int rt = 0; rt < NT; rt++) {
                            if (rt != nbt) {
                                const double a = gc[(cb + 4 * kh + tig) * KS + 8 * rt + gid];
#pragma unroll
                                for (int ct = 0; ct < CT; ct++) dmma(mu[rt][ct][0], mu[rt][ct][1], a, bf[ct]);
                            }
                        }
                    }
                }
                sched_sync(bid);
            }
            if (cont) t++;
            cont = cont && (flagbits < 0 || (0.0 > rel_tol));
        }
        if (have) my_sweeps += t;

        unsigned int g32 = 0;
        if (lane == 0 && has) g32 = atomicAdd(next_group, 1u);
        grp = has ? (int64_t)gridDim.x * WARPS + __shfl_sync(0xffffffffu, g32, 0) : ngroups;
    }
#pragma unroll
    for (int x = 16; x > 0; x >>= 1) my_sweeps += __shfl_xor_sync(0xffffffffu, my_sweeps, x);
    if (lane == 0 && my_sweeps) atomicAdd(sweeps, my_sweeps);
}

template <int NT, bool REM, int CT>
void launch(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, double l1,
            unsigned max_iter, double rel_tol, unsigned long long* sweeps, unsigned int* counter, cudaStream_t st)
{
    using L = Lay<NT, REM, CT>;
    const size_t smem = sizeof(double) * (size_t)L::TOTAL;
    auto kern = k_scd_phase<NT, REM, CT>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = ceil_div(ncol, L::NC);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(groups, 148));
    NNLM_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    kern<<<grid, 32 * WARPS, smem, st>>>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter);
    NNLM_LAUNCHED();
}

#define NNLM_SCDP_ARGS double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, \
    double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, unsigned int* counter, cudaStream_t st
#define NNLM_SCDP_PASS X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st

// explicit-instantiation entry points (several translation units keep the build parallel).
// nt = number of row tiles, rem = the rank has 1..3 coordinates beyond the last full tile (kept next to the chain)
void launch_ct4(int nt, bool rem, NNLM_SCDP_ARGS);      // 32-column groups, nt <= 6
void launch_ct2_lo(int nt, bool rem, NNLM_SCDP_ARGS);   // 16-column groups, nt 1..4
void launch_ct2_hi(int nt, bool rem, NNLM_SCDP_ARGS);   // 16-column groups, nt 5..8

}  // namespace scd_phase
}  // namespace nnlm
