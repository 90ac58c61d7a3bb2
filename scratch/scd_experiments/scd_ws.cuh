// scd_ws.cuh — K3/K4, the warp-specialised sequential-coordinate-descent solver for the square loss, rank k <= 64
// (reference src/base_algorithms.cpp:3-37 preceded by mu = WtW*h - WtA (+beta2), src/update_with_missing.cpp:39-41).
//
// Why warp-specialised. The solver has two kinds of work that want the same fp64 pipe of a scheduler (SMSP):
//   * per column, a strictly sequential chain over the coordinates (one DFMA + DADD + two selects per step);
//   * per block of 8 coordinates, a rank-8 update of mu = V h - q for all rows and all columns, which runs as
//     DMMA.8x8x4 at the full fp64 rate (64 FMA/clk/SM) but occupies the pipe for 16 cycles per instruction.
// Measured on B200 (scratch/smsp_map.cu): a dependent DFMA chain takes 8 cycles per step alone or next to DMMA warps on
// OTHER schedulers, and 56 cycles per step when three DMMA warps share its scheduler. In the single-role kernels
// (scd_dmma.cuh, scd_chain.cuh) every chain instruction queued behind the other warps' MMAs: pipe 55-60 % busy, the rest
// waiting. Here the roles are pinned to schedulers: warp w runs on scheduler w % 4, so
//   * warps 0, 4, 8, 12 (scheduler 0) are CHAIN warps: lane = column, they own h and the sequential steps and never
//     issue an MMA;
//   * the other 12 warps (schedulers 1-3) are MMA warps: each owns the mu tiles (DMMA C fragments) of one group of
//     8*CT columns and does nothing but MMAs.
// Hand-off per block b of 8 coordinates, through shared memory and mbarriers (one chain warp serves three MMA warps):
//   MMA warp  -> chain warp : the diagonal tile mu[8b..8b+7, columns]                    (tsm, barrier full_t)
//   chain warp: P_r = h_r - mu_r / V_rr; eight steps  cand = P_c, d_c = clamp(cand) - h_c, P_r -= (V_rc/V_rr) d_c (r > c)
//   chain warp -> MMA warp  : the eight d                                                (dsm, barriers full_d[2])
//   MMA warp  : mu[next diagonal tile] += V d first (the only part the next chain waits for), hand it over, then the
//               other row tiles while the chain warp is busy with its other groups.
// Rows beyond the last full tile (k mod 8 in 1..3, e.g. k = 50) are kept by the chain warp itself, one thread per
// column (3 extra FMAs per step with V from shared memory), and their coordinates form a short last block that needs
// no tile from the MMA warp: at k = 50 an MMA warp holds 6 row tiles of 32 columns in 96 registers — all 16 warps fit
// at 128 registers — and a sweep costs 6 x 13 MMAs per 8 columns, the unpadded count. h stays in X (L2-resident), read
// and written by the chain warp 64 B per thread and block.
// Arithmetic differences from the reference, all at rounding level: reciprocal and pre-multiplied V_rc/V_rr instead of
// a division per step; inside a block the candidate of coordinate r is h_r - mu_r/V_rr - sum_c (V_rc/V_rr) d_c instead
// of going through mu; `tmp != Hj(k)` becomes d = 0; the exit test 2|d|/(h_new+h_old+1e-16) > tol is evaluated as
// (tol/2)(h_new+h_old) + (tol/2)1e-16 - |d| < 0; rows outside the block see its 8 updates summed 4 at a time inside
// the MMA. Control flow per column is the reference's (a column stops when its max relative change <= rel_tol or at
// max_iter; finished columns are frozen while the rest of the group keeps sweeping); sweep counts are summed.
#pragma once
#include <algorithm>

#include "kernels.cuh"

namespace nnlm {
namespace scd_ws {

constexpr int WARPS = 16, SLOTS = 12;     // 4 chain warps + 12 MMA warps; one group of columns per MMA warp ("slot")

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip_sign(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// bounded wait: a protocol error traps instead of hanging the device
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_addr(bar);
    uint32_t done = 0;
    for (unsigned spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();
    }
}

// offset of step c's multipliers inside a block's 32-entry record (7-c used entries, padded to an even count)
__host__ __device__ constexpr int woff(int c) { return c == 0 ? 0 : c == 1 ? 8 : c == 2 ? 14 : c == 3 ? 20 : c == 4 ? 24 : c == 5 ? 28 : 30; }

// shared-memory layout (doubles unless noted)
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
    static constexpr int SLOT0 = VREM + (REM ? KB * 4 : 0);
    static constexpr int DSM = 0, TSM = 16 * DP, FZ = TSM + 8 * TP, CTL = FZ + 64, BAR = CTL + 2, SLOT = BAR + 4;   // per slot
    static constexpr int TOTAL = SLOT0 + SLOTS * SLOT;
};

template <int NT, bool REM, int CT>
__global__ void __launch_bounds__(32 * WARPS, 1)
k_scd_ws(double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ Qp, int splits,
         const uint8_t* __restrict__ mask, int k, int64_t ncol, double l1, unsigned max_iter, double rel_tol,
         unsigned long long* __restrict__ sweeps)
{
    using L = Lay<NT, REM, CT>;
    constexpr int KT = L::KT, KB = L::KB, KS = L::KS, NC = L::NC, NBLK = L::NBLK, DP = L::DP, TP = L::TP;
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
    if (threadIdx.x < SLOTS) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::SLOT0 + threadIdx.x * L::SLOT + L::BAR);
        bar_init(bars + 0, 1); bar_init(bars + 1, 1); bar_init(bars + 2, 1);      // full_d[0], full_d[1], full_t
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NBLK * 64; e += 32 * WARPS) {
        const int b = e >> 6, c = (e >> 3) & 7, r = e & 7;
        const int rr = 8 * b + r, cc = 8 * b + c;
        if (r > c && rr < k && cc < k) wl[b * 32 + woff(c) + (r - c - 1)] = rinv[rr] * G[rr + k * cc];
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ngroups = (ncol + NC - 1) / NC;
    const int64_t gstride = (int64_t)gridDim.x * SLOTS;

    if ((warp & 3) != 0) {
        // =============================== MMA warp: slot s, scheduler 1 + s % 3 ===============================
        const int s = 3 * (warp >> 2) + (warp & 3) - 1;
        const int gid = lane >> 2, tig = lane & 3;
        double* slot = sm + L::SLOT0 + s * L::SLOT;
        double* dsm = slot + L::DSM;
        double* tsm = slot + L::TSM;
        const int* ctl = reinterpret_cast<const int*>(slot + L::CTL);
        uint64_t* bars = reinterpret_cast<uint64_t*>(slot + L::BAR);
        unsigned hd = 0;                                            // hand-offs received on full_d so far
        for (int64_t grp = (int64_t)blockIdx.x + (int64_t)gridDim.x * s; grp < ngroups; grp += gstride) {
            const int64_t col0 = grp * NC;
            const int cnt = (int)min((int64_t)NC, ncol - col0);
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
            // the first diagonal tile, one thread per column
#pragma unroll
            for (int ct = 0; ct < CT; ct++)
                *reinterpret_cast<double2*>(tsm + gid * TP + 8 * ct + 2 * tig) = make_double2(mu[0][ct][0], mu[0][ct][1]);
            __syncwarp();
            if (lane == 0) bar_arrive(bars + 2);

            // ---- sweeps: until the chain warp says stop ----
            bool stop = false;
            while (!stop) {
#pragma unroll
                for (int b = 0; b < NBLK; b++) {
                    const int buf = hd & 1;
                    bar_wait(bars + buf, (hd >> 1) & 1);
                    hd++;
                    if (ctl[buf] != 0) { stop = true; break; }
                    const double* dcur = dsm + buf * 8 * DP;
                    const int nkh = (REM && b == NT) ? 1 : 2;              // the remainder block is one half-block
                    // next tile-backed block: b+1, or 0 after the last block of the sweep; none before the remainder block
                    const int nbt = (b + 1 < NT) ? b + 1 : ((REM && b == NT - 1) ? -1 : 0);
                    if (nbt >= 0) {
#pragma unroll
                        for (int kh = 0; kh < 2; kh++) {
                            if (kh < nkh) {
                                const double a = gc[(8 * b + 4 * kh + tig) * KS + 8 * nbt + gid];
#pragma unroll
                                for (int ct = 0; ct < CT; ct++)
                                    dmma(mu[nbt][ct][0], mu[nbt][ct][1], a, dcur[(4 * kh + tig) * DP + 8 * ct + gid]);
                            }
                        }
#pragma unroll
                        for (int ct = 0; ct < CT; ct++)
                            *reinterpret_cast<double2*>(tsm + gid * TP + 8 * ct + 2 * tig) = make_double2(mu[nbt][ct][0], mu[nbt][ct][1]);
                        __syncwarp();
                        if (lane == 0) bar_arrive(bars + 2);
                    }
#pragma unroll
                    for (int kh = 0; kh < 2; kh++) {
                        if (kh < nkh) {
                            double bf[CT];
#pragma unroll
                            for (int ct = 0; ct < CT; ct++) bf[ct] = dcur[(4 * kh + tig) * DP + 8 * ct + gid];
#pragma unroll
                            for (int rt = 0; rt < NT; rt++) {
                                if (rt != nbt) {
                                    const double a = gc[(8 * b + 4 * kh + tig) * KS + 8 * rt + gid];
#pragma unroll
                                    for (int ct = 0; ct < CT; ct++) dmma(mu[rt][ct][0], mu[rt][ct][1], a, bf[ct]);
                                }
                            }
                        }
                    }
                }
            }
        }
        return;
    }

    // ================================= chain warp ci, scheduler 0: serves slots ci, ci+4, ci+8 =================================
    const int ci = warp >> 2;
    const int colx = lane < NC ? lane : NC - 1;
    const double tolh = 0.5 * rel_tol, c0 = tolh * TINY_NUM;
    unsigned long long my_sweeps = 0;
    unsigned hd[3] = {0, 0, 0}, ht[3] = {0, 0, 0};                  // hand-offs sent on full_d / received on full_t, per slot

    for (int64_t base = (int64_t)blockIdx.x + (int64_t)gridDim.x * ci; base < ngroups; base += gstride) {
        // ---- per-group state of this round ----
        bool act[3], cont[3];
        unsigned t[3];
        int64_t colg[3];                                            // this lane's column (element offset of its h), per group
        double mrem[3][3];                                          // mu of the remainder rows (REM)
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int64_t grp = base + (int64_t)gridDim.x * 4 * p;
            act[p] = grp < ngroups;
            cont[p] = false; t[p] = 0; colg[p] = 0;
#pragma unroll
            for (int r = 0; r < 3; r++) mrem[p][r] = 0.0;
            if (!act[p]) continue;
            const int s = ci + 4 * p;
            double* slot = sm + L::SLOT0 + s * L::SLOT;
            unsigned long long* fzs = reinterpret_cast<unsigned long long*>(slot + L::FZ);
            const int64_t col0 = grp * NC;
            const int cnt = (int)min((int64_t)NC, ncol - col0);
            const bool have = lane < cnt;
            colg[p] = (col0 + (have ? lane : 0)) * k;
            // coordinate mask of this lane's column: bit c set = leave coordinate c alone
            unsigned long long mb[2] = {0, 0};
            bool any_free = false;
#pragma unroll
            for (int w = 0; w < 2; w++) {
                if (mask != nullptr && have) {
                    const uint8_t* mc = mask + (col0 + lane) * k;
#pragma unroll 1
                    for (int r = 64 * w; r < k && r < 64 * w + 64; r++) mb[w] |= (unsigned long long)(mc[r] != 0) << (r - 64 * w);
                }
                const int kw = k - 64 * w;
                const unsigned long long kmask = kw >= 64 ? ~0ull : (kw <= 0 ? 0ull : ((1ull << kw) - 1ull));
                any_free = any_free || (mb[w] & kmask) != kmask;
                mb[w] &= kmask;                                      // padding coordinates have h = mu = 0: d = 0 by itself
            }
            cont[p] = have && any_free;                              // fully masked column: src/update_with_missing.cpp:33-34
            fzs[lane] = mb[0];
            fzs[32 + lane] = mb[1];
            if (REM) {
                // mu of the remainder rows: l1 - q + V[rem, :] h
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    double q = 0.0;
                    if (have && KT + r < k) {
#pragma unroll 1
                        for (int sp = 0; sp < splits; sp++) q += Qp[((int64_t)sp * ncol + col0 + lane) * k + KT + r];
                    }
                    mrem[p][r] = (have && KT + r < k) ? l1 - q : 0.0;
                }
                if (have) {
#pragma unroll 1
                    for (int c = 0; c < k; c++) {
                        const double hc = X[colg[p] + c];
#pragma unroll
                        for (int r = 0; r < 3; r++) mrem[p][r] = fma(vrem[c * 4 + r], hc, mrem[p][r]);
                    }
                }
            }
        }
        __syncwarp();

        // ---- sweeps ----
        for (unsigned it = 0; it < max_iter; it++) {
            bool galive[3];
            bool any = false;
#pragma unroll
            for (int p = 0; p < 3; p++) { galive[p] = act[p] && __any_sync(0xffffffffu, cont[p]); any = any || galive[p]; }
            if (!any) break;
            int flagbits[3] = {0, 0, 0};
#pragma unroll 1
            for (int b = 0; b < NBLK; b++) {
#pragma unroll
                for (int p = 0; p < 3; p++) {
                    if (!galive[p]) continue;
                    const int s = ci + 4 * p;
                    double* slot = sm + L::SLOT0 + s * L::SLOT;
                    double* tsm = slot + L::TSM;
                    const unsigned long long* fzs = reinterpret_cast<const unsigned long long*>(slot + L::FZ);
                    int* ctl = reinterpret_cast<int*>(slot + L::CTL);
                    uint64_t* bars = reinterpret_cast<uint64_t*>(slot + L::BAR);
                    const bool tile = !REM || b < NT;                 // tile-backed block (8 steps) or the remainder block (3 steps)
                    const bool mine = lane < NC && cont[p];           // lanes that own a live column (others compute d = 0)
                    const int cb = 8 * b;
                    // h of the block, straight from X (in flight while the tile arrives)
                    double h8[8];
#pragma unroll
                    for (int r = 0; r < 8; r++) h8[r] = 0.0;
                    if (mine) {
                        if (cb + 8 <= k && (reinterpret_cast<uintptr_t>(X + colg[p] + cb) & 15) == 0) {
                            const double2* hp = reinterpret_cast<const double2*>(X + colg[p] + cb);
                            const double2 v0 = hp[0], v1 = hp[1], v2 = hp[2], v3 = hp[3];
                            h8[0] = v0.x; h8[1] = v0.y; h8[2] = v1.x; h8[3] = v1.y; h8[4] = v2.x; h8[5] = v2.y; h8[6] = v3.x; h8[7] = v3.y;
                        } else {
#pragma unroll
                            for (int r = 0; r < 8; r++) if (cb + r < k) h8[r] = X[colg[p] + cb + r];
                        }
                    }
                    const unsigned fz8 = (unsigned)(fzs[(cb >> 6) * 32 + colx] >> (cb & 63)) & 0xffu;
                    double P[8], dd[8];
                    if (tile) {
                        bar_wait(bars + 2, ht[p] & 1);
                        ht[p]++;
#pragma unroll
                        for (int r = 0; r < 8; r++) P[r] = fma(-tsm[r * TP + colx], rinv[cb + r], h8[r]);
                    } else {
#pragma unroll
                        for (int r = 0; r < 8; r++) P[r] = (r < 3) ? fma(-mrem[p][r], rinv[cb + r], h8[r]) : 0.0;
                    }
                    const double* wb = wl + b * 32;
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        if (c >= 3 && !tile) { dd[c] = 0.0; continue; }
                        const double cand = P[c], hc = h8[c];
                        const bool neg = __double2hiint(cand) < 0;
                        const bool live = mine && !((fz8 >> c) & 1u);
                        const double dpos = cand - hc;
                        double d = neg ? flip_sign(hc) : dpos;             // integer sign flip: keeps the fp64 pipe for the FMAs
                        double hn = neg ? 0.0 : cand;
                        d = live ? d : 0.0;
                        hn = live ? hn : hc;
                        dd[c] = d;
                        h8[c] = hn;
#pragma unroll
                        for (int r = c + 1; r < 8; r++) P[r] = fma(-wb[woff(c) + (r - c - 1)], d, P[r]);
                        if (REM) {
#pragma unroll
                            for (int r = 0; r < 3; r++) mrem[p][r] = fma(vrem[(cb + c) * 4 + r], d, mrem[p][r]);
                        }
                        // 2|d| > tol (hn + hc + 1e-16)  <=>  (tol/2)(hn + hc) + (tol/2)1e-16 - |d| < 0 : collect the sign bits
                        flagbits[p] |= __double2hiint(fma(hn + hc, tolh, c0 - fabs(d)));
                    }
                    // publish d (all lanes of the tile width: dead columns publish zeros) and the new h
                    const int buf = hd[p] & 1;
                    double* dcur = slot + L::DSM + buf * 8 * DP;
                    if (lane < NC) {
#pragma unroll
                        for (int c = 0; c < 8; c++) dcur[c * DP + lane] = dd[c];
                    }
                    if (mine) {
                        if (cb + 8 <= k && (reinterpret_cast<uintptr_t>(X + colg[p] + cb) & 15) == 0) {
                            double2* hp = reinterpret_cast<double2*>(X + colg[p] + cb);
                            hp[0] = make_double2(h8[0], h8[1]); hp[1] = make_double2(h8[2], h8[3]);
                            hp[2] = make_double2(h8[4], h8[5]); hp[3] = make_double2(h8[6], h8[7]);
                        } else {
#pragma unroll
                            for (int r = 0; r < 8; r++) if (cb + r < k) X[colg[p] + cb + r] = h8[r];
                        }
                    }
                    if (lane == 0) ctl[buf] = 0;
                    __syncwarp();
                    if (lane == 0) bar_arrive(bars + buf);
                    hd[p]++;
                }
            }
#pragma unroll
            for (int p = 0; p < 3; p++) {
                if (!galive[p]) continue;
                if (cont[p]) t[p]++;
                cont[p] = cont[p] && (flagbits[p] < 0 || (0.0 > rel_tol));
            }
            __syncwarp();
        }
        // ---- tell the MMA warps of this round to move on (each has one diagonal tile pending) ----
#pragma unroll
        for (int p = 0; p < 3; p++) {
            if (!act[p]) continue;
            double* slot = sm + L::SLOT0 + (ci + 4 * p) * L::SLOT;
            int* ctl = reinterpret_cast<int*>(slot + L::CTL);
            uint64_t* bars = reinterpret_cast<uint64_t*>(slot + L::BAR);
            bar_wait(bars + 2, ht[p] & 1);
            ht[p]++;
            const int buf = hd[p] & 1;
            if (lane == 0) ctl[buf] = 1;
            __syncwarp();
            if (lane == 0) bar_arrive(bars + buf);
            hd[p]++;
            my_sweeps += t[p];
        }
    }
#pragma unroll
    for (int x = 16; x > 0; x >>= 1) my_sweeps += __shfl_xor_sync(0xffffffffu, my_sweeps, x);
    if (lane == 0 && my_sweeps) atomicAdd(sweeps, my_sweeps);
}

template <int NT, bool REM, int CT>
void launch(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, double l1,
            unsigned max_iter, double rel_tol, unsigned long long* sweeps, cudaStream_t st)
{
    using L = Lay<NT, REM, CT>;
    const size_t smem = sizeof(double) * (size_t)L::TOTAL;
    auto kern = k_scd_ws<NT, REM, CT>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = ceil_div(ncol, L::NC);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(groups, 148));
    kern<<<grid, 32 * WARPS, smem, st>>>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps);
    NNLM_LAUNCHED();
}

#define NNLM_SCDW_ARGS double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol, \
    double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, cudaStream_t st
#define NNLM_SCDW_PASS X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st

// explicit-instantiation entry points (several translation units keep the build parallel).
// nt = number of row tiles, rem = the rank has 1..3 coordinates beyond the last full tile (kept by the chain warps)
void launch_ct4(int nt, bool rem, NNLM_SCDW_ARGS);      // 32-column groups, nt <= 6
void launch_ct2_lo(int nt, bool rem, NNLM_SCDW_ARGS);   // 16-column groups, nt 1..4
void launch_ct2_hi(int nt, bool rem, NNLM_SCDW_ARGS);   // 16-column groups, nt 5..8

}  // namespace scd_ws
}  // namespace nnlm
