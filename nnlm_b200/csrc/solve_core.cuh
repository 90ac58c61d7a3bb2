// solve_core.cuh — the warp-level non-negative least-squares column solvers shared by solve_ls.cu (dense A: one Gram
// for all columns) and solve_ls_missing.cu (NA path: one Gram per column).
//   METHOD 1: sequential coordinate descent, reference src/base_algorithms.cpp:3-37 (scd_ls_update), preceded by
//             mu = WtW*H.col(j) - Wt*A.col(j) (+beta2) of src/update_with_missing.cpp:39-41 / :109-111
//   METHOD 2: Lee-Seung multiplicative rule applied coordinate after coordinate, src/base_algorithms.cpp:40-68
//
// One warp owns one column. Lane l owns rows l, l+32, ... of h and mu in registers; the regularised Gram sits in
// shared memory, column-major with the row count padded to KR = 32*RPL so `mu += d * V[:,c]` is one conflict-free
// shared load + one DFMA per owned row. The coordinate loop is strictly sequential (each step sees the mu left by the
// previous one) exactly as the reference; all state is fp64 because the data-dependent control flow
// (`tmp != Hj(k)`, the relative-change exit) decides the sweep count that is returned as average_epoch.
//
// The exit test `rel_err > rel_tol` with rel_err = max_k 2|d|/(new+old+TINY) is evaluated without the division as
// 2|d| > rel_tol*(new+old+TINY) whenever the denominator is positive (always, for non-negative iterates); this can
// only differ from the quotient form when the quotient is within one ulp of rel_tol.
#pragma once
#include "common.cuh"

namespace nnlm {

// h: in/out column (rows lane+32*s); q: Wt*A.col(j) (same distribution); mk[s]: ballot of masked rows 32*s..32*s+31;
// gs: shared Gram, gs[r + KR*c] = V[r,c], rows >= k zero. Returns the number of sweeps performed.
// TRI: gs holds only the lower triangle of the (symmetric) Gram, V[r,c] at r(r+1)/2 + c for r >= c — half the shared memory per
// column, so the NA path keeps 16 instead of 10 columns in flight per SM at k = 50 (its solver is latency-bound); the
// price is an address select per load and a strided (conflicting) access for the rows below the diagonal.
template <int RPL, int METHOD, bool TRI = false>
__device__ __forceinline__ unsigned warp_solve_ls(double (&h)[RPL], const double (&q)[RPL], const unsigned (&mk)[RPL],
                                                  const double* gs, int k, double l1, unsigned max_iter, double rel_tol,
                                                  int ldg = 32 * RPL)
{
    // ldg: column pitch of the shared Gram. The dense solver pads every column to 32*RPL rows of zeros; the NA solver packs
    // the columns at pitch k (more columns in flight per SM): lanes whose row is >= k then read finite values of the next
    // column into a mu that is never consumed (their h is 0, their 1/V_rr is 0 and they never own a coordinate).
    const int KR = ldg;
    const int lane = threadIdx.x & 31;
    int rr[RPL], tb[RPL];                               // TRI: this lane's rows (clamped into the matrix) and their triangle offsets
#pragma unroll
    for (int s = 0; s < RPL; s++) { rr[s] = min(lane + 32 * s, k - 1); tb[s] = rr[s] * (rr[s] + 1) / 2; }
    auto gi = [&](int s, int c) -> int {                // index of V[row s of this lane, c]
        if (!TRI) return lane + 32 * s + KR * c;
        return rr[s] >= c ? tb[s] + c : c * (c + 1) / 2 + rr[s];
    };
    unsigned t = 0;
    bool cont = true;                                   // rel_err starts at 1 + rel_tol
    if (METHOD == 1) {
        // mu = V h - WtA (+ l1)
        double mu[RPL];
#pragma unroll
        for (int s = 0; s < RPL; s++) mu[s] = 0.0;
#pragma unroll
        for (int sc = 0; sc < RPL; sc++) {
            for (int lc = 0; lc < 32; lc++) {
                const int c = 32 * sc + lc;
                if (c >= k) break;
                const double hc = shfl_d(h[sc], lc);
#pragma unroll
                for (int s = 0; s < RPL; s++) mu[s] = fma(gs[gi(s, c)], hc, mu[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < RPL; s++) { mu[s] -= q[s]; if (l1 != 0.0) mu[s] += l1; }

        // Sweeps. Per step the dependent path is: owner lane forms cand = max(0, h_c - mu_c / V_cc) and d (a multiplication
        // with the reciprocal taken once per column; clamp on the bit pattern) -> ONE broadcast of d -> one FMA per owned row.
        // d = 0 stands for the reference's `tmp != Hj(k)` test (adding 0 * V leaves mu bit-identical), so the step has no
        // data-dependent branch; the exit test is evaluated by the owner lane only and combined once per sweep.
        // (A first version broadcast h_c and mu_c, divided on every lane and branched on the result: ~550 cycles per step.)
        double rinv[RPL];
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int r = lane + 32 * s;
            rinv[s] = (r < k) ? 1.0 / gs[TRI ? tb[s] + r : r + KR * r] : 0.0;
        }
        bool anymask = false;
#pragma unroll
        for (int s = 0; s < RPL; s++) anymask = anymask || mk[s] != 0u;
        if (!anymask) {
            // unmasked column (the common case): no per-step branches at all. The Gram column of step c+1 is fetched
            // during step c; the owner lane's h and its exit test are selects / sign bits, not divergent code.
            const double tolh = 0.5 * rel_tol, c0 = tolh * TINY_NUM;
            for (; t < max_iter && cont; t++) {
                int flagi = 0;
#pragma unroll
                for (int sc = 0; sc < RPL; sc++) {
                    const int cnt = min(32, k - 32 * sc);
                    double gn[RPL];
#pragma unroll
                    for (int s = 0; s < RPL; s++) gn[s] = gs[gi(s, min(32 * sc, k - 1))];
                    for (int lc = 0; lc < cnt; lc++) {
                        const int c = 32 * sc + lc;
                        double g[RPL];
#pragma unroll
                        for (int s = 0; s < RPL; s++) g[s] = gn[s];
                        const int cn = (lc + 1 < cnt) ? c + 1 : c;
#pragma unroll
                        for (int s = 0; s < RPL; s++) gn[s] = gs[gi(s, cn)];
                        const double hc = h[sc];
                        double cand = fma(-mu[sc], rinv[sc], hc);
                        const int keep = ~(__double2hiint(cand) >> 31);                       // max(cand, +0) on the bit pattern
                        cand = __hiloint2double(__double2hiint(cand) & keep, __double2loint(cand) & keep);
                        const double down = cand - hc;
                        const double d = shfl_d(down, lc);
#pragma unroll
                        for (int s = 0; s < RPL; s++) mu[s] = fma(d, g[s], mu[s]);
                        const bool own = lane == lc;
                        h[sc] = own ? cand : hc;
                        // 2|d| > tol (new + old + 1e-16)  <=>  (tol/2)(new + old) + (tol/2)1e-16 - |d| < 0
                        const int over = __double2hiint(fma(cand + hc, tolh, c0 - fabs(down)));
                        flagi |= own ? over : 0;
                    }
                }
                cont = __any_sync(0xffffffffu, flagi < 0) || (0.0 > rel_tol);
            }
        }
        for (; anymask && t < max_iter && cont; t++) {
            bool flag = false;
#pragma unroll
            for (int sc = 0; sc < RPL; sc++) {
#pragma unroll 4
                for (int lc = 0; lc < 32; lc++) {
                    const int c = 32 * sc + lc;
                    if (c >= k) break;
                    if ((mk[sc] >> lc) & 1u) continue;
                    const double hc = h[sc];
                    double cand = fma(-mu[sc], rinv[sc], hc);
                    cand = __hiloint2double(__double2hiint(cand) & ~(__double2hiint(cand) >> 31),
                                            __double2loint(cand) & ~(__double2hiint(cand) >> 31));      // max(cand, +0)
                    const double down = cand - hc;
                    const double d = shfl_d(down, lc);
#pragma unroll
                    for (int s = 0; s < RPL; s++) mu[s] = fma(d, gs[gi(s, c)], mu[s]);
                    if (lane == lc) {
                        h[sc] = cand;
                        flag = flag || (2 * fabs(down) > rel_tol * (cand + hc + TINY_NUM));
                    }
                }
            }
            cont = __any_sync(0xffffffffu, flag) || (0.0 > rel_tol);
        }
    } else {
        for (; t < max_iter && cont; t++) {
            bool flag = false;
#pragma unroll
            for (int sc = 0; sc < RPL; sc++) {
                for (int lc = 0; lc < 32; lc++) {
                    const int c = 32 * sc + lc;
                    if (c >= k) break;
                    if ((mk[sc] >> lc) & 1u) continue;
                    double part = 0.0;
#pragma unroll
                    for (int s = 0; s < RPL; s++) part = fma(gs[gi(s, c)], h[s], part);
                    const double den = warp_sum(part) + l1;
                    const double ratio = shfl_d(q[sc], lc) / (den + TINY_NUM);
                    if (lane == lc) h[sc] *= ratio;
                    const double e = 2 * fabs(ratio - 1) / (ratio + 1);
                    flag = flag || (e > rel_tol);
                }
            }
            cont = flag || (0.0 > rel_tol);
        }
    }
    return t;
}

}  // namespace nnlm
