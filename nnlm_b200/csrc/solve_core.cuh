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
template <int RPL, int METHOD>
__device__ __forceinline__ unsigned warp_solve_ls(double (&h)[RPL], const double (&q)[RPL], const unsigned (&mk)[RPL],
                                                  const double* gs, int k, double l1, unsigned max_iter, double rel_tol)
{
    constexpr int KR = 32 * RPL;
    const int lane = threadIdx.x & 31;
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
                for (int s = 0; s < RPL; s++) mu[s] = fma(gs[lane + 32 * s + KR * c], hc, mu[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < RPL; s++) { mu[s] -= q[s]; if (l1 != 0.0) mu[s] += l1; }

        for (; t < max_iter && cont; t++) {
            bool flag = false;
#pragma unroll
            for (int sc = 0; sc < RPL; sc++) {
                for (int lc = 0; lc < 32; lc++) {
                    const int c = 32 * sc + lc;
                    if (c >= k) break;
                    if ((mk[sc] >> lc) & 1u) continue;
                    const double hc = shfl_d(h[sc], lc);
                    const double muc = shfl_d(mu[sc], lc);
                    double cand = hc - muc / gs[c + KR * c];
                    if (cand < 0) cand = 0;
                    if (cand != hc) {
                        const double d = cand - hc;
#pragma unroll
                        for (int s = 0; s < RPL; s++) mu[s] = fma(d, gs[lane + 32 * s + KR * c], mu[s]);
                        const double num = 2 * fabs(hc - cand), den = cand + hc + TINY_NUM;
                        const bool over = (den > 0) ? (num > rel_tol * den) : (num / den > rel_tol);
                        flag = flag || over;
                        if (lane == lc) h[sc] = cand;
                    }
                }
            }
            cont = flag || (0.0 > rel_tol);
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
                    for (int s = 0; s < RPL; s++) part = fma(gs[lane + 32 * s + KR * c], h[s], part);
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
