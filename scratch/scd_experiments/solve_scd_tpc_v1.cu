// solve_scd_tpc.cu — K3/K4, the throughput-oriented sequential-coordinate-descent solver for the square loss
// (reference src/base_algorithms.cpp:3-37 preceded by mu = WtW*h - WtA (+beta2), src/update_with_missing.cpp:39-41).
//
// The coordinate loop is strictly sequential inside one column (each step sees the mu left by the previous one), but
// columns are independent (src/update_with_missing.cpp:29-30). The warp-per-column layout of solve_core.cuh keeps one
// column's k values across the lanes and is bound by the shuffle -> divide -> broadcast latency chain (measured: 58 % of
// the ANLS step at 50000 x 10000, k = 50). Here ONE THREAD owns ONE column:
//   * mu[k] lives in registers with compile-time indices (the sweep over c is fully unrolled),
//   * h[k] lives in shared memory as hs[c][lane] (one conflict-free 8-byte access per step),
//   * the regularised Gram is read from shared memory as warp-wide broadcasts (all lanes are at the same coordinate c),
//   * the rank-1 update mu += d * V[:,c] is k independent DFMAs per thread: the fp64 pipe is the bound, not latency.
// The division mu_c / V_cc is replaced by a multiplication with the reciprocal computed once per half-iteration
// (<= 1 ulp difference per step, far inside the 1e-5 parity bound; the exit test was already division-free).
// A step whose d is zero on every lane of the warp skips the rank-1 update (the reference's `tmp != Hj(k)` branch).
// Control flow per column is identical to the reference: a column stops sweeping when its max relative change drops to
// rel_tol or at max_iter; its sweep count is summed into total_raw_iter.
#include <algorithm>

#include "kernels.cuh"

namespace nnlm {

namespace {

template <int KB, int TPC_WARPS>
__global__ void __launch_bounds__(32 * TPC_WARPS, 1)
k_scd_tpc(double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ Qp, int splits,
          const uint8_t* __restrict__ mask, int k, int64_t ncol, double l1, unsigned max_iter, double rel_tol,
          unsigned long long* __restrict__ sweeps)
{
    extern __shared__ __align__(16) double sm[];
    double* gs = sm;                          // [KB][KB], gs[r + KB*c] = V[r,c]; padding rows/cols are zero
    double* rinv = gs + KB * KB;              // [KB] 1 / V[c,c]
    double* hs = rinv + KB + (threadIdx.x >> 5) * (KB * 32);   // this warp's [KB][32] slab (swizzled columns)

    for (int e = threadIdx.x; e < KB * KB; e += 32 * TPC_WARPS) {
        const int c = e / KB, r = e % KB;
        gs[e] = (r < k && c < k) ? G[r + k * c] : 0.0;
    }
    for (int c = threadIdx.x; c < KB; c += 32 * TPC_WARPS) rinv[c] = (c < k) ? 1.0 / G[c + k * c] : 0.0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t ngroups = (ncol + 31) / 32;
    // groups are dealt round-robin over the CTAs first, then over the warps of a CTA, so every SM gets the same load
    const int64_t gwarp = (int64_t)blockIdx.x + (int64_t)gridDim.x * (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * TPC_WARPS;
    unsigned long long my_sweeps = 0;
    // slot of (row r, column-in-group col): XOR swizzle keeps both the staging pass (consecutive r, fixed col) and the
    // per-step access (fixed r, consecutive lanes) free of bank conflicts
    auto slot = [](int r, int col) { return r * 32 + (col ^ (r & 31)); };

    for (int64_t grp = gwarp; grp < ngroups; grp += nwarps) {
        const int64_t col0 = grp * 32;
        const int cnt = (int)min((int64_t)32, ncol - col0);
        const bool have = lane < cnt;
        const int total = cnt * k;                 // contiguous doubles of this group in X / Qp

        // ---- q = sum of split-K partials, staged coalesced, then mu = -q (+ l1) ----
        for (int e = lane; e < total; e += 32) {
            double acc = 0.0;
            for (int sp = 0; sp < splits; sp++) acc += Qp[((int64_t)sp * ncol + col0) * k + e];
            hs[slot(e % k, e / k)] = acc;
        }
        __syncwarp();
        double mu[KB];
#pragma unroll
        for (int r = 0; r < KB; r++) mu[r] = (have && r < k) ? (l1 - hs[slot(r, lane)]) : 0.0;
        __syncwarp();
        // ---- h staged the same way; mask bits per thread ----
        for (int e = lane; e < total; e += 32) hs[slot(e % k, e / k)] = X[col0 * k + e];
        unsigned long long mbits = 0;
        if (mask != nullptr && have) {
            const uint8_t* mc = mask + (col0 + lane) * k;
            for (int r = 0; r < k; r++) mbits |= (unsigned long long)(mc[r] != 0) << r;
        }
        __syncwarp();
        if (!have) {
#pragma unroll 1
            for (int r = 0; r < k; r++) hs[slot(r, lane)] = 0.0;
        }
        const unsigned long long kmask = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
        const bool all_masked = (mbits & kmask) == kmask;            // src/update_with_missing.cpp:33-34
        // ---- mu += V h ----
#pragma unroll
        for (int c = 0; c < KB; c++) {
            if (c < k) {
                const double hc = hs[slot(c, lane)];
                const double2* gc = reinterpret_cast<const double2*>(gs + KB * c);
#pragma unroll
                for (int r2 = 0; r2 < KB / 2; r2++) {
                    const double2 g = gc[r2];
                    mu[2 * r2] = fma(g.x, hc, mu[2 * r2]);
                    mu[2 * r2 + 1] = fma(g.y, hc, mu[2 * r2 + 1]);
                }
            }
        }

        // ---- sweeps ----
        bool cont = have && !all_masked;        // rel_err starts at 1 + rel_tol
        unsigned t = 0;
        for (unsigned it = 0; it < max_iter; it++) {
            if (!__any_sync(0xffffffffu, cont)) break;
            bool flag = false;
#pragma unroll
            for (int c = 0; c < KB; c++) {
                if (c < k) {
                    const int sl = slot(c, lane);
                    const double hc = hs[sl];
                    double cand = fma(-mu[c], rinv[c], hc);
                    if (cand < 0) cand = 0;
                    const bool frozen = !cont || ((mbits >> c) & 1ull);
                    const double d = frozen ? 0.0 : cand - hc;
                    if (__any_sync(0xffffffffu, d != 0.0)) {
                        const double2* gc = reinterpret_cast<const double2*>(gs + KB * c);
#pragma unroll
                        for (int r2 = 0; r2 < KB / 2; r2++) {
                            const double2 g = gc[r2];
                            mu[2 * r2] = fma(d, g.x, mu[2 * r2]);
                            mu[2 * r2 + 1] = fma(d, g.y, mu[2 * r2 + 1]);
                        }
                        if (d != 0.0) {
                            hs[sl] = cand;
                            flag = flag || (2 * fabs(d) > rel_tol * (cand + hc + TINY_NUM));
                        }
                    }
                }
            }
            if (cont) t++;
            cont = cont && (flag || (0.0 > rel_tol));
        }
        my_sweeps += t;
        __syncwarp();
        for (int e = lane; e < total; e += 32) X[col0 * k + e] = hs[slot(e % k, e / k)];
        __syncwarp();
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) my_sweeps += __shfl_xor_sync(0xffffffffu, my_sweeps, s);
    if (lane == 0 && my_sweeps) atomicAdd(sweeps, my_sweeps);
}

template <int KB>
void launch_kb(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol,
               double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, cudaStream_t st)
{
    constexpr int W = (KB <= 56) ? 12 : 8;           // 3 (2) warps per scheduler; bounded by the 227 KB of shared memory
    const size_t smem = sizeof(double) * ((size_t)KB * KB + KB + (size_t)W * KB * 32);
    auto kern = k_scd_tpc<KB, W>;
    NNLM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = ceil_div(ncol, 32);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(groups, 148));
    kern<<<grid, 32 * W, smem, st>>>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps);
    NNLM_LAUNCHED();
}

}  // namespace

bool scd_tpc_supported(int k) { return k >= 1 && k <= 64; }
size_t scd_tpc_scratch_doubles() { return 2; }

void launch_scd_tpc(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol,
                    double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, double* /*scratch*/, cudaStream_t st)
{
    NNLM_REQUIRE(scd_tpc_supported(k), "thread-per-column SCD supports rank k <= 64");
    if (ncol <= 0) return;
    const int kb = (k + 7) / 8;
    switch (kb) {
        case 1: launch_kb<8>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
        case 2: launch_kb<16>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
        case 3: launch_kb<24>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
        case 4: launch_kb<32>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
        case 5: launch_kb<40>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
        case 6: launch_kb<48>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
        case 7: launch_kb<56>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
        default: launch_kb<64>(X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, st); break;
    }
}

}  // namespace nnlm
