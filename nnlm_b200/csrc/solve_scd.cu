// solve_scd.cu — dispatch of the blocked DMMA SCD solvers over tile widths and padded ranks:
// scd_chain.cuh (second generation) for k <= 64, scd_dmma.cuh (first generation) for k in (64, 128].
#include <cstdlib>

#include "scd_chain.cuh"
#include "scd_dmma.cuh"

namespace nnlm {

bool scd_tpc_supported(int k) { return k >= 1 && k <= 128; }
size_t scd_tpc_scratch_doubles() { return 2; }

void launch_scd_tpc(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol,
                    double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, double* scratch, cudaStream_t st)
{
    NNLM_REQUIRE(scd_tpc_supported(k), "the blocked SCD solver supports rank k <= 128");
    if (ncol <= 0) return;
    unsigned int* counter = reinterpret_cast<unsigned int*>(scratch);
    const int nb = (k + 7) / 8;
    if (nb > 8) {          // k in (64, 128]: the Gram alone takes up to 135 KB of shared memory -> 8-column tiles, 8 warps
        if (nb <= 12) scd_dmma::launch_ct1_big_a(nb, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
        else scd_dmma::launch_ct1_big_b(nb, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
        return;
    }
    // Tile width, measured at 50000 x 10000, k = 50 with scd_chain (bench.py, ANLS iterations/s with 8- / 16- / 32-column
    // tiles on both halves: 371 / 458 / 418): 16 columns per warp keeps 12 warps per SM resident at 168 registers and
    // amortises the sequential part over twice the columns of the 8-column tile; 32 columns need 255 registers (8 warps)
    // and run two unbalanced rounds. Shards with fewer than two 16-column groups per SM take 8-column tiles, so that every
    // SM still gets work. NNLM_SCD_CT / NNLM_SCD_IMPL override (experiments).
    static const int force_ct = [] { const char* e = getenv("NNLM_SCD_CT"); return e ? atoi(e) : 0; }();
    static const int impl = [] { const char* e = getenv("NNLM_SCD_IMPL"); return e ? atoi(e) : 2; }();   // 1: scd_dmma (first generation, kept for A/B), 2: scd_chain
    if (impl == 2) {
        int ct = ncol >= (int64_t)16 * 148 * 2 ? 2 : 1;
        if (force_ct == 1 || force_ct == 2 || force_ct == 4) ct = force_ct;
        const int nh = (k + 3) / 4;
        auto fn = ct == 4 ? (nh <= 8 ? scd_chain::launch_ct4_lo : scd_chain::launch_ct4_hi)
                : ct == 2 ? (nh <= 8 ? scd_chain::launch_ct2_lo : scd_chain::launch_ct2_hi)
                          : (nh <= 8 ? scd_chain::launch_ct1_lo : scd_chain::launch_ct1_hi);
        fn(nh, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
        return;
    }
    int ct = ceil_div(ncol, 16) >= 148 * 12 ? 2 : 1;
    if (force_ct == 1 || force_ct == 2 || force_ct == 4) ct = force_ct;
    if (ct == 4) scd_dmma::launch_ct4(nb, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
    else if (ct == 2) scd_dmma::launch_ct2(nb, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
    else scd_dmma::launch_ct1(nb, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
}

}  // namespace nnlm
