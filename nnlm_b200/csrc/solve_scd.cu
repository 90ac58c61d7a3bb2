// solve_scd.cu — dispatch of the blocked DMMA SCD solver (scd_chain.cuh) over tile widths and padded ranks, k <= 128.
#include <cstdlib>

#include "scd_chain.cuh"

namespace nnlm {

bool scd_tpc_supported(int k) { return k >= 1 && k <= 128; }
size_t scd_tpc_scratch_doubles() { return 2; }

void launch_scd_tpc(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol,
                    double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, double* scratch, cudaStream_t st)
{
    NNLM_REQUIRE(scd_tpc_supported(k), "the blocked SCD solver supports rank k <= 128");
    if (ncol <= 0) return;
    unsigned int* counter = reinterpret_cast<unsigned int*>(scratch);
    const int nh = (k + 3) / 4;                 // half-blocks of 4 coordinates (the K extent of one DMMA)
    if (nh > 16) {         // k in (64, 128]: the Gram alone takes up to 135 KB of shared memory -> 8-column tiles, 8 warps
        auto fn = nh <= 20 ? scd_chain::launch_big_a : nh <= 24 ? scd_chain::launch_big_b : nh <= 28 ? scd_chain::launch_big_c
                                                                                                  : scd_chain::launch_big_d;
        fn(nh, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
        return;
    }
    // Tile width, measured at 50000 x 10000, k = 50 with scd_chain (bench.py, ANLS iterations/s with 8- / 16- / 32-column
    // tiles on both halves: 371 / 458 / 418): 16 columns per warp keeps 12 warps per SM resident at 168 registers and
    // amortises the sequential part over twice the columns of the 8-column tile; 32 columns need 255 registers (8 warps)
    // and run two unbalanced rounds. On smaller shards the 8-column tile gives every scheduler more warps to overlap and wins:
    // measured per iteration (scratch/ct_threshold.py, W/H columns per GPU): 12500/2500 0.708 ms with 8-column tiles vs 0.765 with
    // 16-column tiles on the W side; 25000/5000 1.153 vs 1.054 ms; H side of 10000 columns 466 vs 458 iters/s for the 8-column
    // tile. The switch sits at 16 x 148 x 8 = 18944 columns. NNLM_SCD_CT / NNLM_SCD_CT2_MIN override (experiments).
    static const int force_ct = [] { const char* e = getenv("NNLM_SCD_CT"); return e ? atoi(e) : 0; }();
    static const int64_t ct2_min = [] { const char* e = getenv("NNLM_SCD_CT2_MIN"); return e ? atoll(e) : (int64_t)16 * 148 * 8; }();
    int ct = ncol >= ct2_min ? 2 : 1;
    if (force_ct == 1 || force_ct == 2 || force_ct == 4) ct = force_ct;
    auto fn = ct == 4 ? (nh <= 8 ? scd_chain::launch_ct4_lo : scd_chain::launch_ct4_hi)
            : ct == 2 ? (nh <= 8 ? scd_chain::launch_ct2_lo : scd_chain::launch_ct2_hi)
                      : (nh <= 8 ? scd_chain::launch_ct1_lo : scd_chain::launch_ct1_hi);
    fn(nh, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
}

}  // namespace nnlm
