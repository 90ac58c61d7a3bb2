// solve_scd.cu — dispatch of the blocked DMMA SCD solver (scd_dmma.cuh) over tile widths and padded ranks.
#include <cstdlib>

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
    // Tile width. Measured at 50000 x 10000, k = 50 (ncu gpu__time_duration): 50000 columns take 1.26 / 1.22 / 1.67 ms with
    // 32- / 16- / 8-column tiles, 10000 columns 0.65 / 0.62 / 0.44 ms: the sequential part is latency-bound, so the narrowest
    // tile that still leaves every SM its 12 resident warps wins. NNLM_SCD_CT overrides (experiments).
    static const int force_ct = [] { const char* e = getenv("NNLM_SCD_CT"); return e ? atoi(e) : 0; }();
    int ct = ceil_div(ncol, 16) >= 148 * 12 ? 2 : 1;
    if (force_ct == 1 || force_ct == 2 || force_ct == 4) ct = force_ct;
    if (ct == 4) scd_dmma::launch_ct4(nb, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
    else if (ct == 2) scd_dmma::launch_ct2(nb, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
    else scd_dmma::launch_ct1(nb, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
}

}  // namespace nnlm
