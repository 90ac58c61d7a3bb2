// solve_scd_tile.cu — dispatch of the tiled SCD solver (scd_tile.cuh) over its instantiation sets.
#include "scd_tile.cuh"

namespace nnlm {

bool scd_tpc_supported(int k) { return k >= 1 && k <= 64; }
size_t scd_tpc_scratch_doubles() { return 2; }

void launch_scd_tpc(double* X, const double* G, const double* Qp, int splits, const uint8_t* mask, int k, int64_t ncol,
                    double l1, unsigned max_iter, double rel_tol, unsigned long long* sweeps, double* scratch, cudaStream_t st)
{
    NNLM_REQUIRE(scd_tpc_supported(k), "tiled SCD supports rank k <= 64");
    if (ncol <= 0) return;
    unsigned int* counter = reinterpret_cast<unsigned int*>(scratch);
    // 32-column tiles (2 row groups) when there are enough columns to give every scheduler two or three warps,
    // else 16-column tiles (4 row groups): twice the warps, half the work per step
    const bool wide = ceil_div(ncol, 32) >= 148 * 8;
    const int kq4 = (k + 3) / 4;
    if (wide) {
        if (kq4 <= 8) scd_tile::launch_wide_lo(kq4, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
        else if (kq4 <= 12) scd_tile::launch_wide_mid(kq4, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
        else scd_tile::launch_wide_hi(kq4, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
    } else {
        if (kq4 <= 10) scd_tile::launch_narrow_lo(kq4, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
        else scd_tile::launch_narrow_hi(kq4, X, G, Qp, splits, mask, k, ncol, l1, max_iter, rel_tol, sweeps, counter, st);
    }
}

}  // namespace nnlm
