// scd_tile_wide_lo.cu — instantiations of the tiled SCD solver (scd_tile.cuh), T = 2 row groups, padded rank 4*kq4 for kq4 in {1 2 3 4 5 6 7 8}
#include "scd_tile.cuh"

namespace nnlm { namespace scd_tile {
void launch_wide_lo(int kq4, NNLM_SCD_TILE_ARGS)
{
    switch (kq4) {
        case 1: launch<2, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 2: launch<4, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 3: launch<6, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 4: launch<8, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 5: launch<10, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 6: launch<12, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 7: launch<14, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 8: launch<16, 2, 2>(NNLM_SCD_TILE_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_tile: unsupported rank for this instantiation set");
    }
}
} }
