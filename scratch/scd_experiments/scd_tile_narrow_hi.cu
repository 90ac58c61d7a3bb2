// scd_tile_narrow_hi.cu — instantiations of the tiled SCD solver (scd_tile.cuh), T = 4 row groups, padded rank 4*kq4 for kq4 in {11 12 13 14 15 16}
#include "scd_tile.cuh"

namespace nnlm { namespace scd_tile {
void launch_narrow_hi(int kq4, NNLM_SCD_TILE_ARGS)
{
    switch (kq4) {
        case 11: launch<11, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 12: launch<12, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 13: launch<13, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 14: launch<14, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 15: launch<15, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 16: launch<16, 4, 2>(NNLM_SCD_TILE_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_tile: unsupported rank for this instantiation set");
    }
}
} }
