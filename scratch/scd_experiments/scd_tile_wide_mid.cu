// scd_tile_wide_mid.cu — instantiations of the tiled SCD solver (scd_tile.cuh), T = 2 row groups, padded rank 4*kq4 for kq4 in {9 10 11 12}
#include "scd_tile.cuh"

namespace nnlm { namespace scd_tile {
void launch_wide_mid(int kq4, NNLM_SCD_TILE_ARGS)
{
    switch (kq4) {
        case 9: launch<18, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 10: launch<20, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 11: launch<22, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 12: launch<24, 2, 2>(NNLM_SCD_TILE_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_tile: unsupported rank for this instantiation set");
    }
}
} }
