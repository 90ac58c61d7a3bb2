// scd_tile_wide_hi.cu — instantiations of the tiled SCD solver (scd_tile.cuh), T = 2 row groups, padded rank 4*kq4 for kq4 in {13 14 15 16}
#include "scd_tile.cuh"

namespace nnlm { namespace scd_tile {
void launch_wide_hi(int kq4, NNLM_SCD_TILE_ARGS)
{
    switch (kq4) {
        case 13: launch<26, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 14: launch<28, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 15: launch<30, 2, 2>(NNLM_SCD_TILE_PASS); break;
        case 16: launch<32, 2, 2>(NNLM_SCD_TILE_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_tile: unsupported rank for this instantiation set");
    }
}
} }
