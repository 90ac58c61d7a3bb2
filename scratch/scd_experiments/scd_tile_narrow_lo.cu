// scd_tile_narrow_lo.cu — instantiations of the tiled SCD solver (scd_tile.cuh), T = 4 row groups, padded rank 4*kq4 for kq4 in {1 2 3 4 5 6 7 8 9 10}
#include "scd_tile.cuh"

namespace nnlm { namespace scd_tile {
void launch_narrow_lo(int kq4, NNLM_SCD_TILE_ARGS)
{
    switch (kq4) {
        case 1: launch<1, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 2: launch<2, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 3: launch<3, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 4: launch<4, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 5: launch<5, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 6: launch<6, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 7: launch<7, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 8: launch<8, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 9: launch<9, 4, 2>(NNLM_SCD_TILE_PASS); break;
        case 10: launch<10, 4, 2>(NNLM_SCD_TILE_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_tile: unsupported rank for this instantiation set");
    }
}
} }
