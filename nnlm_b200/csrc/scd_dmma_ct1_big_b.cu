// scd_dmma_ct1_big_b.cu — instantiations of the blocked DMMA SCD solver (scd_dmma.cuh), 8-column tiles, padded rank 8*nb, nb 13..16
#include "scd_dmma.cuh"

namespace nnlm { namespace scd_dmma {
void launch_ct1_big_b(int nb, NNLM_SCD_ARGS)
{
    switch (nb) {
        case 13: launch<13, 1>(NNLM_SCD_PASS); break;
        case 14: launch<14, 1>(NNLM_SCD_PASS); break;
        case 15: launch<15, 1>(NNLM_SCD_PASS); break;
        case 16: launch<16, 1>(NNLM_SCD_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_dmma: rank not in this instantiation set");
    }
}
} }
