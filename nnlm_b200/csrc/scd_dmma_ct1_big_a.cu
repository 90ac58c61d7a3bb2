// scd_dmma_ct1_big_a.cu — instantiations of the blocked DMMA SCD solver (scd_dmma.cuh), 8-column tiles, padded rank 8*nb, nb 9..12
#include "scd_dmma.cuh"

namespace nnlm { namespace scd_dmma {
void launch_ct1_big_a(int nb, NNLM_SCD_ARGS)
{
    switch (nb) {
        case 9: launch<9, 1>(NNLM_SCD_PASS); break;
        case 10: launch<10, 1>(NNLM_SCD_PASS); break;
        case 11: launch<11, 1>(NNLM_SCD_PASS); break;
        case 12: launch<12, 1>(NNLM_SCD_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_dmma: rank not in this instantiation set");
    }
}
} }
