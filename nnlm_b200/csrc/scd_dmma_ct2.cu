// scd_dmma_ct2.cu — instantiations of the blocked DMMA SCD solver (scd_dmma.cuh) for 16-column tiles, padded rank 8*nb
#include "scd_dmma.cuh"

namespace nnlm { namespace scd_dmma {
void launch_ct2(int nb, NNLM_SCD_ARGS)
{
    switch (nb) {
        case 1: launch<1, 2>(NNLM_SCD_PASS); break;
        case 2: launch<2, 2>(NNLM_SCD_PASS); break;
        case 3: launch<3, 2>(NNLM_SCD_PASS); break;
        case 4: launch<4, 2>(NNLM_SCD_PASS); break;
        case 5: launch<5, 2>(NNLM_SCD_PASS); break;
        case 6: launch<6, 2>(NNLM_SCD_PASS); break;
        case 7: launch<7, 2>(NNLM_SCD_PASS); break;
        case 8: launch<8, 2>(NNLM_SCD_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_dmma: rank k > 64 is not instantiated");
    }
}
} }
