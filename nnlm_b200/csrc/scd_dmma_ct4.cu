// scd_dmma_ct4.cu — instantiations of the blocked DMMA SCD solver (scd_dmma.cuh) for 32-column tiles, padded rank 8*nb
#include "scd_dmma.cuh"

namespace nnlm { namespace scd_dmma {
void launch_ct4(int nb, NNLM_SCD_ARGS)
{
    switch (nb) {
        case 1: launch<1, 4>(NNLM_SCD_PASS); break;
        case 2: launch<2, 4>(NNLM_SCD_PASS); break;
        case 3: launch<3, 4>(NNLM_SCD_PASS); break;
        case 4: launch<4, 4>(NNLM_SCD_PASS); break;
        case 5: launch<5, 4>(NNLM_SCD_PASS); break;
        case 6: launch<6, 4>(NNLM_SCD_PASS); break;
        case 7: launch<7, 4>(NNLM_SCD_PASS); break;
        case 8: launch<8, 4>(NNLM_SCD_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_dmma: rank k > 64 is not instantiated");
    }
}
} }
