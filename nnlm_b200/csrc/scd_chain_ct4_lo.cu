// scd_chain_ct4_lo.cu — instantiations of the chain/DMMA SCD solver (scd_chain.cuh): 32-column tiles, 1..8 half-blocks of 4 coordinates
#include "scd_chain.cuh"

namespace nnlm { namespace scd_chain {
void launch_ct4_lo(int nh, NNLM_SCDC_ARGS)
{
    switch (nh) {
        case 1: launch<1, 4>(NNLM_SCDC_PASS); break;
        case 2: launch<2, 4>(NNLM_SCDC_PASS); break;
        case 3: launch<3, 4>(NNLM_SCDC_PASS); break;
        case 4: launch<4, 4>(NNLM_SCDC_PASS); break;
        case 5: launch<5, 4>(NNLM_SCDC_PASS); break;
        case 6: launch<6, 4>(NNLM_SCDC_PASS); break;
        case 7: launch<7, 4>(NNLM_SCDC_PASS); break;
        case 8: launch<8, 4>(NNLM_SCDC_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_chain: rank not in this instantiation set");
    }
}
} }
