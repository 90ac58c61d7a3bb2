// scd_chain_ct4_hi.cu — instantiations of the chain/DMMA SCD solver (scd_chain.cuh): 32-column tiles, 9..16 half-blocks of 4 coordinates
#include "scd_chain.cuh"

namespace nnlm { namespace scd_chain {
void launch_ct4_hi(int nh, NNLM_SCDC_ARGS)
{
    switch (nh) {
        case 9: launch<9, 4>(NNLM_SCDC_PASS); break;
        case 10: launch<10, 4>(NNLM_SCDC_PASS); break;
        case 11: launch<11, 4>(NNLM_SCDC_PASS); break;
        case 12: launch<12, 4>(NNLM_SCDC_PASS); break;
        case 13: launch<13, 4>(NNLM_SCDC_PASS); break;
        case 14: launch<14, 4>(NNLM_SCDC_PASS); break;
        case 15: launch<15, 4>(NNLM_SCDC_PASS); break;
        case 16: launch<16, 4>(NNLM_SCDC_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_chain: rank not in this instantiation set");
    }
}
} }
