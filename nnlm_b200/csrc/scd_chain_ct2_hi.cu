// scd_chain_ct2_hi.cu — instantiations of the chain/DMMA SCD solver (scd_chain.cuh): 16-column tiles, 9..16 half-blocks of 4 coordinates
#include "scd_chain.cuh"

namespace nnlm { namespace scd_chain {
void launch_ct2_hi(int nh, NNLM_SCDC_ARGS)
{
    switch (nh) {
        case 9: launch<9, 2>(NNLM_SCDC_PASS); break;
        case 10: launch<10, 2>(NNLM_SCDC_PASS); break;
        case 11: launch<11, 2>(NNLM_SCDC_PASS); break;
        case 12: launch<12, 2>(NNLM_SCDC_PASS); break;
        case 13: launch<13, 2>(NNLM_SCDC_PASS); break;
        case 14: launch<14, 2>(NNLM_SCDC_PASS); break;
        case 15: launch<15, 2>(NNLM_SCDC_PASS); break;
        case 16: launch<16, 2>(NNLM_SCDC_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_chain: rank not in this instantiation set");
    }
}
} }
