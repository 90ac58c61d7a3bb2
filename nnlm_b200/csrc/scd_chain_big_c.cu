// scd_chain_big_c.cu — instantiations of the chain/DMMA SCD solver (scd_chain.cuh): 8-column tiles, 25..28 half-blocks (k 97..112)
#include "scd_chain.cuh"

namespace nnlm { namespace scd_chain {
void launch_big_c(int nh, NNLM_SCDC_ARGS)
{
    switch (nh) {
        case 25: launch<25, 1>(NNLM_SCDC_PASS); break;
        case 26: launch<26, 1>(NNLM_SCDC_PASS); break;
        case 27: launch<27, 1>(NNLM_SCDC_PASS); break;
        case 28: launch<28, 1>(NNLM_SCDC_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_chain: rank not in this instantiation set");
    }
}
} }
