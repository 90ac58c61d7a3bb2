// scd_chain_big_a.cu — instantiations of the chain/DMMA SCD solver (scd_chain.cuh): 8-column tiles, 17..20 half-blocks (k 65..80)
#include "scd_chain.cuh"

namespace nnlm { namespace scd_chain {
void launch_big_a(int nh, NNLM_SCDC_ARGS)
{
    switch (nh) {
        case 17: launch<17, 1>(NNLM_SCDC_PASS); break;
        case 18: launch<18, 1>(NNLM_SCDC_PASS); break;
        case 19: launch<19, 1>(NNLM_SCDC_PASS); break;
        case 20: launch<20, 1>(NNLM_SCDC_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_chain: rank not in this instantiation set");
    }
}
} }
