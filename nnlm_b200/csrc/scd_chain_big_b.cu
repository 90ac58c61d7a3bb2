// scd_chain_big_b.cu — instantiations of the chain/DMMA SCD solver (scd_chain.cuh): 8-column tiles, 21..24 half-blocks (k 81..96)
#include "scd_chain.cuh"

namespace nnlm { namespace scd_chain {
void launch_big_b(int nh, NNLM_SCDC_ARGS)
{
    switch (nh) {
        case 21: launch<21, 1>(NNLM_SCDC_PASS); break;
        case 22: launch<22, 1>(NNLM_SCDC_PASS); break;
        case 23: launch<23, 1>(NNLM_SCDC_PASS); break;
        case 24: launch<24, 1>(NNLM_SCDC_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_chain: rank not in this instantiation set");
    }
}
} }
