// scd_chain_big_d.cu — instantiations of the chain/DMMA SCD solver (scd_chain.cuh): 8-column tiles, 29..32 half-blocks (k 113..128)
#include "scd_chain.cuh"

namespace nnlm { namespace scd_chain {
void launch_big_d(int nh, NNLM_SCDC_ARGS)
{
    switch (nh) {
        case 29: launch<29, 1>(NNLM_SCDC_PASS); break;
        case 30: launch<30, 1>(NNLM_SCDC_PASS); break;
        case 31: launch<31, 1>(NNLM_SCDC_PASS); break;
        case 32: launch<32, 1>(NNLM_SCDC_PASS); break;
        default: throw Error(NNLM_E_ARG, "scd_chain: rank not in this instantiation set");
    }
}
} }
