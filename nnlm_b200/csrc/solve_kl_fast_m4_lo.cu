// solve_kl_fast_m4_lo.cu — instantiations of the cluster KL solver (solve_kl_fast.cuh): method 4, 1..8 entries per thread
#include "solve_kl_fast.cuh"

namespace nnlm { namespace klf {
void launch_m4_lo(NNLM_KLF_ARGS)
{
    if (sh.E < 5) launch_range<4, 1>(NNLM_KLF_PASS);
    else launch_range<4, 5>(NNLM_KLF_PASS);
}
} }
