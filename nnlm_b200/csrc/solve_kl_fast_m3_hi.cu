// solve_kl_fast_m3_hi.cu — instantiations of the cluster KL solver (solve_kl_fast.cuh): method 3, 9..16 entries per thread
#include "solve_kl_fast.cuh"

namespace nnlm { namespace klf {
void launch_m3_hi(NNLM_KLF_ARGS)
{
    if (sh.E < 13) launch_range<3, 9>(NNLM_KLF_PASS);
    else launch_range<3, 13>(NNLM_KLF_PASS);
}
} }
