// solve_kl_fast_m4_hi.cu — instantiations of the cluster KL solver (solve_kl_fast.cuh): method 4, 9..16 entries per thread
#include "solve_kl_fast.cuh"

namespace nnlm { namespace klf {
void launch_m4_hi(NNLM_KLF_ARGS)
{
    if (sh.E < 13) launch_range<4, 9>(NNLM_KLF_PASS);
    else launch_range<4, 13>(NNLM_KLF_PASS);
}
} }
