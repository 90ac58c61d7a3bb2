"""Small problems through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, nnlm_b200, oracle
from nnlm_b200 import _capi as K
from conftest import umat
def run(n, m, k, method, loss, na=0.0, prec=K.PREC_FAST, T=2, inner=None):
    A = oracle.synth_matrix(n, m, min(k, 8), na_frac=na)
    W0 = 0.01 * umat(11, n, k); H0 = 0.01 * umat(12, k, m)
    r = nnlm_b200.nnmf(A, k, method=method, loss=loss, init={"W": W0, "H": H0}, max_iter=T, rel_tol=-1, trace=1, show_warning=False,
                       check_k=False, precision=prec, inner_max_iter=inner)
    print(n, m, k, method, loss, na, prec, "mse", r.mse[-1], flush=True)
run(777, 333, 7, "scd", "mse")                      # cross_tc<32>, scd_chain, error_tc, factor_prep
run(900, 400, 50, "scd", "mse")                     # cross_tc<64>
run(700, 300, 100, "scd", "mse", T=1)               # cross_tc<128>, two k-blocks in error_tc
run(1300, 500, 9, "lee", "mkl")                     # solve_kl_fast (S=1..4)
run(5000, 300, 5, "scd", "mkl", inner=2)            # solve_kl_fast clusters of 8 on the H-half
run(800, 350, 6, "scd", "mse", na=0.2)              # mask planes, z slices, exact contraction, packed solve
run(600, 250, 6, "lee", "mse", prec=K.PREC_EXACT)   # fp64 path
run(600, 250, 5, "scd", "mse", na=0.1, prec=K.PREC_EXACT)
