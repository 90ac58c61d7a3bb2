import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import nnlm_b200
from nnlm_b200 import _capi as K
from conftest import umat
n = 600
for k in (8, 16, 24, 32, 40, 50, 64):
    for m in (148 * 4 * 32, 148 * 12 * 32):
        Wt = umat(1, k, n); A = np.asfortranarray(umat(2, n, k) @ umat(3, k, m)); H0 = umat(4, k, m)
        nnlm_b200.nnlm_update(H0, Wt, A, method=1, max_iter=50, rel_tol=-1, precision=K.PREC_EXACT)
