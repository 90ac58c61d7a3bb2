import sys, os, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import nnlm_b200
from nnlm_b200 import _capi as K
from conftest import umat
k, n, m = 50, 50000, 1024
Wt = np.asfortranarray(umat(1, k, n)); A = np.asfortranarray(umat(2, n, k) @ umat(3, k, m) + 0.1 * umat(4, n, m))
ref = Wt @ A
Q, st = nnlm_b200.cross(Wt, A, precision=K.PREC_FAST)
d = (Q - ref) / ref
print("drain", os.environ.get("NNLM_TC_DRAIN"), "rel fro", np.linalg.norm(Q - ref) / np.linalg.norm(ref), "mean rel (bias)", d.mean(), "max |rel|", np.abs(d).max(), "std", d.std())
