import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracle, nnlm_b200
from nnlm_b200 import _capi as K
from conftest import umat
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
n, m, k = 5000, 2000, int(os.environ.get("KK", "128"))
A = oracle.synth_matrix(n, m, k); At = np.asfortranarray(A.T)
W0 = 0.01 * umat(11, n, k); H0 = 0.01 * umat(12, k, m)
Wt0 = np.asfortranarray(W0.T)
kw = dict(method=1, max_iter=50, rel_tol=1e-9)
Wo, _ = oracle.update(Wt0, H0, At, n_threads=0, **kw)
Ho, _ = oracle.update(H0, Wo, A, n_threads=0, **kw)
Wg, _ = nnlm_b200.nnlm_update(Wt0, H0, At, precision=K.PREC_FAST, **kw)
Hg, _ = nnlm_b200.nnlm_update(H0, Wg, A, precision=K.PREC_FAST, **kw)
Q, _ = nnlm_b200.cross(H0, At, precision=K.PREC_FAST)
Qr = H0 @ At
e = Q - Qr
# component of the error along the all-ones direction in the k-space vs orthogonal
along = e.mean(axis=0, keepdims=True) * np.ones((k, 1))
print(f"drain={os.environ.get('NNLM_TC_DRAIN','default')} k={k}: Q max err/sum|F||A| {np.max(np.abs(e) / (np.abs(H0) @ np.abs(At))):.2e} rel {rel(Q, Qr):.2e} "
      f"(orthogonal-to-ones part {np.linalg.norm(e - along) / np.linalg.norm(Qr):.2e}); W rel {rel(Wg, Wo):.2e}; H rel {rel(Hg, Ho):.2e}")
