"""Where does the T=1 H error at k=128 (FAST) come from? W-half error propagated vs H-half's own error."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, oracle, nnlm_b200
from nnlm_b200 import _capi as K
from conftest import umat
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
for (n, m, k) in [(5000, 2000, 128), (5000, 2000, 50)]:
    A = oracle.synth_matrix(n, m, k); At = np.asfortranarray(A.T)
    W0 = 0.01 * umat(11, n, k); H0 = 0.01 * umat(12, k, m)
    Wt0 = np.asfortranarray(W0.T)
    kw = dict(method=1, max_iter=50, rel_tol=1e-9)
    Wo, _ = oracle.update(Wt0, H0, At, n_threads=0, **kw)
    Ho, _ = oracle.update(H0, Wo, A, n_threads=0, **kw)
    for prec in (K.PREC_EXACT, K.PREC_FAST):
        Wg, _ = nnlm_b200.nnlm_update(Wt0, H0, At, precision=prec, **kw)
        Hg_from_Wo, _ = nnlm_b200.nnlm_update(H0, Wo, A, precision=prec, **kw)
        Hg_from_Wg, _ = nnlm_b200.nnlm_update(H0, Wg, A, precision=prec, **kw)
        Ho_from_Wg, _ = oracle.update(H0, Wg, A, n_threads=0, **kw)
        zo, zg = (Wo == 0), (Wg == 0)
        print(f"{n}x{m} k={k} prec={prec}: W rel {rel(Wg, Wo):.2e}; zero-pattern mismatches {int((zo != zg).sum())} of {zo.size} "
              f"(max |value| at mismatch {max(np.abs(Wg[zo != zg]).max() if (zo != zg).any() else 0, np.abs(Wo[zo != zg]).max() if (zo != zg).any() else 0):.2e}); "
              f"H-half alone (oracle W in) rel {rel(Hg_from_Wo, Ho):.2e}; GPU W -> GPU H rel {rel(Hg_from_Wg, Ho):.2e}; GPU W -> ORACLE H rel {rel(Ho_from_Wg, Ho):.2e}")
        Q, _ = nnlm_b200.cross(H0, At, precision=prec)
        Qr = H0 @ At
        print(f"     cross-product W-half: max err / sum|F||A| = {np.max(np.abs(Q - Qr) / (np.abs(H0) @ np.abs(At))):.2e}, rel Frobenius {rel(Q, Qr):.2e}")
        # error of W along / across the dominant direction
        d = Wg - Wo
        print(f"     W error: max abs {np.abs(d).max():.2e}, at entries where oracle W is zero: {np.abs(d[zo]).max():.2e}; col norms of oracle W min {np.linalg.norm(Wo,axis=1).min():.2e}")
