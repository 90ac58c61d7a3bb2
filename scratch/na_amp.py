"""Config-4-like problems: T = 1 on the fast path vs the exact (fp64) path of the same library, and the exact path's own response to a
1e-12 perturbation of H0 (the conditioning of the first iteration from the tiny BASELINE init)."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnlm_b200.session import Session
from nnlm_b200 import _capi as K
from nnlm_b200.session import synth_init
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
def t1(n, m, k, prec, W0, H0, na):
    s = Session(k=k, method=1, inner_max_iter=50, inner_rel_tol=1e-9, precision=prec, device=0, synthetic=dict(n=n, m=m, na_frac=na))
    s.set_factors(W0, H0); s.run(1); W, H = s.get_factors(); s.close(); return W, H
for na in (0.2, 0.0):
    for (n, m, k) in ((5000, 2000, 50), (20000, 5000, 50), (50000, 10000, 50)):
        W0, H0 = synth_init(n, m, k)
        Wf, Hf = t1(n, m, k, K.PREC_FAST, W0, H0, na)
        We, He = t1(n, m, k, K.PREC_EXACT, W0, H0, na)
        rng = np.random.default_rng(1)
        Hp = H0 * (1.0 + 1e-12 * rng.standard_normal(H0.shape))
        Wq, Hq = t1(n, m, k, K.PREC_EXACT, W0, Hp, na)
        zc = int((np.abs(We).max(axis=0) == 0).sum())
        print(f"{n}x{m} k={k} NA {na}: T=1 fast vs exact rel W {rel(Wf, We):.2e} rel H {rel(Hf, He):.2e} (x{rel(Hf, He) / max(rel(Wf, We), 1e-300):.0f}) | exact path, H0 perturbed 1e-12: rel W {rel(Wq, We):.2e} rel H {rel(Hq, He):.2e} (x{rel(Hq, He) / max(rel(Wq, We), 1e-300):.0f}) | zero columns of W {zc}", flush=True)
