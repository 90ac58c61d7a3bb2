"""Config 4 at FULL size (50000 x 10000, 20 % NA, k = 50): T = 1 from the BASELINE init — oracle (16 cores) vs the exact GPU path (twice: is it
deterministic?) vs the fast GPU path; per-column breakdown of the H differences."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from nnlm_b200.session import Session
from nnlm_b200 import _capi as K
n, m, k, na = 50000, 10000, 50, 0.2
if len(sys.argv) > 1: n, m = int(sys.argv[1]), int(sys.argv[2])
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
def colrel(a, b): return np.linalg.norm(a - b, axis=0) / np.maximum(np.linalg.norm(b, axis=0), 1e-300)
W0 = 0.01 * oracle.splitmix_uniform(11, n * k).reshape((n, k), order="F")
H0 = 0.01 * oracle.splitmix_uniform(12, k * m).reshape((k, m), order="F")
def gpu(prec):
    s = Session(k=k, method=1, inner_max_iter=50, inner_rel_tol=1e-9, precision=prec, device=0, synthetic=dict(n=n, m=m, na_frac=na))
    s.set_factors(W0, H0); _, sw = s.run(1); W, H = s.get_factors(); s.close(); return W, H, sw
We1, He1, sw1 = gpu(K.PREC_EXACT)
We2, He2, sw2 = gpu(K.PREC_EXACT)
Wf, Hf, swf = gpu(K.PREC_FAST)
print(f"exact run 1 vs run 2: rel W {rel(We2, We1):.2e} rel H {rel(He2, He1):.2e} sweeps {sw1} {sw2}; fast vs exact: rel W {rel(Wf, We1):.2e} rel H {rel(Hf, He1):.2e} sweeps {swf}", flush=True)
oracle.set_threads(oracle.host_cores())
t0 = time.perf_counter()
A = oracle.synth_matrix(n, m, k, na_frac=na)
At = oracle.transpose(A)
kw = dict(n_threads=0, method=1, max_iter=50, rel_tol=1e-9, with_missing=1)
Wt, sw_w = oracle.update(np.asfortranarray(W0.T), H0.copy(order="F"), At, **kw)
Ho, sw_h = oracle.update(H0.copy(order="F"), Wt, A, **kw)
Wo = np.asfortranarray(Wt.T)
print(f"oracle T=1 in {time.perf_counter() - t0:.1f} s on {oracle.host_cores()} cores", flush=True)
for name, W, H in (("exact GPU", We1, He1), ("fast GPU", Wf, Hf)):
    cr = colrel(H, Ho)
    print(f"{name} vs oracle: rel W {rel(W, Wo):.2e} rel H {rel(H, Ho):.2e} | H columns: median {np.median(cr):.2e}, 99 % {np.quantile(cr, 0.99):.2e}, max {cr.max():.2e}, above 1e-5: {int((cr > 1e-5).sum())}, above 1e-3: {int((cr > 1e-3).sum())}", flush=True)
