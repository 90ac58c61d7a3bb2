"""Config 5 at FULL size (200000 x 20000, k = 128, dense): T = 1 from the BASELINE init, oracle vs the fast GPU path on ONE GPU.
Needs ~70 GB of host memory for the oracle's A and A.t(): falls back to 100000 x 20000 when less than 120 GB is available."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from nnlm_b200.session import Session
from nnlm_b200 import _capi as K
avail = 0
for line in open("/proc/meminfo"):
    if line.startswith("MemAvailable"): avail = int(line.split()[1]) / 1e6
n, m, k = 200000, 20000, 128
if avail < 120: n = 100000
if avail < 70: print(f"only {avail:.0f} GB of host memory available: skipped"); sys.exit(0)
print(f"host memory available {avail:.0f} GB, cores {oracle.host_cores()}, problem {n} x {m}, k = {k}", flush=True)
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
W0 = 0.01 * oracle.splitmix_uniform(11, n * k).reshape((n, k), order="F")
H0 = 0.01 * oracle.splitmix_uniform(12, k * m).reshape((k, m), order="F")
s = Session(k=k, method=1, inner_max_iter=50, inner_rel_tol=1e-9, precision=K.PREC_FAST, device=0, synthetic=dict(n=n, m=m, na_frac=0.0))
s.set_factors(W0, H0); _, sw = s.run(1); Wg, Hg = s.get_factors(); s.close()
oracle.set_threads(oracle.host_cores())
t0 = time.perf_counter()
A = oracle.synth_matrix(n, m, k)
At = oracle.transpose(A)
kw = dict(n_threads=0, method=1, max_iter=50, rel_tol=1e-9, with_missing=0)
Wt, sw_w = oracle.update(np.asfortranarray(W0.T), H0.copy(order="F"), At, **kw)
del At
Ho, sw_h = oracle.update(H0.copy(order="F"), Wt, A, **kw)
print(f"oracle T=1 in {time.perf_counter() - t0:.1f} s", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "rounded":
    # the reference's own answer when A is rounded to fp32 (24 bits: what ANY 4-byte storage of A keeps)
    A = A.astype(np.float32).astype(np.float64, order="F")
    At = oracle.transpose(A)
    Wt2, _ = oracle.update(np.asfortranarray(W0.T), H0.copy(order="F"), At, **kw)
    del At
    Ho2, _ = oracle.update(H0.copy(order="F"), Wt2, A, **kw)
    print(f"oracle on fp32-rounded A vs oracle: rel W {rel(Wt2, Wt):.2e} rel H {rel(Ho2, Ho):.2e}", flush=True)
del A
cr = np.linalg.norm(Hg - Ho, axis=0) / np.maximum(np.linalg.norm(Ho, axis=0), 1e-300)
print(f"fast GPU vs oracle at {n} x {m}, k = {k}: rel W {rel(Wg, np.asfortranarray(Wt.T)):.2e} rel H {rel(Hg, Ho):.2e} | H columns: median {np.median(cr):.2e}, max {cr.max():.2e}, above 1e-5: {int((cr > 1e-5).sum())} of {m}; sweeps gpu {sw} oracle {int(sw_w) + int(sw_h)}; zero columns of W {int((np.abs(Wt).max(axis=1) == 0).sum())}", flush=True)
