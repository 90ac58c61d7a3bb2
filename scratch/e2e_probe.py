"""Why does the e2e wall clock of nnmf() vary 2x between calls? Per-call host timers (alloc / set-up / loop / teardown)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, nnlm_b200
from nnlm_b200.session import Session, synth_matrix, synth_init
n, m, k = 50000, 10000, 50
W0, H0 = synth_init(n, m, k)
s = Session(k=k, method=1, precision=2, synthetic=dict(n=n, m=m), timing=True)
s.set_factors(W0, H0); s.run(5); s.close()
pin = lambda a: np.asfortranarray(torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory().numpy().T)
Ap = pin(synth_matrix(n, m, k)); Wp, Hp = pin(W0), pin(H0)
for i in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = nnlm_b200.nnmf(Ap, k, init={"W": Wp, "H": Hp}, max_iter=20, rel_tol=-1, trace=0, show_warning=False, check_k=False, precision=2)
    w = time.perf_counter() - t0
    st = r.stats
    print(f"call {i}: wall {w*1e3:7.1f} ms | in C: total {st['host_total_ms']:7.1f} setup {st['host_setup_ms']:6.1f} (alloc {st['host_alloc_ms']:6.1f}, upload ev {st['upload_ms']:5.1f}) "
          f"loop {st['host_loop_ms']:5.1f} finish {st['host_finish_ms']:4.1f} teardown {st['host_teardown_ms']:6.1f}", flush=True)
