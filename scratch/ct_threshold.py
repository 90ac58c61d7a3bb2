# which tile width wins on a half-size shard (the N = 2 regime: 25000 x 5000 per GPU)? run with NNLM_SCD_CT2_MIN set
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnlm_b200.session import Session, synth_init
for (n, m) in ((25000, 5000), (12500, 2500), (36000, 7200)):
    k = 50
    W0, H0 = synth_init(n, m, k)
    s = Session(k=k, method=1, inner_max_iter=50, inner_rel_tol=1e-9, precision=2, device=0, synthetic=dict(n=n, m=m, na_frac=0.0), timing=True)
    s.set_factors(W0, H0)
    s.run(5); s.reset_stats()
    ms, _ = s.run(20)
    st = s.stats()
    print(os.environ.get("NNLM_SCD_CT2_MIN"), n, m, "ms/iter", round(ms / 20, 4), "solve", round(st["solve_ms"] / 20, 4))
    s.close()
