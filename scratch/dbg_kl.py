import sys, warnings
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from test_gpu_parity import run_both, synth, rel
warnings.simplefilter("ignore")
rng = np.random.default_rng(5)
n, m, k = 150, 70, 5
A = synth(n, m, k)
Wm = rng.random((n, k)) < 0.2
Hm = rng.random((k, m)) < 0.1
Hm2 = Hm.copy(); Hm2[:, 7] = True
for name, kw in [("pen only", dict(alpha=(0.02, 0.01, 0.005), beta=(0.01, 0.0, 0.01))),
                 ("L2 only", dict(alpha=(0.02, 0, 0), beta=(0.01, 0.0, 0.0))),
                 ("angle only", dict(alpha=(0.0, 0.01, 0), beta=(0.0, 0.0, 0.0))),
                 ("L1 only", dict(alpha=(0.0, 0.0, 0.005), beta=(0.0, 0.0, 0.01))),
                 ("mask only", dict(Wm=Wm, Hm=Hm)),
                 ("mask+fullcol", dict(Wm=Wm, Hm=Hm2)),
                 ("Wm only", dict(Wm=Wm)), ("Hm only", dict(Hm=Hm))]:
    for method, inner in [(3, 2), (4, 2), (3, 1)]:
        for T in (1, 2, 6):
            ref, got = run_both(A, k, method, T, inner, **kw)
            print(f"{name:14s} method {method} inner {inner} T {T}: W {rel(got.W, ref['W']):.2e} H {rel(got.H, ref['H']):.2e}", flush=True)
