import sys, warnings
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from test_gpu_parity import run_both, synth, rel, oracle_sensitivity
warnings.simplefilter("ignore")
for k in (33, 50, 64, 100, 128):
    A = synth(400, 300, k)
    ref, got = run_both(A, k, 1, 1, 50)
    print("k", k, "T=1 W", rel(got.W, ref["W"]), "H", rel(got.H, ref["H"]), got.average_epochs, ref["average_epochs"], flush=True)
    ref, got = run_both(A, k, 1, 6, 50)
    print("   T=6 mse", got.mse, ref["mse"], flush=True)
rng = np.random.default_rng(5)
n, m, k = 150, 70, 5
A = synth(n, m, k)
Hm = rng.random((k, m)) < 0.1
print("sens", oracle_sensitivity(A, k, 3, 2, 2, Hm=Hm))
ref, got = run_both(A, k, 3, 2, 2, Hm=Hm)
print("err", rel(got.W, ref["W"]), rel(got.H, ref["H"]))
