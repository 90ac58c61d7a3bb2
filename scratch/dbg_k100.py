import sys, warnings
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from test_gpu_parity import run_both, synth, rel
warnings.simplefilter("ignore")
for (n,m,k) in [(400,300,100),(400,300,64),(400,300,33),(2000,1500,100),(400,300,128)]:
    A = synth(n,m,k)
    for T in (1,2,3):
        ref, got = run_both(A, k, 1, T, 50)
        print(n,m,k,T, rel(got.W, ref["W"]), rel(got.H, ref["H"]), got.average_epochs, ref["average_epochs"], flush=True)
