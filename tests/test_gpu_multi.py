"""Multi-GPU sharded path (needs >= 2 GPUs on the box; skipped otherwise): launches tests/multi_gpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

from nnlm_b200 import _capi as K

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_matches_oracle_and_single_gpu():
    n_dev, _ = K.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n_dev < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29511", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_n_gpus_behind_the_single_nnmf_call():
    """nnlm_options.n_gpus / NNLM_B200_GPUS: ONE nnmf() call on one host thread sharded over N GPUs inside the library
    (worker threads + NCCL, SURVEY.md §8b "Threading") gives the single-GPU result and the oracle's."""
    import numpy as np
    import nnlm_b200
    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import umat
    n_dev, _ = K.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n_dev < 4 else 4
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    for (n, m, k, method, prec, na) in [(1003, 517, 6, 1, K.PREC_EXACT, 0.0), (2003, 1017, 6, 1, K.PREC_FAST, 0.0),
                                        (700, 300, 4, 4, K.PREC_EXACT, 0.0), (700, 300, 4, 1, K.PREC_EXACT, 0.15)]:
        A = oracle.synth_matrix(n, m, k, na_frac=na)
        W0 = 0.01 * umat(11, n, k); H0 = 0.01 * umat(12, k, m)
        inner = 50 if method < 3 else 1
        kw = dict(init={"W": W0, "H": H0}, method="scd" if method in (1, 3) else "lee", loss="mse" if method < 3 else "mkl",
                  max_iter=6, rel_tol=-1, trace=2, inner_max_iter=inner, show_warning=False, check_k=False, precision=prec)
        one = nnlm_b200.nnmf(A, k, n_gpus=1, **kw)
        many = nnlm_b200.nnmf(A, k, n_gpus=world, **kw)
        ref = oracle.nnmf(A, k, W0, H0, max_iter=6, rel_tol=-1, n_threads=0, inner_max_iter=inner, method=method, trace=2)
        assert many.stats["n_gpus_used"] == world and one.stats["n_gpus_used"] == 1
        assert rel(many.W, one.W) < 1e-6 and rel(many.H, one.H) < 1e-6
        assert rel(many.W, ref["W"]) < 1e-5 and rel(many.H, ref["H"]) < 1e-5
        np.testing.assert_allclose(many.mse, ref["mse"], rtol=1e-6)
        np.testing.assert_allclose(many.average_epochs, one.average_epochs, rtol=1e-12)
    # interrupt: rank 0 polls, every rank stops at the same iteration
    calls = {"n": 0}
    def stop():
        calls["n"] += 1
        return calls["n"] > 3
    with pytest.raises(K.Interrupted):
        nnlm_b200.nnmf(A, k, n_gpus=world, interrupt=stop, **kw)
