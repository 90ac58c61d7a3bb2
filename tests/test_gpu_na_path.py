"""The NA path of the square loss (src/update_with_missing.cpp:58-117) on the GPU: per-column Gram as fp64 tensor-core tiles
(complement or direct), then one warp per column in the batched solver (nnlm_b200/csrc/solve_ls_missing.cu). One
half-iteration against the oracle for every register layout of the warp solver (k <= 32, 64, 96, 128), with coordinate
masks, regularisation, mostly-empty columns, and with the per-column Gram scratch forced into several chunks."""
import os
import subprocess
import sys

import numpy as np
import pytest

import nnlm_b200
import oracle
from nnlm_b200 import _capi as K
from conftest import umat

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def synth_na(n, m, k, na, seed=0):
    A = umat(seed + 1, n, k) @ umat(seed + 2, k, m) + 0.1 * umat(seed + 3, n, m)
    A[umat(seed + 4, n, m) < na] = np.nan
    return np.asfortranarray(A)


@pytest.mark.parametrize("method", [1, 2])
@pytest.mark.parametrize("k", [7, 33, 50, 64, 70, 100, 128])
def test_update_missing_every_rank_layout(method, k):
    n, m = 400, 150
    Wt = umat(1, k, n); A = synth_na(n, m, k, 0.2, seed=40); H0 = umat(3, k, m)
    href, tref = oracle.update(H0, Wt, A, method=method, max_iter=15, rel_tol=1e-9, n_threads=0)
    hgot, tgot = nnlm_b200.nnlm_update(H0, Wt, A, method=method, max_iter=15, rel_tol=1e-9, precision=K.PREC_EXACT)
    assert rel(hgot, href) < 1e-9
    assert tgot == tref


def test_update_missing_masked_regularised_direct_branch():
    """70 % missing (the Gram is summed over the present rows), coordinate masks, one fully masked column, one complete
    column, all three penalties."""
    n, m, k = 300, 120, 12
    rng = np.random.default_rng(11)
    Wt = umat(1, k, n); A = synth_na(n, m, k, 0.7, seed=41); H0 = umat(3, k, m)
    A[:, 5] = np.asfortranarray(umat(42, n, k) @ umat(43, k, m))[:, 5]
    mask = rng.random((k, m)) < 0.15
    mask[:, 9] = True
    H0[mask] = 0
    beta = (0.02, 0.01, 0.005)
    href, tref = oracle.update(H0, Wt, A, method=1, max_iter=30, rel_tol=1e-9, n_threads=0, mask=mask, beta=beta)
    hgot, tgot = nnlm_b200.nnlm_update(H0, Wt, A, method=1, max_iter=30, rel_tol=1e-9, precision=K.PREC_EXACT, mask=mask, beta=beta)
    assert (hgot[mask] == 0).all()
    assert rel(hgot, href) < 1e-9
    assert tgot == tref


_CHUNK_SCRIPT = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import numpy as np
import nnlm_b200, oracle
from nnlm_b200 import _capi as K
from conftest import umat
n, m, k = 200, 333, 20
A = umat(1, n, k) @ umat(2, k, m) + 0.1 * umat(3, n, m)
A[umat(4, n, m) < 0.25] = np.nan
A = np.asfortranarray(A)
Wt = umat(5, k, n); H0 = umat(6, k, m)
href, tref = oracle.update(H0, Wt, A, method=1, max_iter=20, rel_tol=1e-9, n_threads=0)
hgot, tgot = nnlm_b200.nnlm_update(H0, Wt, A, method=1, max_iter=20, rel_tol=1e-9, precision=K.PREC_EXACT)
err = np.linalg.norm(hgot - href) / np.linalg.norm(href)
assert err < 1e-9 and tgot == tref, (err, tgot, tref)
print("ok", err)
"""


def test_update_missing_scratch_in_several_chunks():
    """NNLM_NA_SCRATCH_DOUBLES = 40 Grams of 20 x 20: the 333 columns go through the Gram scratch in 9 chunks."""
    env = dict(os.environ, NNLM_NA_SCRATCH_DOUBLES=str(40 * 20 * 20))
    out = subprocess.run([sys.executable, "-c", _CHUNK_SCRIPT.format(root=ROOT)], env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok")
