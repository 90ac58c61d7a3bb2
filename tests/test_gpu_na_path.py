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


# ---- the NA path on the tensor cores (fast precision): na_gram.cu ------------------------------------------------------
def _pairs(k):
    a, b = np.tril_indices(k)
    return a, b


@pytest.mark.parametrize("n,m,k,scale", [(700, 90, 5, 1.0), (3000, 300, 50, 1.0), (2000, 150, 50, 1e6), (1500, 140, 128, 1.0)])
def test_tensor_core_mask_contraction_is_exact_to_the_slicing(n, m, k, scale):
    """S_j = sum_{i missing} y_i y_i' from the 0/1-mask x fixed-point-slice contraction vs numpy fp64. The products and the
    fp32 partial sums are exact; what is left is the 2^-44 truncation of every product against its column bound
    max|y_a| max|y_b| ("show the Gram's relative error vs fp64 and why the planes suffice")."""
    rng = np.random.default_rng(4)
    Wt = umat(1, k, n) * scale
    Wt[rng.random((k, n)) < 0.3] = 0.0                       # NNLS factors are sparse
    Wt[:, ::7] *= 37.0                                       # and not uniformly scaled
    if k > 3:
        Wt[3, :] = 0.0                                       # a vanished factor column (mid-trajectory)
    A = oracle.synth_matrix(n, m, min(k, 8), na_frac=0.2)
    miss = ~np.isfinite(A)
    S_pairs, S_rows = nnlm_b200.api.na_corrections(Wt, A)
    a, b = _pairs(k)
    Z = (Wt[a, :] * Wt[b, :])                                 # (pairs, n)
    ref = (Z @ miss.astype(np.float64)).T                     # (m, pairs)
    bound = (np.abs(Wt).max(axis=1)[a] * np.abs(Wt).max(axis=1)[b])[None, :] * miss.sum(axis=0)[:, None]
    err = np.abs(S_pairs - ref)
    worst = float((err / np.maximum(bound, 1e-300)).max())
    relf = float(np.linalg.norm(S_pairs - ref) / np.linalg.norm(ref))
    print(f"{n}x{m} k={k}: |S - S_fp64| / (n_missing max|y_a| max|y_b|) <= {worst:.2e}; relative Frobenius {relf:.2e}")
    assert worst < 2.0 ** -42 and relf < 1e-11
    if k > 3:
        assert (S_pairs[:, (a == 3) | (b == 3)] == 0).all()  # exact zeros stay exact
    ref_rows = (Wt @ miss.astype(np.float64)).T
    assert np.linalg.norm(S_rows - ref_rows) / np.linalg.norm(ref_rows) < 1e-11


@pytest.mark.parametrize("method", [1, 2])
@pytest.mark.parametrize("k", [5, 50, 100])
def test_update_missing_fast_path_parity(method, k):
    n, m = 3000, 260
    Wt = umat(1, k, n); A = oracle.synth_matrix(n, m, min(k, 20), na_frac=0.2, seed_base=30); H0 = umat(3, k, m)
    A[:, 5] = np.nan; A[::2, 7] = np.nan; A[:10, 9] = np.nan          # a wholly missing column, a half-missing one
    A[:, 11] = np.where(np.isnan(A[:, 11]), 1.0, A[:, 11])            # and a complete one
    href, tref = oracle.update(H0, Wt, A, method=method, max_iter=30, rel_tol=1e-9, n_threads=0, with_missing=1)
    hgot, tgot = nnlm_b200.nnlm_update(H0, Wt, A, method=method, max_iter=30, rel_tol=1e-9, with_missing=1, precision=K.PREC_FAST)
    e = np.linalg.norm(hgot - href) / np.linalg.norm(href)
    print(f"NA fast path method {method} k={k}: rel {e:.2e}, sweeps {tgot} / {tref}")
    assert e < 1e-5
    assert abs(tgot - tref) <= 1e-2 * tref        # a few columns leave the sweep loop one test earlier or later (1e-9 exit test)


def test_nnmf_missing_fast_path_with_masks_and_penalties():
    n, m, k = 1200, 500, 6
    A = oracle.synth_matrix(n, m, k, na_frac=0.15)
    rng = np.random.default_rng(8)
    Wm = rng.random((n, k)) < 0.05; Hm = rng.random((k, m)) < 0.05
    W0 = 0.01 * umat(11, n, k); H0 = 0.01 * umat(12, k, m)
    W0[Wm] = 0; H0[Hm] = 0
    alpha, beta = (0.05, 0.01, 0.02), (0.03, 0.0, 0.01)
    ref = oracle.nnmf(A, k, W0, H0, Wm=Wm, Hm=Hm, alpha=alpha, beta=beta, max_iter=5, rel_tol=-1, n_threads=0, inner_max_iter=50,
                      method=1, trace=1)
    got = nnlm_b200.nnmf(A, k, alpha=alpha, beta=beta, init={"W": W0, "H": H0}, mask={"W": Wm, "H": Hm}, max_iter=5, rel_tol=-1,
                         trace=1, show_warning=False, check_k=False, precision=K.PREC_FAST)
    assert got.stats["precision_used"] == K.PREC_FAST
    ew = np.linalg.norm(got.W - ref["W"]) / np.linalg.norm(ref["W"]); eh = np.linalg.norm(got.H - ref["H"]) / np.linalg.norm(ref["H"])
    print(f"NA nnmf fast: rel W {ew:.2e}, rel H {eh:.2e}")
    assert ew < 1e-5 and eh < 1e-5
    assert (got.W[Wm] == 0).all() and (got.H[Hm] == 0).all()
    np.testing.assert_allclose(got.mse, ref["mse"], rtol=1e-6)
    np.testing.assert_allclose(got.target_loss, ref["target_loss"], rtol=1e-6)
