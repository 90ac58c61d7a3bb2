"""The blocked SCD solver picks its tile width from the number of columns (nnlm_b200/csrc/solve_scd.cu): the 16-column tile
only runs from 16 x 148 x 8 = 18944 columns on, which none of the other parity tests reach. One half-iteration (src/update_with_missing.cpp:3-55
+ src/base_algorithms.cpp:3-37) on wide, short problems against the oracle, for ranks on both sides of every padding boundary
(k mod 8, k mod 4) and with coordinate masks."""
import numpy as np
import pytest

import nnlm_b200
import oracle
from nnlm_b200 import _capi as K
from conftest import umat

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def synth(n, m, k, noise=0.1, seed=0):
    return np.asfortranarray(umat(seed + 1, n, k) @ umat(seed + 2, k, m) + noise * umat(seed + 3, n, m))


@pytest.mark.parametrize("k", [3, 8, 9, 12, 13, 50, 60, 61, 64])
def test_update_wide_16_column_tiles(k):
    n, m = 100, 28500
    Wt = umat(1, k, n); A = synth(n, m, k, seed=31); H0 = umat(3, k, m)
    href, tref = oracle.update(H0, Wt, A, method=1, max_iter=30, rel_tol=1e-9, n_threads=0)
    hgot, tgot = nnlm_b200.nnlm_update(H0, Wt, A, method=1, max_iter=30, rel_tol=1e-9, precision=K.PREC_EXACT)
    assert rel(hgot, href) < 1e-9
    assert tgot == tref


def test_update_wide_tiles_masked_and_ragged():
    """28499 columns (a ragged last tile), coordinate masks, one fully masked column, L1 penalty."""
    n, m, k = 80, 28499, 21
    rng = np.random.default_rng(7)
    Wt = umat(1, k, n); A = synth(n, m, k, seed=32); H0 = umat(3, k, m)
    mask = rng.random((k, m)) < 0.15
    mask[:, 17] = True
    H0[mask] = 0
    beta = (0.01, 0.0, 0.02)
    href, tref = oracle.update(H0, Wt, A, method=1, max_iter=25, rel_tol=1e-9, n_threads=0, mask=mask, beta=beta)
    hgot, tgot = nnlm_b200.nnlm_update(H0, Wt, A, method=1, max_iter=25, rel_tol=1e-9, precision=K.PREC_EXACT, mask=mask, beta=beta)
    assert (hgot[mask] == 0).all()
    assert rel(hgot, href) < 1e-9
    assert tgot == tref


def test_update_wide_tiles_early_exit_counts():
    """Columns converge at different sweeps: the summed sweep count (total_raw_iter) must equal the oracle's."""
    n, m, k = 100, 30000, 10
    Wt = umat(1, k, n); A = synth(n, m, k, seed=33); H0 = umat(3, k, m)
    href, tref = oracle.update(H0, Wt, A, method=1, max_iter=200, rel_tol=1e-6, n_threads=0)
    hgot, tgot = nnlm_b200.nnlm_update(H0, Wt, A, method=1, max_iter=200, rel_tol=1e-6, precision=K.PREC_EXACT)
    assert rel(hgot, href) < 1e-8
    assert abs(tgot - tref) <= max(2, tref // 100000)
