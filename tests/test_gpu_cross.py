"""The cross-product kernels in isolation (Q = Wt A, the contraction of src/update_with_missing.cpp:39,45):
fp64 CUDA-core path vs numpy to rounding, tcgen05 fp16-split path vs numpy to the 2^-22 representation bound."""
import numpy as np
import pytest

import nnlm_b200
from nnlm_b200 import _capi as K
from conftest import umat

pytestmark = pytest.mark.gpu

SHAPES = [(7, 100, 50), (50, 1000, 300), (50, 1003, 257), (64, 4099, 130), (33, 20000, 1000), (3, 64, 128), (1, 10, 3),
          (100, 3001, 200), (128, 2048, 256), (65, 777, 129)]


def inputs(k, n, m, signed=False):
    Wt = umat(1, k, n) * np.linspace(0.5, 40.0, k)[:, None]          # rows of very different magnitude
    A = umat(2, n, m) * 25.0
    if signed:
        Wt = Wt - 0.3 * Wt.max()
        A = A - 10.0
    return np.asfortranarray(Wt), np.asfortranarray(A)


@pytest.mark.parametrize("k,n,m", SHAPES)
def test_cross_exact_path(k, n, m):
    Wt, A = inputs(k, n, m)
    Q, st = nnlm_b200.cross(Wt, A, precision=K.PREC_EXACT)
    ref = Wt @ A
    assert np.max(np.abs(Q - ref) / np.abs(ref)) < 1e-13
    assert st["launches"] > 0 and st["precision_used"] == K.PREC_EXACT


@pytest.mark.parametrize("signed", [False, True])
@pytest.mark.parametrize("k,n,m", SHAPES)
def test_cross_tensor_core_path(k, n, m, signed):
    Wt, A = inputs(k, n, m, signed)
    Q, st = nnlm_b200.cross(Wt, A, precision=K.PREC_FAST)
    ref = Wt @ A
    assert st["precision_used"] == K.PREC_FAST
    # error relative to the magnitude of the terms summed (|Wt| |A|): operands carry 2^-22 worst-case relative error each
    bound = np.abs(Wt) @ np.abs(A)
    err = np.max(np.abs(Q - ref) / bound)
    assert err < 3e-7, err
    if not signed:
        assert np.linalg.norm(Q - ref) / np.linalg.norm(ref) < 1e-7


def test_cross_tensor_core_missing_entries_read_as_zero():
    Wt, A = inputs(20, 700, 90)
    A[umat(9, 700, 90) < 0.2] = np.nan
    Q, _ = nnlm_b200.cross(Wt, A, precision=K.PREC_EXACT)
    ref = Wt @ np.nan_to_num(A, nan=0.0)
    assert np.max(np.abs(Q - ref) / np.abs(ref)) < 1e-13
