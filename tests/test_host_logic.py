"""Host-side mirror of the R front-end (no GPU): method codes (R/misc.R:28-35), reformat.input (R/misc.R:48-129),
mse.mkl (R/misc.R:9-16), argument errors of nnmf()/nnlm() that fire before the .Call."""
import numpy as np
import pytest

import nnlm_b200
from nnlm_b200.api import get_method_code, mse_mkl, reformat_input


def test_method_codes():
    assert get_method_code("scd", "mse") == 1
    assert get_method_code("lee", "mse") == 2
    assert get_method_code("scd", "mkl") == 3
    assert get_method_code("lee", "mkl") == 4
    with pytest.raises(ValueError):
        get_method_code("foo", "mse")


def test_reformat_input_defaults():
    r = reformat_input(None, None, 7, 5, 3)
    assert r["K"] == 3 and r["kW0"] == 0 and r["kH0"] == 0
    assert r["Wi"] is None and r["Hi"] is None and r["Wm"] is None and r["Hm"] is None


def test_reformat_input_known_profiles():
    n, m, k = 7, 5, 2
    W0 = np.ones((n, 2)); H0 = np.ones((1, m))
    r = reformat_input({"W0": W0, "H0": H0}, None, n, m, k, np.random.default_rng(0))
    assert r["K"] == 5 and r["kW0"] == 2 and r["kH0"] == 1
    # Wi = [W | W0 | W1], Hi = [H ; H1 ; H0]  (R/misc.R:120-123)
    assert r["Wi"].shape == (n, 5) and r["Hi"].shape == (5, m)
    assert (r["Wi"][:, 2:4] == 1).all() and (r["Hi"][4:, :] == 1).all()
    assert r["Wm"][:, 2:4].all() and not r["Wm"][:, :2].any() and not r["Wm"][:, 4:].any()
    assert r["Hm"][4:, :].all() and not r["Hm"][:4, :].any()


def test_reformat_input_mask_shape_error():
    with pytest.raises(ValueError):
        reformat_input(None, {"W": np.zeros((3, 3), dtype=bool)}, 7, 5, 2)
    with pytest.raises(ValueError):
        reformat_input(None, {"W": np.zeros((7, 2))}, 7, 5, 2)       # not logical


def test_mse_mkl():
    obs = np.array([1.0, 2.0, np.nan]); pred = np.array([1.5, 2.0, 3.0])
    e = mse_mkl(obs, pred)
    assert e["MSE"] == pytest.approx(0.125)
    assert e["MKL"] == pytest.approx(((1 + 1e-16) * np.log((1 + 1e-16) / (1.5 + 1e-16)) - 1 + 1.5) / 2)


def test_nnlm_dimension_error_message():
    with pytest.raises(ValueError, match="Dimensions of x and y do not match."):      # tests/testthat/test-nnlm.R:54
        nnlm_b200.nnlm(np.random.default_rng(0).random((5, 4)), np.ones(4))


def test_nnmf_check_k():
    A = np.random.default_rng(0).random((50, 10))
    with pytest.raises(ValueError, match="is not recommended"):                       # tests/testthat/test-nnmf.R:60
        nnlm_b200.nnmf(A, 20)
