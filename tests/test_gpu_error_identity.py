"""Error evaluation from quantities already on the device (src/nnmf.cpp:135-140):
sum (A - W'H)^2 = ||A||^2 - 2 <H, WtA> + <WtW, HHt>, all fp64, against the fused pass over A (k_error) and the oracle."""
import numpy as np
import pytest

import nnlm_b200
import oracle
from nnlm_b200 import _capi as K
from nnlm_b200.session import Session
from conftest import umat

pytestmark = pytest.mark.gpu


def problem(n, m, k):
    return oracle.synth_matrix(n, m, k), 0.01 * umat(11, n, k), 0.01 * umat(12, k, m)


def identity_bound(A, mse, prec):
    """The identity subtracts quantities of size ||A||^2 to get N*mse: its relative error is the relative error of the
    cross-product term times ||A||^2 / (N mse). fp64 path: 1e-15 x that (asserted as 1e-10 flat). Tensor-core path: the
    fp32 TMEM accumulation leaves a ~1e-8 relative, nearly constant bias on <H, WtA>, so 3e-8 x that."""
    if prec == K.PREC_EXACT:
        return 1e-10
    return 3e-8 * float((A ** 2).sum()) / (A.size * mse)


@pytest.mark.parametrize("prec", [K.PREC_EXACT, K.PREC_FAST])
@pytest.mark.parametrize("n,m,k", [(1200, 700, 8), (3000, 1500, 50)])
def test_identity_mse_matches_fused_pass(prec, n, m, k):
    A, W0, H0 = problem(n, m, k)
    with Session(A, k=k, method=1, precision=prec) as s:
        s.set_factors(W0, H0)
        s.run(3)
        mse_id, used = s.mse()
        mse_full, _, _ = s.error()
        assert used, "the identity must be available right after run()"
        print(f"{n}x{m} k={k} prec={prec}: identity {mse_id:.12e}, fused pass {mse_full:.12e}, rel {abs(mse_id / mse_full - 1):.1e}")
        tol = identity_bound(A, mse_full, prec)
        assert abs(mse_id - mse_full) <= tol * mse_full, (tol, abs(mse_id / mse_full - 1))
        # stale after new factors: falls back to the fused pass
        s.set_factors(W0, H0)
        _, used2 = s.mse()
        assert not used2


def test_identity_with_masks_and_penalties():
    n, m, k = 900, 400, 6
    A, W0, H0 = problem(n, m, k)
    rng = np.random.default_rng(2)
    Wm = rng.random((n, k)) < 0.1; Hm = rng.random((k, m)) < 0.1
    W0[Wm] = 0; H0[Hm] = 0
    with Session(A, k=k, method=2, precision=K.PREC_EXACT, Wm=Wm, Hm=Hm, alpha=(0.1, 0.02, 0.01), beta=(0.05, 0.0, 0.02)) as s:
        s.set_factors(W0, H0)
        s.run(4)
        a, used = s.mse()
        b, _, _ = s.error()
        assert used and abs(a - b) <= 1e-10 * b


@pytest.mark.parametrize("prec", [K.PREC_EXACT, K.PREC_FAST])
def test_nnmf_mkl_trace_final_matches_reference_vectors(prec):
    """mkl_trace='final': mse / target / epochs at every record as the reference, mkl only at the last record."""
    n, m, k = 1500, 700, 6
    A, W0, H0 = problem(n, m, k)
    ref = oracle.nnmf(A, k, W0, H0, max_iter=12, rel_tol=-1, n_threads=0, inner_max_iter=50, method=1, trace=2)
    got = nnlm_b200.nnmf(A, k, init={"W": W0, "H": H0}, max_iter=12, rel_tol=-1, trace=2, show_warning=False,
                         precision=prec, mkl_trace="final")
    assert got.stats["mse_from_identity"] == 1
    rt = 1e-9 if prec == K.PREC_EXACT else max(1e-5, identity_bound(A, ref["mse"].min(), prec))
    np.testing.assert_allclose(got.mse, ref["mse"], rtol=rt)
    np.testing.assert_allclose(got.target_loss, ref["target_loss"], rtol=rt)
    np.testing.assert_allclose(got.average_epochs, ref["average_epochs"], rtol=2e-3)
    assert np.isnan(got.mkl[:-1]).all()
    np.testing.assert_allclose(got.mkl[-1], ref["mkl"][-1], rtol=1e-6)
    full = nnlm_b200.nnmf(A, k, init={"W": W0, "H": H0}, max_iter=12, rel_tol=-1, trace=2, show_warning=False, precision=prec)
    assert full.stats["mse_from_identity"] == 0
    np.testing.assert_allclose(full.mkl, ref["mkl"], rtol=1e-6)
    np.testing.assert_allclose(full.W, got.W, rtol=0, atol=0)            # tracing never changes the factors


def test_missing_path_keeps_the_fused_pass():
    n, m, k = 600, 300, 4
    A = oracle.synth_matrix(n, m, k, na_frac=0.1)
    W0 = 0.01 * umat(11, n, k); H0 = 0.01 * umat(12, k, m)
    ref = oracle.nnmf(A, k, W0, H0, max_iter=4, rel_tol=-1, n_threads=0, inner_max_iter=50, method=1, trace=1)
    got = nnlm_b200.nnmf(A, k, init={"W": W0, "H": H0}, max_iter=4, rel_tol=-1, trace=1, show_warning=False,
                         precision=K.PREC_EXACT, mkl_trace="final")
    assert got.stats["mse_from_identity"] == 0
    np.testing.assert_allclose(got.mse, ref["mse"], rtol=1e-9)
    np.testing.assert_allclose(got.mkl, ref["mkl"], rtol=1e-7)


def test_predict_mirrors_the_reference_method():
    """predict.nnmf (R/nnmf_methods.R:22-48): 'A' = W H; 'H' / 'W' solve new columns / rows with nnlm()."""
    n, m, k = 300, 120, 4
    A, W0, H0 = problem(n, m, k)
    r = nnlm_b200.nnmf(A, k, init={"W": W0, "H": H0}, max_iter=30, rel_tol=-1, show_warning=False, precision=K.PREC_EXACT)
    np.testing.assert_allclose(nnlm_b200.predict(r), r.W @ r.H)
    newx = oracle.synth_matrix(n, 7, k, seed_base=20)
    rng = np.random.default_rng(1)
    init = rng.random((k, 7))
    ph = nnlm_b200.predict(r, newx, "H", init=init)
    ref, _ = oracle.nnlm(r.W, newx, init, max_iter=10000, rel_tol=1e-12)
    np.testing.assert_allclose(ph.coefficients, ref, rtol=1e-7, atol=1e-10)
    neww = oracle.synth_matrix(5, m, k, seed_base=30)
    initw = rng.random((k, 5))
    pw = nnlm_b200.predict(r, neww, "W", init=initw)
    refw, _ = oracle.nnlm(r.H.T, neww.T, initw, max_iter=10000, rel_tol=1e-12)
    assert pw.coefficients.shape == (5, k)
    np.testing.assert_allclose(pw.coefficients, refw.T, rtol=1e-7, atol=1e-10)
    with pytest.raises(ValueError):
        nnlm_b200.predict(r, newx[:-1], "H")


@pytest.mark.parametrize("n,m,k,na", [(1000, 300, 6, 0.0), (2500, 1300, 50, 0.0), (1300, 700, 100, 0.0), (1500, 640, 20, 0.15)])
def test_tensor_core_error_evaluation_matches_the_oracle(n, m, k, na):
    """error_tc.cu (fast precision): mse and mkl of a mid-trajectory (W, H) from the tcgen05 tile of W'H folded against A in the
    epilogue, vs the oracle's mse / mkl vectors and vs the fp64 pass of error_eval.cu (NNLM_ERR_FP64 not set here)."""
    A = oracle.synth_matrix(n, m, min(k, 20), na_frac=na)
    W0 = 0.01 * umat(11, n, k); H0 = 0.01 * umat(12, k, m)
    ref = oracle.nnmf(A, k, W0, H0, max_iter=4, rel_tol=-1, n_threads=0, inner_max_iter=50, method=1, trace=1)
    got = nnlm_b200.nnmf(A, k, init={"W": W0, "H": H0}, max_iter=4, rel_tol=-1, trace=1, show_warning=False, check_k=False,
                         precision=K.PREC_FAST)
    exact = nnlm_b200.nnmf(A, k, init={"W": W0, "H": H0}, max_iter=4, rel_tol=-1, trace=1, show_warning=False, check_k=False,
                           precision=K.PREC_EXACT)
    print(f"{n}x{m} k={k} na={na}: mse rel {np.abs(got.mse / ref['mse'] - 1).max():.1e}, mkl rel {np.abs(got.mkl / ref['mkl'] - 1).max():.1e} "
          f"(fp64 pass: {np.abs(exact.mse / ref['mse'] - 1).max():.1e}, {np.abs(exact.mkl / ref['mkl'] - 1).max():.1e})")
    # k = 100 over a rank-20 matrix: the trajectory itself is chaotic from the second iteration on (the fp64 path
    # drifts from the oracle just the same), so only the first record isolates the evaluation there
    upto = 1 if k > 50 else len(ref["mse"])
    np.testing.assert_allclose(got.mse[:upto], ref["mse"][:upto], rtol=2e-6)
    np.testing.assert_allclose(got.mkl[:upto], ref["mkl"][:upto], rtol=2e-6)
    np.testing.assert_allclose(got.target_loss[:upto], ref["target_loss"][:upto], rtol=2e-6)
    # and the two evaluations agree on the SAME factors: the fp64 pass on the fast path's final (W, H)
    with Session(A, k=k, method=1, precision=K.PREC_FAST) as s:
        s.set_factors(got.W, got.H)
        mse_tc, mkl_tc, _ = s.error()
    with Session(A, k=k, method=1, precision=K.PREC_EXACT) as s:
        s.set_factors(got.W, got.H)
        mse_64, mkl_64, _ = s.error()
    assert abs(mse_tc / mse_64 - 1) < 2e-6 and abs(mkl_tc / mkl_64 - 1) < 2e-6, (mse_tc, mse_64, mkl_tc, mkl_64)


def test_tensor_core_error_with_vanishing_reconstruction():
    """Entries where W'H is exactly 0 against a > 0 (masked-out factors): the KL term stays finite and equal to the reference's
    -(a+e) log(0+e) (src/nnmf.cpp:139), through the branch that avoids 1 + x in fp32."""
    n, m, k = 700, 260, 4
    A = oracle.synth_matrix(n, m, k)
    W0 = 0.01 * umat(11, n, k); H0 = 0.01 * umat(12, k, m)
    Hm = np.zeros((k, m), dtype=bool); Hm[:, :7] = True          # seven columns of H pinned to zero -> Ahat = 0 there
    H0[Hm] = 0
    ref = oracle.nnmf(A, k, W0, H0, Hm=Hm, max_iter=3, rel_tol=-1, n_threads=0, inner_max_iter=50, method=1, trace=1)
    got = nnlm_b200.nnmf(A, k, init={"W": W0, "H": H0}, mask={"H": Hm}, max_iter=3, rel_tol=-1, trace=1, show_warning=False,
                         check_k=False, precision=K.PREC_FAST)
    assert np.isfinite(got.mkl).all()
    np.testing.assert_allclose(got.mkl, ref["mkl"], rtol=1e-5)
    np.testing.assert_allclose(got.mse, ref["mse"], rtol=2e-6)
