"""nnmf() / nnlm(): the reference's R front-end restated in Python over the C ABI.

Every default, coercion and error message follows the R source (cited inline) so the tests read like the reference's
testthat files. Names use `_` where R uses `.` (max.iter -> max_iter, ...).
"""
from __future__ import annotations

import ctypes as C
import math
import time
import warnings
from dataclasses import dataclass, field

import numpy as np

from . import _capi as K

WARN_NOT_CONVERGED = "Target tolerance not reached. Try a larger max.iter."       # src/nnmf.cpp:209


def get_method_code(method="scd", loss="mse") -> int:
    """R/misc.R:28-35: 1 = scd+mse, 2 = lee+mse, 3 = scd+mkl, 4 = lee+mkl."""
    if method not in ("scd", "lee"):
        raise ValueError("'arg' should be one of 'scd', 'lee'")
    if loss not in ("mse", "mkl"):
        raise ValueError("'arg' should be one of 'mse', 'mkl'")
    return 1 + (2 if loss == "mkl" else 0) + (1 if method == "lee" else 0)


def mse_mkl(obs, pred, na_rm=True, show_warning=True):
    """R/misc.R:9-16 (host-side helper of nnlm(); plain numpy, not on the accelerated path)."""
    obs = np.asarray(obs, dtype=np.float64)
    pred = np.asarray(pred, dtype=np.float64)
    mean = np.nanmean if na_rm else np.mean
    if not show_warning and ((obs[np.isfinite(obs)] < 0).any() or (pred[np.isfinite(pred)] < 0).any()):
        mkl = float("nan")
    else:
        with np.errstate(divide="ignore", invalid="ignore"):
            mkl = float(mean((obs + 1e-16) * np.log((obs + 1e-16) / (pred + 1e-16)) - obs + pred))
    mse = float(mean((obs - pred) ** 2))
    return {"MSE": mse, "MKL": mkl}


def _empty(x):
    return x is None or np.size(x) == 0


def reformat_input(init, mask, n, m, k, rng=None):
    """R/misc.R:48-129. Builds Wi (n x K), Hi (K x m), Wm, Hm from init/mask lists with optional known profiles
    W0 (n x kW0, fixed) / H0 (kH0 x m, fixed) and their free partners H1 / W1. Missing pieces of a supplied init are
    drawn uniform(0,1) (R: runif) from `rng`. Returns dict(Wi, Hi, Wm, Hm, kW0, kH0, K); Wi/Hi/Wm/Hm are None when the
    corresponding R matrix would have zero rows/columns (-> C++ default init / no mask)."""
    mask = dict(mask or {})
    init = dict(init or {})
    rng = rng or np.random.default_rng()
    known_w = not _empty(init.get("W0"))
    known_h = not _empty(init.get("H0"))
    kW0 = kH0 = 0
    if known_w:
        init["W0"] = np.asarray(init["W0"], dtype=np.float64).reshape(n, -1)
        kW0 = init["W0"].shape[1]
        mask["W0"] = np.ones((n, kW0), dtype=bool)
    else:
        mask.pop("W0", None); mask.pop("H1", None); init.pop("H1", None)
    if known_h:
        init["H0"] = np.asarray(init["H0"], dtype=np.float64).reshape(-1, m)
        kH0 = init["H0"].shape[0]
        mask["H0"] = np.ones((kH0, m), dtype=bool)
    else:
        mask.pop("H0", None); mask.pop("W1", None); init.pop("W1", None)
    Kt = k + kW0 + kH0

    def shapes(src):
        ew = not all(_empty(src.get(x)) for x in ("W", "W0", "W1"))
        eh = not all(_empty(src.get(x)) for x in ("H", "H0", "H1"))
        return {"W": (n, k * ew), "W0": (n, kW0 * ew), "W1": (n, kH0 * ew),
                "H": (k * eh, m), "H1": (kW0 * eh, m), "H0": (kH0 * eh, m)}

    dm = shapes(mask)
    for name in ("W", "W0", "W1", "H", "H0", "H1"):
        if _empty(mask.get(name)):
            mask[name] = np.zeros(dm[name], dtype=bool)
        else:
            a = np.asarray(mask[name])
            if a.shape != dm[name]:
                raise ValueError(f"Dimension of matrix mask${name} is expected to be {dm[name]}, but got {a.shape}")
            if a.dtype != bool:
                raise ValueError(f"Matrix mask${name} must be logical.")
            mask[name] = a
    di = shapes(init)
    for name in ("W", "W0", "W1", "H", "H0", "H1"):
        if _empty(init.get(name)):
            init[name] = rng.random(di[name])
        else:
            a = np.asarray(init[name], dtype=np.float64)
            if a.shape != di[name]:
                raise ValueError(f"Dimension of matrix init${name} is expected to be {di[name]}, but got {a.shape}")
            if not np.isfinite(a).all():
                raise ValueError(f"Matrix init${name} contains missing values.")
            init[name] = a
    Wm = np.concatenate([mask["W"], mask["W0"], mask["W1"]], axis=1)
    Hm = np.concatenate([mask["H"], mask["H1"], mask["H0"]], axis=0)
    Wi = np.concatenate([init["W"], init["W0"], init["W1"]], axis=1)
    Hi = np.concatenate([init["H"], init["H1"], init["H0"]], axis=0)
    return dict(Wm=Wm if Wm.size else None, Hm=Hm if Hm.size else None,
                Wi=Wi if Wi.size else None, Hi=Hi if Hi.size else None, kW0=kW0, kH0=kH0, K=Kt)


@dataclass
class Nnmf:
    """The `nnmf` S3 object of R/nnmf.R:207-224."""
    W: np.ndarray
    H: np.ndarray
    mse: np.ndarray
    mkl: np.ndarray
    target_loss: np.ndarray
    average_epochs: np.ndarray
    n_iteration: int
    run_time: float
    options: dict = field(default_factory=dict)
    stats: dict = field(default_factory=dict)
    converged: bool = True


@dataclass
class Nnlm:
    """The `nnlm` S3 object of R/nnlm.R:122-144."""
    coefficients: np.ndarray
    n_iteration: int
    error: dict
    options: dict = field(default_factory=dict)


def _options(precision, device, n_gpus=0, mkl_trace="all"):
    o = K.Options()
    o.precision = int(precision)
    o.device = int(device)
    o.n_gpus = int(n_gpus)
    if mkl_trace not in ("all", "final"):
        raise ValueError("mkl_trace must be 'all' or 'final'")
    o.mkl_trace = 1 if mkl_trace == "final" else 0
    return o


def nnmf(A, k=1, alpha=(0.0, 0.0, 0.0), beta=(0.0, 0.0, 0.0), method="scd", loss="mse", init=None, mask=None,
         W_norm=-1, check_k=True, max_iter=500, rel_tol=1e-4, n_threads=1, trace=None, verbose=0, show_warning=True,
         inner_max_iter=None, inner_rel_tol=1e-9, *, rng=None, interrupt=None, precision=K.PREC_AUTO, device=-1, n_gpus=0,
         mkl_trace="all"):
    """Non-negative matrix factorisation A ~ W H by alternating NNLS — R/nnmf.R:135-225 over nnlm_nnmf (c_nnmf).

    Extra keyword-only arguments (no counterpart in R): rng (numpy Generator for the default init — R's RNG stream is not
    reproduced), interrupt (callable polled once per outer iteration, like Rcpp::checkUserInterrupt), precision, device,
    n_gpus (0 = environment NNLM_B200_GPUS or 1; N > 1 shards this one call over N GPUs inside the library), mkl_trace
    ('all' = the reference: KL distance at every error record; 'final' = square-loss methods evaluate it for the last
    record only, earlier entries of `mkl` are NaN and tracing costs no pass over A).
    """
    code = get_method_code(method, loss)
    if inner_max_iter is None:
        inner_max_iter = 50 if loss == "mse" else 1                                  # R/nnmf.R:139
    if trace is None:
        trace = 100 / inner_max_iter                                                # R/nnmf.R:138
    A = K.f64(A, copy=False)
    if A.ndim != 2:
        raise ValueError("Matrix A must be numeric.")
    n, m = A.shape
    rng = rng or np.random.default_rng()
    im = reformat_input(init, mask, n, m, int(k), rng)
    Kt = im["K"]
    alpha = K.vec3(alpha)
    beta = K.vec3(beta)
    if check_k and np.all(alpha == 0) and np.all(beta == 0):                        # R/nnmf.R:157-166
        min_k = min(n, m)
        if Kt > min_k or np.isnan(A).any():
            isna = np.isnan(A)
            if isna.any():
                min_k = min(min_k, int((m - isna.sum(axis=1)).min()), int((n - isna.sum(axis=0)).min()))
            del isna
    else:
        min_k = Kt
    if check_k and Kt > min_k and np.all(alpha == 0) and np.all(beta == 0):
        raise ValueError(f"k larger than {min_k} is not recommended, unless properly masked or regularized.\n"
                         "Set check.k = FALSE if you want to skip this checking.")
    if n_threads < 0:
        n_threads = 0
    verbose = int(verbose)
    if trace <= 0:
        trace = 999999                                                              # R/nnmf.R:172-174
    trace = int(trace)
    max_iter = int(max_iter)

    # default init of src/nnmf.cpp:82-98 (0.01 * U(0,1), masked entries zero), drawn on the host: the core ABI always
    # receives explicit factors
    Wm = K.lgl(im["Wm"]); Hm = K.lgl(im["Hm"])
    if im["Wi"] is None:
        W = 0.01 * rng.random((n, Kt))
        if Wm is not None:
            W[Wm != 0] = 0.0
    else:
        W = im["Wi"]
    if im["Hi"] is None:
        H = 0.01 * rng.random((Kt, m))
        if Hm is not None:
            H[Hm != 0] = 0.0
    else:
        H = im["Hi"]
    W = K.f64(W); H = K.f64(H)

    tr = max(trace, 1)
    cap = int(math.ceil(max_iter / tr)) + 1
    mse = np.zeros(cap); mkl = np.zeros(cap); tgt = np.zeros(cap); ep = np.zeros(cap)
    n_err = C.c_uint32(0); n_iter = C.c_uint32(0); conv = C.c_int32(0)
    err = C.create_string_buffer(512)
    stats = K.Stats()
    opt = _options(precision, device, n_gpus, mkl_trace)
    cb = K.INTERRUPT_FN(lambda _u: 1 if interrupt() else 0) if interrupt is not None else K.INTERRUPT_FN()
    t0 = time.perf_counter()
    rc = K.lib().nnlm_nnmf(K.d(A), C.c_int64(n), C.c_int64(m), C.c_int32(Kt), K.d(W), K.d(H), K.i32(Wm), K.i32(Hm),
                           K.d(alpha), K.d(beta), C.c_uint32(max_iter), C.c_double(rel_tol), C.c_int32(n_threads),
                           C.c_int32(verbose), C.c_uint32(int(inner_max_iter)), C.c_double(inner_rel_tol),
                           C.c_int32(code), C.c_uint32(tr), K.d(mse), K.d(mkl), K.d(tgt), K.d(ep), C.c_uint32(cap),
                           C.byref(n_err), C.byref(n_iter), C.byref(conv), cb, None, C.byref(opt), C.byref(stats),
                           err, C.c_size_t(512))
    run_time = time.perf_counter() - t0
    K.check(rc, err)
    if show_warning and not conv.value:
        warnings.warn(WARN_NOT_CONVERGED, RuntimeWarning, stacklevel=2)
    ne = n_err.value
    if W_norm is not None and W_norm > 0:                                           # R/nnmf.R:197-205
        if math.isfinite(W_norm):
            scale = (W ** W_norm).sum(axis=0) ** (1.0 / W_norm)
        else:
            scale = W.max(axis=0)
        W = W / scale
        H = scale[:, None] * H
    return Nnmf(W=W, H=H, mse=mse[:ne].copy(), mkl=mkl[:ne].copy(), target_loss=tgt[:ne].copy(),
                average_epochs=ep[:ne].copy(), n_iteration=int(n_iter.value), run_time=run_time,
                options=dict(method=method, loss=loss, alpha=alpha, beta=beta, init=init, mask=mask,
                             n_threads=n_threads, trace=trace, verbose=verbose, max_iter=max_iter, rel_tol=rel_tol,
                             inner_max_iter=inner_max_iter, inner_rel_tol=inner_rel_tol),
                stats=stats.as_dict(), converged=bool(conv.value))


def nnlm(x, y, alpha=(0.0, 0.0, 0.0), method="scd", loss="mse", init=None, mask=None, check_x=True, max_iter=10000,
         rel_tol=1e-12, n_threads=1, show_warning=True, *, rng=None, precision=K.PREC_AUTO, device=-1):
    """Non-negative linear model y ~ x beta — R/nnlm.R:70-145 over nnlm_nnlm (c_nnlm)."""
    code = get_method_code(method, loss)
    x = K.f64(x, copy=False)
    y_arr = np.asarray(y, dtype=np.float64)
    is_y_vector = y_arr.ndim == 1
    y2 = K.f64(y_arr.reshape(-1, 1) if is_y_vector else y_arr, copy=False)
    if show_warning and loss == "mkl" and ((x < 0).any() or (y2[np.isfinite(y2)] < 0).any()):
        warnings.warn("x or y have negative values. One should instead use method == 'mse'.", RuntimeWarning, stacklevel=2)
    if not np.isfinite(x).all():
        raise ValueError("Matrix  contains missing values.")
    if x.shape[0] != y2.shape[0]:
        raise ValueError("Dimensions of x and y do not match.")                      # R/nnlm.R:86-87
    if max_iter <= 0:
        raise ValueError("max.iter must be positive.")
    n, p = x.shape
    q = y2.shape[1]
    if check_x:
        if n < p or np.linalg.cond(x) > 1.0 / np.finfo(np.float64).eps:         # R: rcond(x) < .Machine$double.eps
            warnings.warn("x does not have a full column rank. Solution may not be unique.", RuntimeWarning, stacklevel=2)
    alpha = K.vec3(alpha)
    if show_warning and alpha[0] < alpha[1]:
        warnings.warn("If alpha[1] < alpha[2], be aware that that algorithm may not converge or unique.", RuntimeWarning,
                      stacklevel=2)
    mk = None
    if not _empty(mask):
        mk = np.asarray(mask)
        if mk.shape != (p, q) or mk.dtype != bool:
            raise ValueError("Matrix  must be logical.")
    if _empty(init):
        rng = rng or np.random.default_rng()
        coef = rng.random((p, q))                                                   # beta.randu(), src/nnlm.cpp:38-39
        if mk is not None:
            coef = (~mk).astype(np.float64)                                         # R/nnlm.R:108-109
    else:
        coef = np.asarray(init, dtype=np.float64)
        if coef.shape != (p, q):
            raise ValueError(f"Dimension of matrix  is expected to be ({p}, {q})")
        if (coef < 0).any():
            raise ValueError("Matrix  must be non-negative.")
    coef = K.f64(coef)
    mkc = K.lgl(mk)
    nit = C.c_int64(0)
    err = C.create_string_buffer(512)
    opt = _options(precision, device)
    rc = K.lib().nnlm_nnlm(K.d(x), K.d(y2), C.c_int64(n), C.c_int64(p), C.c_int64(q), K.d(coef), K.i32(mkc), K.d(alpha),
                           C.c_uint32(int(max_iter)), C.c_double(rel_tol), C.c_int32(n_threads), C.c_int32(code),
                           C.byref(nit), C.byref(opt), None, err, C.c_size_t(512))
    K.check(rc, err)
    pred = x @ coef
    e = mse_mkl(y2, pred, na_rm=True, show_warning=False)
    target = 0.5 * e["MSE"] if loss == "mse" else e["MKL"]
    target += (alpha[0] - alpha[1]) * float((coef ** 2).sum()) + alpha[1] * float((coef.sum(axis=0) ** 2).sum()) \
        + alpha[2] * float(coef.sum())                                               # R/nnlm.R:136-137
    e["target.error"] = target
    out = coef[:, 0].copy() if is_y_vector else coef
    return Nnlm(coefficients=out, n_iteration=int(nit.value), error=e,
                options=dict(method=method, loss=loss, max_iter=max_iter, rel_tol=rel_tol))


def predict(obj: Nnmf, newdata=None, which="A", method=None, loss=None, **kw):
    """predict.nnmf — R/nnmf_methods.R:22-48: 'A' returns W H; 'W' / 'H' solve for new rows / columns with nnlm()."""
    if which not in ("A", "W", "H"):
        raise ValueError("'arg' should be one of 'A', 'W', 'H'")
    method = method or obj.options.get("method", "scd")
    loss = loss or obj.options.get("loss", "mse")
    if which == "A":
        return obj.W @ obj.H
    x = np.asarray(newdata, dtype=np.float64)
    if x.ndim != 2:
        raise ValueError("Matrix newdata must be numeric.")
    if which == "W":
        if x.shape[1] != obj.H.shape[1]:
            raise ValueError(f"Dimension of matrix newdata is expected to be (NA, {obj.H.shape[1]}), but got {x.shape}")
        out = nnlm(obj.H.T, x.T, method=method, loss=loss, **kw)
        out.coefficients = out.coefficients.T
        return out
    if x.shape[0] != obj.W.shape[0]:
        raise ValueError(f"Dimension of matrix newdata is expected to be ({obj.W.shape[0]}, NA), but got {x.shape}")
    return nnlm(obj.W, x, method=method, loss=loss, **kw)


def nnlm_update(H, Wt, A, mask=None, beta=(0.0, 0.0, 0.0), max_iter=10, rel_tol=1e-8, n_threads=1, method=1,
                with_missing=-1, *, precision=K.PREC_AUTO, device=-1):
    """One half-iteration — update()/update_with_missing() of src/update_with_missing.cpp. Returns (H_new, total_iter)."""
    H = K.f64(H); Wt = K.f64(Wt, copy=False); A = K.f64(A, copy=False); mk = K.lgl(mask)
    k, m = H.shape
    n = A.shape[0]
    if Wt.shape != (k, n) or A.shape != (n, m):
        raise ValueError("nnlm_update: shapes must be H k x m, Wt k x n, A n x m")
    b = K.vec3(beta)
    tot = C.c_int64(0)
    err = C.create_string_buffer(512)
    opt = _options(precision, device)
    rc = K.lib().nnlm_update(K.d(H), K.d(Wt), K.d(A), K.i32(mk), K.d(b), C.c_int32(k), C.c_int64(n), C.c_int64(m),
                             C.c_uint32(int(max_iter)), C.c_double(rel_tol), C.c_int32(n_threads), C.c_int32(method),
                             C.c_int32(with_missing), C.byref(tot), C.byref(opt), None, err, C.c_size_t(512))
    K.check(rc, err)
    return H, int(tot.value)


def na_mask(A):
    """Bit-plane of !isfinite(A) over the column-major linear index, built on the device (nnlm_na_mask).
    Returns (bits uint32[ceil(n*m/32)], per-column missing counts int64[m])."""
    A = K.f64(A, copy=False)
    n, m = A.shape
    words = (n * m + 31) // 32
    bits = np.zeros(words, dtype=np.uint32)
    cols = np.zeros(m, dtype=np.int64)
    err = C.create_string_buffer(512)
    rc = K.lib().nnlm_na_mask(K.d(A), C.c_int64(n), C.c_int64(m), bits.ctypes.data_as(C.POINTER(C.c_uint32)),
                              cols.ctypes.data_as(C.POINTER(C.c_int64)), err, C.c_size_t(512))
    K.check(rc, err)
    return bits, cols


def na_corrections(Wt, A, *, device=-1):
    """Diagnostic: per-column Gram / row-sum corrections of the NA path from the tensor-core mask contraction
    (nnlm_na_corrections). Returns (S_pairs m x k(k+1)/2 with pair (a >= b) at a(a+1)/2 + b, S_rows m x k)."""
    Wt = K.f64(Wt, copy=False); A = K.f64(A, copy=False)
    k, n = Wt.shape
    m = A.shape[1]
    width = (k * (k + 1) // 2 + k + 127) // 128 * 128
    S = np.empty((m, width), dtype=np.float64)
    w = C.c_int64(0)
    err = C.create_string_buffer(512)
    opt = _options(K.PREC_FAST, device)
    rc = K.lib().nnlm_na_corrections(K.d(Wt), K.d(A), C.c_int32(k), C.c_int64(n), C.c_int64(m), K.d(S), C.c_int64(S.size),
                                     C.byref(w), C.byref(opt), err, C.c_size_t(512))
    K.check(rc, err)
    kk2 = k * (k + 1) // 2
    return S[:, :kk2].copy(), S[:, kk2:kk2 + k].copy()


def cross(Wt, A, *, precision=K.PREC_AUTO, device=-1):
    """Diagnostic: Q = Wt @ A (non-finite entries of A read as zero) through the library's cross-product kernels."""
    Wt = K.f64(Wt, copy=False); A = K.f64(A, copy=False)
    k, n = Wt.shape
    m = A.shape[1]
    assert A.shape[0] == n
    Q = np.empty((k, m), dtype=np.float64, order="F")
    err = C.create_string_buffer(512)
    opt = _options(precision, device)
    st = K.Stats()
    rc = K.lib().nnlm_cross(K.d(Wt), K.d(A), C.c_int32(k), C.c_int64(n), C.c_int64(m), K.d(Q), C.byref(opt), C.byref(st),
                            err, C.c_size_t(512))
    K.check(rc, err)
    return Q, st.as_dict()
