"""Device-resident ANLS session: A stays in HBM, factors move on request. This is the benchmark's "inputs already
resident" path (nnlm_session_* in include/nnlm_b200.h); nnmf() itself uploads and frees A on every call like c_nnmf."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as K


def splitmix_uniform(seed: int, count: int, offset: int = 0) -> np.ndarray:
    """numpy twin of the device generator in csrc/synth.cu."""
    with np.errstate(over="ignore"):
        idx = np.arange(offset, offset + count, dtype=np.uint64)
        z = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + idx
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synth_init(n, m, k, seed_base=0):
    """W0 = 0.01*u(base+11) (n x k), H0 = 0.01*u(base+12) (k x m): the scale of src/nnmf.cpp:85,95."""
    W0 = 0.01 * splitmix_uniform(seed_base + 11, n * k).reshape((n, k), order="F")
    H0 = 0.01 * splitmix_uniform(seed_base + 12, k * m).reshape((k, m), order="F")
    return W0, H0


def synth_block(n_global, row0, nr, col0, mc, k, seed_base=0, noise=0.1, na_frac=0.0):
    """Rows [row0, row0+nr) x columns [col0, col0+mc) of the synthetic matrix, generated on the device (nnlm_synth_block)."""
    A = np.empty((nr, mc), dtype=np.float64, order="F")
    err = C.create_string_buffer(512)
    rc = K.lib().nnlm_synth_block(K.d(A), C.c_int64(n_global), C.c_int64(row0), C.c_int64(nr), C.c_int64(col0), C.c_int64(mc),
                                  C.c_int32(k), C.c_uint64(seed_base), C.c_double(noise), C.c_double(na_frac), err,
                                  C.c_size_t(512))
    K.check(rc, err)
    return A


def synth_matrix(n, m, k, col0=0, seed_base=0, noise=0.1, na_frac=0.0):
    """The synthetic A of BASELINE.md §4, generated on the device and copied to the host (nnlm_synth_matrix)."""
    A = np.empty((n, m), dtype=np.float64, order="F")
    err = C.create_string_buffer(512)
    rc = K.lib().nnlm_synth_matrix(K.d(A), C.c_int64(n), C.c_int64(m), C.c_int32(k), C.c_int64(col0),
                                   C.c_uint64(seed_base), C.c_double(noise), C.c_double(na_frac), err, C.c_size_t(512))
    K.check(rc, err)
    return A


class Session:
    def __init__(self, A=None, k=1, method=1, alpha=(0, 0, 0), beta=(0, 0, 0), inner_max_iter=50, inner_rel_tol=1e-9,
                 Wm=None, Hm=None, precision=K.PREC_AUTO, device=-1, synthetic=None, timing=False, comm=None, shards=None,
                 shape=None):
        """A: host matrix (n x m), or synthetic=dict(n=, m=, seed_base=0, noise=0.1, na_frac=0.0)."""
        self._h = C.c_void_p()
        self.k = int(k)
        a = K.vec3(alpha); b = K.vec3(beta)
        opt = K.Options(); opt.precision = int(precision); opt.device = int(device); opt.verbose_timing = int(bool(timing))
        err = C.create_string_buffer(512)
        self._comm = comm              # keep the communicator alive as long as the session
        if comm is not None:
            opt.comm = comm.handle
            if synthetic is None and shards is None:
                raise ValueError("a sharded session needs synthetic=... or shards=(Acol, Arow) with shape=(n, m)")
        if synthetic is not None:
            self.n, self.m = int(synthetic["n"]), int(synthetic["m"])
            rc = K.lib().nnlm_session_create_synthetic(
                C.byref(self._h), C.c_int64(self.n), C.c_int64(self.m), C.c_int32(self.k),
                C.c_uint64(int(synthetic.get("seed_base", 0))), C.c_double(synthetic.get("noise", 0.1)),
                C.c_double(synthetic.get("na_frac", 0.0)), K.d(a), K.d(b), C.c_uint32(int(inner_max_iter)),
                C.c_double(inner_rel_tol), C.c_int32(method), C.byref(opt), err, C.c_size_t(512))
        elif shards is not None:
            # host shards of this rank: Acol = A[:, c0:c0+mc] (n x mc), Arow = A[r0:r0+nr, :] (nr x m)
            Acol = K.f64(shards[0], copy=False); Arow = K.f64(shards[1], copy=False)
            self.n, self.m = int(shape[0]), int(shape[1])
            wm = K.lgl(Wm); hm = K.lgl(Hm)
            rc = K.lib().nnlm_session_create_sharded(
                C.byref(self._h), K.d(Acol), K.d(Arow), C.c_int64(self.n), C.c_int64(self.m), C.c_int32(self.k), K.i32(wm),
                K.i32(hm), K.d(a), K.d(b), C.c_uint32(int(inner_max_iter)), C.c_double(inner_rel_tol), C.c_int32(method),
                C.byref(opt), err, C.c_size_t(512))
        else:
            A = K.f64(A, copy=False)
            self.n, self.m = A.shape
            wm = K.lgl(Wm); hm = K.lgl(Hm)
            rc = K.lib().nnlm_session_create(
                C.byref(self._h), K.d(A), C.c_int64(self.n), C.c_int64(self.m), C.c_int32(self.k), K.i32(wm), K.i32(hm),
                K.d(a), K.d(b), C.c_uint32(int(inner_max_iter)), C.c_double(inner_rel_tol), C.c_int32(method),
                C.byref(opt), err, C.c_size_t(512))
        K.check(rc, err)

    def set_factors(self, W, H):
        W = K.f64(W, copy=False); H = K.f64(H, copy=False)
        assert W.shape == (self.n, self.k) and H.shape == (self.k, self.m)
        err = C.create_string_buffer(512)
        K.check(K.lib().nnlm_session_set_factors(self._h, K.d(W), K.d(H), err, C.c_size_t(512)), err)

    def get_factors(self):
        W = np.empty((self.n, self.k), order="F"); H = np.empty((self.k, self.m), order="F")
        err = C.create_string_buffer(512)
        K.check(K.lib().nnlm_session_get_factors(self._h, K.d(W), K.d(H), err, C.c_size_t(512)), err)
        return W, H

    def run(self, iters):
        """`iters` ANLS iterations (W-half then H-half). Returns (device milliseconds, summed inner sweeps)."""
        ms = C.c_double(0); sw = C.c_int64(0)
        err = C.create_string_buffer(512)
        K.check(K.lib().nnlm_session_run(self._h, C.c_uint32(int(iters)), C.byref(ms), C.byref(sw), err, C.c_size_t(512)), err)
        return ms.value, sw.value

    def error(self):
        mse = C.c_double(0); mkl = C.c_double(0); tgt = C.c_double(0)
        err = C.create_string_buffer(512)
        K.check(K.lib().nnlm_session_error(self._h, C.byref(mse), C.byref(mkl), C.byref(tgt), err, C.c_size_t(512)), err)
        return mse.value, mkl.value, tgt.value

    def mse(self):
        """(mse, from_identity): the MSE from the Gram identity when the last H-half's products are current (nnlm_session_mse)."""
        v = C.c_double(0); f = C.c_int32(0)
        err = C.create_string_buffer(512)
        K.check(K.lib().nnlm_session_mse(self._h, C.byref(v), C.byref(f), err, C.c_size_t(512)), err)
        return v.value, bool(f.value)

    def stats(self):
        st = K.Stats()
        K.lib().nnlm_session_stats(self._h, C.byref(st))
        return st.as_dict()

    def reset_stats(self):
        K.lib().nnlm_session_reset_stats(self._h)

    def close(self):
        if self._h:
            K.lib().nnlm_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
