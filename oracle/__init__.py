"""ctypes binding of the CPU oracle (oracle/nnlm_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
--impl reference arm. Nothing under nnlm_b200/ imports this package.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnnlm_oracle.so")
_STAMP = _SO + ".cpu"
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _cpu_stamp() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    """Compile the oracle with the distro g++ (-march=native). Rebuilt when the host CPU differs from the one the
    existing .so was built on (the .so travels from the build container to the GPU box)."""
    stamp = _cpu_stamp()
    src = os.path.join(_HERE, "nnlm_oracle.cpp")
    fresh = (os.path.exists(_SO) and os.path.exists(_STAMP) and open(_STAMP).read().strip() == stamp
             and os.path.getmtime(_SO) >= os.path.getmtime(src))
    if fresh and not force:
        return _SO
    subprocess.run(["make", "-C", _HERE, "-B", "libnnlm_oracle.so"], check=True, capture_output=True)
    with open(_STAMP, "w") as f:
        f.write(stamp)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(x):
    return np.asfortranarray(np.array(x, dtype=np.float64, copy=True, order="F"))


def _mask(x):
    return None if x is None else np.asfortranarray(np.array(x, order="F").astype(np.int32))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def splitmix_uniform(seed: int, count: int, offset: int = 0) -> np.ndarray:
    """u(seed, idx) of SURVEY.md §8d (numpy twin of the generators in oracle_synth_block and csrc/synth.cu)."""
    out = np.empty(count, dtype=np.float64)
    step = 1 << 24
    with np.errstate(over="ignore"):
        for a in range(0, count, step):
            b = min(count, a + step)
            z = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.arange(offset + a, offset + b, dtype=np.uint64)
            z = z + np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[a:b] = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return out


def set_threads(n: int) -> None:
    """Set the OpenMP thread count (torchrun exports OMP_NUM_THREADS=1 to its workers; bench.py's reference arm calls
    this with len(os.sched_getaffinity(0)) so the CPU arm uses all host cores at every N)."""
    lib().oracle_set_threads(C.c_int(int(n)))


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def synth_block(n_global, row0, nr, col0, mc, k, seed_base=0, noise=0.1, na_frac=0.0, n_threads=0):
    """Host twin of the device generator (nnlm_synth_block): rows [row0,row0+nr) x columns [col0,col0+mc) of the
    synthetic matrix of SURVEY.md §8d, bit-identical to the GPU arm's matrix."""
    A = np.empty((nr, mc), dtype=np.float64, order="F")
    lib().oracle_synth_block(_d(A), C.c_int64(n_global), C.c_int64(row0), C.c_int64(nr), C.c_int64(col0), C.c_int64(mc),
                             C.c_int32(k), C.c_uint64(seed_base), C.c_double(noise), C.c_double(na_frac), C.c_int32(n_threads))
    return A


def synth_matrix(n, m, k, seed_base=0, noise=0.1, na_frac=0.0, n_threads=0):
    return synth_block(n, 0, n, 0, m, k, seed_base, noise, na_frac, n_threads)


def set_faithful_transpose(on: bool) -> None:
    lib().oracle_set_faithful_transpose(C.c_int(1 if on else 0))


def transpose(A, n_threads=0):
    """A.t() as the reference materialises it each iteration (src/nnmf.cpp:117,131)."""
    A = np.asfortranarray(A, dtype=np.float64)
    n, m = A.shape
    At = np.empty((m, n), dtype=np.float64, order="F")
    lib().oracle_transpose(_d(A), C.c_int64(n), C.c_int64(m), _d(At), C.c_int32(n_threads))
    return At


def update(H, Wt, A, mask=None, beta=(0.0, 0.0, 0.0), max_iter=10, rel_tol=1e-8, n_threads=1, method=1,
           with_missing=-1):
    """update()/update_with_missing() (src/update_with_missing.cpp). Returns (H_new, total_iter)."""
    H = _f64(H); Wt = np.asfortranarray(Wt, dtype=np.float64); A = np.asfortranarray(A, dtype=np.float64); mk = _mask(mask)
    k, m = H.shape
    n = A.shape[0]
    assert Wt.shape == (k, n) and A.shape == (n, m)
    b = np.array(list(beta) + [0.0] * 3, dtype=np.float64)[:3]
    tot = C.c_int64(0)
    err = C.create_string_buffer(256)
    rc = lib().oracle_update(_d(H), _d(Wt), _d(A), _i(mk), _d(b), C.c_int32(k), C.c_int64(n), C.c_int64(m),
                             C.c_uint32(max_iter), C.c_double(rel_tol), C.c_int32(n_threads), C.c_int32(method),
                             C.c_int32(with_missing), C.byref(tot), None, None, err, C.c_size_t(256))
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return H, int(tot.value)


def nnmf(A, K, W, H, Wm=None, Hm=None, alpha=(0, 0, 0), beta=(0, 0, 0), max_iter=500, rel_tol=1e-4, n_threads=1,
         verbose=0, inner_max_iter=50, inner_rel_tol=1e-9, method=1, trace=1):
    """c_nnmf (src/nnmf.cpp) with explicit init. Returns a dict with the reference's seven outputs + converged."""
    A = _f64(A); W = _f64(W); H = _f64(H); Wm = _mask(Wm); Hm = _mask(Hm)
    n, m = A.shape
    assert W.shape == (n, K) and H.shape == (K, m)
    a = np.array(list(alpha) + [0.0] * 3, dtype=np.float64)[:3]
    b = np.array(list(beta) + [0.0] * 3, dtype=np.float64)[:3]
    tr = max(int(trace), 1)
    cap = int(math.ceil(max_iter / tr)) + 1
    mse = np.zeros(cap); mkl = np.zeros(cap); tgt = np.zeros(cap); ep = np.zeros(cap)
    n_err = C.c_uint32(0); n_iter = C.c_uint32(0); conv = C.c_int32(0)
    err = C.create_string_buffer(256)
    rc = lib().oracle_nnmf(_d(A), C.c_int64(n), C.c_int64(m), C.c_int32(K), _d(W), _d(H), _i(Wm), _i(Hm), _d(a), _d(b),
                           C.c_uint32(max_iter), C.c_double(rel_tol), C.c_int32(n_threads), C.c_int32(verbose),
                           C.c_uint32(inner_max_iter), C.c_double(inner_rel_tol), C.c_int32(method), C.c_uint32(tr),
                           _d(mse), _d(mkl), _d(tgt), _d(ep), C.c_uint32(cap),
                           C.byref(n_err), C.byref(n_iter), C.byref(conv), None, None, None, None,
                           err, C.c_size_t(256))
    if rc != 0:
        raise RuntimeError(err.value.decode())
    ne = n_err.value
    return dict(W=W, H=H, mse=mse[:ne].copy(), mkl=mkl[:ne].copy(), target_loss=tgt[:ne].copy(),
                average_epochs=ep[:ne].copy(), n_iteration=int(n_iter.value), converged=bool(conv.value))


def nnlm(x, y, coef0, mask=None, alpha=(0, 0, 0), max_iter=10000, rel_tol=1e-12, n_threads=1, method=1):
    """c_nnlm (src/nnlm.cpp) with explicit beta0. Returns (coefficients p x q, n_iteration)."""
    x = _f64(x); y = _f64(y)
    if y.ndim == 1:
        y = _f64(y.reshape(-1, 1))
    n, p = x.shape
    q = y.shape[1]
    coef = _f64(coef0)
    assert coef.shape == (p, q) and y.shape[0] == n
    mk = _mask(mask)
    a = np.array(list(alpha) + [0.0] * 3, dtype=np.float64)[:3]
    nit = C.c_int64(0)
    err = C.create_string_buffer(256)
    rc = lib().oracle_nnlm(_d(x), _d(y), C.c_int64(n), C.c_int64(p), C.c_int64(q), _d(coef), _i(mk), _d(a),
                           C.c_uint32(max_iter), C.c_double(rel_tol), C.c_int32(n_threads), C.c_int32(method),
                           C.byref(nit), None, None, err, C.c_size_t(256))
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return coef, int(nit.value)
