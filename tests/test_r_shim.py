"""The R `.Call` shim (r/src/shim.c) type-checks against the C ABI header. R itself is absent from this image, so R's API
is declared by a mock header; what is verified is that the shim calls nnlm_nnmf / nnlm_nnlm with the argument list
include/nnlm_b200.h declares and registers the reference's two entry points with the reference's arities
(src/RcppExports.cpp:56-65)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_compiles_against_header():
    cc = shutil.which("gcc") or "/usr/bin/gcc"
    if not os.path.exists(cc):
        pytest.skip("no C compiler")
    cmd = [cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-fsyntax-only", "-I", os.path.join(ROOT, "r", "tests", "mock"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "src", "shim.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_registers_reference_entry_points():
    src = open(os.path.join(ROOT, "r", "src", "shim.c")).read()
    assert re.search(r'\{"_NNLM_c_nnlm",\s*\(DL_FUNC\)&_NNLM_c_nnlm,\s*9\}', src)
    assert re.search(r'\{"_NNLM_c_nnmf",\s*\(DL_FUNC\)&_NNLM_c_nnmf,\s*17\}', src)
    assert "R_init_NNLM" in src and "Target tolerance not reached. Try a larger max.iter." in src


# ---- executed end to end (GPU box): shim.c + the miniature runtime r/tests/mock/rmock.c + libnnlm_b200.so -----------------
import struct
import sys

import numpy as np


def _build_driver(tmp_path):
    cc = shutil.which("gcc") or "/usr/bin/gcc"
    exe = str(tmp_path / "shim_driver")
    libdir = os.path.join(ROOT, "nnlm_b200")
    cmd = [cc, "-std=c11", "-O1", "-Wall", "-Wextra", "-Wno-cast-function-type", "-I", os.path.join(ROOT, "r", "tests", "mock"),
           "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "r", "tests", "shim_driver.c"),
           os.path.join(ROOT, "r", "tests", "mock", "rmock.c"), os.path.join(ROOT, "r", "src", "shim.c"),
           "-L", libdir, "-lnnlm_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_shim_driver_links_against_the_library(tmp_path):
    """CPU: the shim, the mock runtime and the driver link against libnnlm_b200.so; without a GPU the call surfaces the
    library's NNLM_E_NO_DEVICE message through Rf_error (no CPU fallback)."""
    exe = _build_driver(tmp_path)
    sys.path.insert(0, ROOT)
    from nnlm_b200 import _capi as K
    if K.device_count()[0] > 0:
        pytest.skip("a CUDA device is present: covered by the gpu test")
    r = subprocess.run([exe, "interrupt"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_shim_executes_nnmf_and_nnlm_like_the_reference_call(tmp_path):
    """`.Call(_NNLM_c_nnmf, 17 args)` / `.Call(_NNLM_c_nnlm, 9 args)` through shim.c give the results of the ctypes path
    (same C ABI), with the reference's list layout (src/nnmf.cpp:211-219, src/nnlm.cpp:49-52) checked by the driver."""
    sys.path.insert(0, ROOT)
    import nnlm_b200
    import oracle
    from nnlm_b200 import _capi as K
    from conftest import umat
    exe = _build_driver(tmp_path)
    n, m, k, T, trace = 300, 120, 5, 8, 2
    A = oracle.synth_matrix(n, m, k)
    W0 = 0.01 * umat(11, n, k); H0 = 0.01 * umat(12, k, m)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<9q", n, m, k, T, trace, 1, 50, 1, 0)); f.write(struct.pack("<d", -1.0))
        f.write(A.tobytes(order="F")); f.write(W0.tobytes(order="F")); f.write(H0.tobytes(order="F"))
    env = dict(os.environ, NNLM_B200_PRECISION="exact")
    r = subprocess.run([exe, "nnmf", fin, fout], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert "Target tolerance not reached. Try a larger max.iter." in r.stderr          # rel.tol = -1 never converges
    raw = open(fout, "rb").read()
    ne, nit, warned = struct.unpack_from("<3q", raw, 0)
    off = 24
    W = np.frombuffer(raw, dtype="<f8", count=n * k, offset=off).reshape((n, k), order="F"); off += 8 * n * k
    H = np.frombuffer(raw, dtype="<f8", count=k * m, offset=off).reshape((k, m), order="F"); off += 8 * k * m
    vecs = [np.frombuffer(raw, dtype="<f8", count=ne, offset=off + 8 * ne * i) for i in range(4)]
    got = nnlm_b200.nnmf(A, k, init={"W": W0, "H": H0}, max_iter=T, rel_tol=-1, trace=trace, show_warning=False,
                         precision=K.PREC_EXACT)
    ref = oracle.nnmf(A, k, W0, H0, max_iter=T, rel_tol=-1, n_threads=0, inner_max_iter=50, method=1, trace=trace)
    assert nit == T == got.n_iteration and warned == 1 and ne == len(got.mse) == len(ref["mse"])
    np.testing.assert_array_equal(W, got.W); np.testing.assert_array_equal(H, got.H)
    np.testing.assert_array_equal(vecs[0], got.mse); np.testing.assert_array_equal(vecs[2], got.target_loss)
    assert np.linalg.norm(W - ref["W"]) / np.linalg.norm(ref["W"]) < 1e-9
    # default init drawn by the shim from (the mock of) R's RNG: 0.01 * U(0,1), k-fastest; result is a valid factorisation
    with open(fin, "wb") as f:
        f.write(struct.pack("<9q", n, m, k, 30, 5, 1, 50, 0, 0)); f.write(struct.pack("<d", -1.0)); f.write(A.tobytes(order="F"))
    r = subprocess.run([exe, "nnmf", fin, fout], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    raw = open(fout, "rb").read()
    ne2 = struct.unpack_from("<q", raw, 0)[0]
    W = np.frombuffer(raw, dtype="<f8", count=n * k, offset=24).reshape((n, k), order="F")
    H = np.frombuffer(raw, dtype="<f8", count=k * m, offset=24 + 8 * n * k).reshape((k, m), order="F")
    mse = np.frombuffer(raw, dtype="<f8", count=ne2, offset=24 + 8 * (n * k + k * m))
    assert (W >= 0).all() and (H >= 0).all() and mse[-1] < 0.01 and np.all(np.diff(mse) <= 1e-12)
    # nnlm: the reference's golden-vector style problem through the 9-argument entry
    p, q = 5, 3
    x = umat(41, 40, p); btrue = umat(42, p, q); y = x @ btrue
    b0 = umat(43, p, q)
    with open(fin, "wb") as f:
        f.write(struct.pack("<5q", 40, p, q, 10000, 1)); f.write(struct.pack("<d", 1e-12))
        f.write(x.tobytes(order="F")); f.write(np.asfortranarray(y).tobytes(order="F")); f.write(b0.tobytes(order="F"))
    r = subprocess.run([exe, "nnlm", fin, fout], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    raw = open(fout, "rb").read()
    coef = np.frombuffer(raw, dtype="<f8", count=p * q, offset=8).reshape((p, q), order="F")
    np.testing.assert_allclose(coef, btrue, rtol=1e-6, atol=1e-9)
    # user interrupt: polled between iterations, unwound without a longjmp across device state, then Rf_onintr()
    r = subprocess.run([exe, "interrupt"], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "interrupt propagated" in r.stdout, r.stderr
