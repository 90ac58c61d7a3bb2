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
