"""CPU-side checks of the drop-in boundary: the built library loads and exports every symbol include/nnlm_b200.h
declares, reports the ABI version, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nnlm_b200
from nnlm_b200 import _capi as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "nnlm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nnlm_[a-z_0-9]+)\s*\(", hdr)) - {"nnlm_interrupt_fn"})


def test_library_exports_every_declared_symbol():
    lib = K.lib()
    names = _declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(K.SYMBOLS) == names


def test_abi_version_matches_header():
    hdr = open(os.path.join(ROOT, "include", "nnlm_b200.h")).read()
    ver = int(re.search(r"#define NNLM_B200_ABI_VERSION (\d+)", hdr).group(1))
    assert K.lib().nnlm_abi_version() == ver


def test_struct_layouts_match_header():
    """The ctypes mirrors have the sizes the compiled library reports for its own structs (nnlm_sizeof)."""
    lib = K.lib()
    lib.nnlm_sizeof.restype = C.c_size_t
    assert C.sizeof(K.Options) == lib.nnlm_sizeof(0) == 48
    assert C.sizeof(K.Stats) == lib.nnlm_sizeof(1) == 176


def test_no_cpu_fallback():
    n_dev, _ = K.device_count()
    if n_dev > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(K.NnlmError) as ei:
        nnlm_b200.nnmf(np.random.default_rng(0).random((20, 10)), 2, max_iter=2)
    assert ei.value.code == K.E_NO_DEVICE
    with pytest.raises(K.NnlmError):
        nnlm_b200.nnlm(np.random.default_rng(0).random((20, 3)), np.ones(20), check_x=False)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "nnlm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "nnlm_oracle" not in txt, f
