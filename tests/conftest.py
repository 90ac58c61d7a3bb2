"""pytest configuration: markers and shared helpers.

`-m "not gpu"` runs here on CPU (oracle vs the reference's golden vectors, host logic, C-ABI symbol checks);
`-m gpu` runs on a B200 and compares the CUDA path with the oracle through the C ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def splitmix_uniform(seed: int, count: int, offset: int = 0) -> np.ndarray:
    """Counter-based U[0,1) generator shared by tests, bench and the CUDA synthetic-data kernel:
    u(seed, idx) = (splitmix64(seed*0x9E3779B97F4A7C15 + idx) >> 11) * 2^-53."""
    with np.errstate(over="ignore"):
        idx = np.arange(offset, offset + count, dtype=np.uint64)
        z = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + idx
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def umat(seed: int, rows: int, cols: int) -> np.ndarray:
    return splitmix_uniform(seed, rows * cols).reshape((rows, cols), order="F")


@pytest.fixture(scope="session")
def nsclc():
    return np.load(os.path.join(ROOT, "tests", "golden", "nsclc.npz"))["A"]
