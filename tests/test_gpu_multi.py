"""Multi-GPU sharded path (needs >= 2 GPUs on the box; skipped otherwise): launches tests/multi_gpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

from nnlm_b200 import _capi as K

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_matches_oracle_and_single_gpu():
    n_dev, _ = K.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n_dev < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29511", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
