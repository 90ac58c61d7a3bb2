"""CPU (gloo, world_size 2) coverage of the host side of the sharded path: id exchange, shard bounds, slice assembly."""
import os
import subprocess
import sys

from nnlm_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_match_engine_rule():
    assert [shard.shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 3), (9, 1)]
    assert shard.shard_bounds(50000, 8, 7) == (43750, 6250)
    assert shard.shard_bounds(3, 4, 3) == (3, 0)


def test_world_size_2_gloo():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "_gloo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
