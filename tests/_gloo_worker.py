"""Worker of tests/test_shard_gloo.py: world_size-2 host-side plumbing of the sharded path on CPU (gloo)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from nnlm_b200 import shard


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # (1) the NCCL unique id created on rank 0 reaches every rank unchanged
    uid = shard.exchange_id(rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, uid)
    assert len(uid) == 128 and all(g == gathered[0] for g in gathered) and any(b != 0 for b in uid)
    # (2) the shard bounds of all ranks tile [0, total) exactly, in rank order, with equal chunk capacity
    for total in (2, 3, 10, 517, 1003, 10000, 50000):
        mine = shard.shard_bounds(total, world, rank)
        allb = [None] * world
        dist.all_gather_object(allb, mine)
        pos = 0
        chunk = -(-total // world)
        for (start, count) in allb:
            assert start == pos and 0 <= count <= chunk
            pos += count
        assert pos == total
    # (3) a factor assembled from per-rank slices (what ncclAllGather does in place) equals the whole factor
    k, total = 3, 11
    full = np.arange(k * total, dtype=np.float64).reshape((k, total), order="F")
    start, count = shard.shard_bounds(total, world, rank)
    chunk = -(-total // world)
    send = np.zeros((k, chunk), order="F"); send[:, :count] = full[:, start:start + count]
    parts = [torch.zeros(k * chunk, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(send.ravel(order="F").copy()))
    got = np.concatenate([p.numpy().reshape((k, chunk), order="F") for p in parts], axis=1)[:, :total]
    assert np.array_equal(got, full)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
