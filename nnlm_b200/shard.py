"""Column/row sharding of the ANLS path across the GPUs of one box: one process per GPU, NCCL inside the C library.

Rank g of R owns the columns [c0, c0+mc) of A (H-half) and the rows [r0, r0+nr) of A (W-half); see csrc/engine.cuh.
The host program only has to (1) compute the shard bounds, (2) move the 128-byte NCCL unique id from rank 0 to the
other ranks — any transport will do; `comm_from_torch` uses torch.distributed (gloo or nccl) — and (3) hand the
communicator to the session.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as K


def shard_bounds(total: int, world: int, rank: int):
    """(start, count) of rank's contiguous slice: equal chunks of ceil(total/world), the tail rank(s) get the remainder.
    Mirrors Engine::Engine in csrc/engine.cu (ncclAllGather needs equal counts, so buffers are padded to chunk*world)."""
    chunk = -(-total // world)
    start = min(total, rank * chunk)
    return start, min(total, start + chunk) - start


def unique_id() -> bytes:
    buf = (C.c_ubyte * K.COMM_ID_BYTES)()
    err = C.create_string_buffer(512)
    K.check(K.lib().nnlm_comm_unique_id(buf, err, C.c_size_t(512)), err)
    return bytes(buf)


class Comm:
    """Owns one nnlm_comm (NCCL communicator of this rank)."""

    def __init__(self, uid: bytes, rank: int, world: int, device: int):
        assert len(uid) == K.COMM_ID_BYTES
        self.rank, self.world, self.device = rank, world, device
        self._h = C.c_void_p()
        buf = (C.c_ubyte * K.COMM_ID_BYTES).from_buffer_copy(uid)
        err = C.create_string_buffer(512)
        K.check(K.lib().nnlm_comm_init(C.byref(self._h), buf, C.c_int32(rank), C.c_int32(world), C.c_int32(device), err,
                                       C.c_size_t(512)), err)

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            K.lib().nnlm_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def exchange_id(rank: int, make_id=unique_id, device=None) -> bytes:
    """Rank 0 creates the id, torch.distributed broadcasts it (works on gloo with CPU tensors and on nccl with CUDA ones)."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(K.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def comm_from_torch(device_index: int) -> Comm:
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", device_index) if dist.get_backend() == "nccl" else None
    return Comm(exchange_id(rank, device=dev), rank, world, device_index)
