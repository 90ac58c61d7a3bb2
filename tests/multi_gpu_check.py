"""Sharded-path check, run under torchrun on >= 2 GPUs (tests/test_gpu_multi.py launches it; also usable by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Every rank builds its column and row shards of the same synthetic matrix, runs T ANLS iterations with the NCCL exchanges
(k x k Gram all-reduce + factor all-gather per half-iteration) and compares the factors with
  (a) the CPU oracle on the whole matrix (small sizes), tolerance 1e-5 relative Frobenius (north star),
  (b) the single-GPU session of the same library (rank 0), tolerance 1e-6 (SURVEY.md §8e "Determinism").
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import nnlm_b200
from nnlm_b200 import _capi as K, shard
from nnlm_b200.session import Session, synth_block, synth_init, synth_matrix


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = shard.comm_from_torch(local)
    ok = True
    cases = [(1003, 517, 6, 1, K.PREC_EXACT, 6, "synthetic"), (1003, 517, 6, 1, K.PREC_FAST, 6, "synthetic"),
             (640, 389, 5, 2, K.PREC_EXACT, 4, "host"), (900, 400, 4, 4, K.PREC_FAST, 3, "host"),
             (700, 300, 4, 1, K.PREC_EXACT, 4, "host-missing")]
    for (n, m, k, method, prec, T, src) in cases:
        inner = 50 if method < 3 else 1
        W0, H0 = synth_init(n, m, k)
        na = 0.15 if src == "host-missing" else 0.0
        if src == "synthetic":
            s = Session(k=k, method=method, inner_max_iter=inner, precision=prec, device=local, comm=comm,
                        synthetic=dict(n=n, m=m, na_frac=na))
        else:
            r0, nr = shard.shard_bounds(n, world, rank); c0, mc = shard.shard_bounds(m, world, rank)
            Acol = synth_block(n, 0, n, c0, mc, k, na_frac=na)
            Arow = synth_block(n, r0, nr, 0, m, k, na_frac=na)
            s = Session(k=k, method=method, inner_max_iter=inner, precision=prec, device=local, comm=comm,
                        shards=(Acol, Arow), shape=(n, m))
        s.set_factors(W0, H0)
        ms, sweeps = s.run(T)
        W, H = s.get_factors()
        mse, mkl, _ = s.error()
        st = s.stats()
        s.close()
        if rank == 0:
            import oracle
            A = synth_matrix(n, m, k, na_frac=na)
            ref = oracle.nnmf(A, k, W0, H0, max_iter=T, rel_tol=-1, inner_max_iter=inner, method=method, trace=999999, n_threads=0)
            with Session(A, k=k, method=method, inner_max_iter=inner, precision=prec, device=local) as s1:
                s1.set_factors(W0, H0)
                _, sweeps1 = s1.run(T)
                W1, H1 = s1.get_factors()
                mse1, mkl1, _ = s1.error()
            e_or = max(rel(W, ref["W"]), rel(H, ref["H"]))
            e_1g = max(rel(W, W1), rel(H, H1))
            good = e_or < 1e-5 and e_1g < 1e-6 and abs(mse - mse1) <= 1e-7 * abs(mse1) and sweeps == sweeps1 and st["comm_bytes"] > 0
            ok = ok and good
            print(f"[{world} ranks] {n}x{m} k={k} method={method} prec={prec} {src}: vs oracle {e_or:.2e}, vs 1 GPU {e_1g:.2e}, "
                  f"mse {mse:.6e} / {mse1:.6e}, sweeps {sweeps}/{sweeps1}, comm {st['comm_bytes']} B -> {'ok' if good else 'FAIL'}", flush=True)
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
