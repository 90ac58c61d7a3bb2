#!/usr/bin/env python
"""bench.py — ANLS iterations/sec on the synthetic 50000 x 10000 dense problem, k = 50, scd/mse (BASELINE.json configs[1]).

A "step" is one ANLS iteration (W-half then H-half, src/nnmf.cpp:109-133) on the resident factors.
  value      whole-job iterations/sec with A already resident in HBM (device time, CUDA events, max over ranks)
  e2e        the same metric through the public nnmf() call with HOST buffers: pinned A/W/H are copied to the device,
             `steps` iterations run (rel.tol = -1, trace = 0 -> the reference's two error evaluations, first/last),
             W and H are copied back — wall clock around the call; value = steps / seconds
  roofline   the cross-product kernel (the one pass over A per half-iteration): algorithmic bytes n*m*s per launch
             divided by its mean launch duration from CUDA events on the library's stream, vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (C++/OpenMP restatement of the reference, all host cores) on ONE full iteration of the
             same problem, including the per-iteration A.t() copy the reference makes (src/nnmf.cpp:117,131)
`--impl reference` times that CPU path alone and prints the same JSON line with "impl": "reference".
Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--small]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ANLS iters/sec on 50k x 10k dense, k=50, MSE loss"
UNIT = "iters/s"


def workload(small: bool, config: int = 2):
    """BASELINE.json configs: 2 is the headline (the metric is quoted on it); 3, 4, 5 are the other GPU configurations."""
    if small:
        return dict(n=5000, m=2000, k=50, method=1, inner=50, na=0.0, name="synthetic dense 5000x2000 (smoke size), k=50, scd/mse")
    if config == 3:
        return dict(n=50000, m=10000, k=50, method=4, inner=1, na=0.0,
                    name="synthetic dense 50000x10000, k=50, method='lee', loss='mkl' (inner.max.iter=1)")
    if config == 4:
        return dict(n=50000, m=10000, k=50, method=1, inner=50, na=0.2,
                    name="synthetic 50000x10000 with 20% NA, k=50, method='scd' (update_with_missing path)")
    if config == 5:
        return dict(n=200000, m=20000, k=128, method=1, inner=50, na=0.0,
                    name="synthetic dense 200000x20000, k=128, method='scd', loss='mse'")
    return dict(n=50000, m=10000, k=50, method=1, inner=50, na=0.0,
                name="synthetic dense 50000x10000, k=50, method='scd', loss='mse'")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            inside = t0 <= t <= t1 + 0.2
            try:
                if inside:
                    sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            if inside:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:      # timed region shorter than one sampling period: use every sample we have
            for t, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_iteration(wl, steps=1, warm=0):
    """The reference's CPU path for one ANLS iteration on the host cores: A.t() copy + update(W) + update(H)
    (src/nnmf.cpp:117-119). Returns (seconds per iteration, threads, description)."""
    import oracle
    from nnlm_b200.session import synth_init, synth_matrix
    n, m, k = wl["n"], wl["m"], wl["k"]
    try:
        A = synth_matrix(n, m, k)                 # device-generated, identical to the GPU arm's matrix
        src = "same matrix as the GPU arm"
    except Exception:                             # no GPU (reference arm on a CPU-only box): host generator
        from nnlm_b200.session import splitmix_uniform as u
        A = np.asfortranarray(u(1, n * k).reshape((n, k), order="F") @ u(2, k * m).reshape((k, m), order="F"))
        A += 0.1 * np.random.default_rng(3).random((n, m))
        src = "host-generated matrix of the same recipe"
    W0, H0 = synth_init(n, m, k)
    Wt = np.asfortranarray(W0.T)
    H = H0
    threads = oracle.max_threads()
    times = []
    for it in range(warm + steps):
        t0 = time.perf_counter()
        At = oracle.transpose(A)                                                     # A.t(), src/nnmf.cpp:117
        Wt, _ = oracle.update(Wt, H, At, n_threads=0, method=1, max_iter=50, rel_tol=1e-9, with_missing=0)
        del At
        H, _ = oracle.update(H, Wt, A, n_threads=0, method=1, max_iter=50, rel_tol=1e-9, with_missing=0)
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
    return float(np.mean(times)), threads, f"{steps} full ANLS iteration(s) of {wl['name']} ({src}), incl. the A.t() copy"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="smoke-size problem (not a bench line)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--precision", type=int, default=2, help="1 exact (fp64 A), 2 fast")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configuration (default: the headline)")
    args = ap.parse_args()
    wl = workload(args.small, args.config)
    if args.config != 2:
        args.no_cpu = True          # the CPU arm is defined on the headline configuration
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 2))          # bounded sample: each iteration costs seconds on the CPU
        sec, threads, sample = cpu_iteration(wl, steps=steps, warm=0)
        v = 1.0 / sec
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": 0, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["name"]},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import nnlm_b200
    from nnlm_b200 import _capi as K, shard
    from nnlm_b200.session import Session, synth_block, synth_init, synth_matrix

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    comm = None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = shard.comm_from_torch(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n, m, k = wl["n"], wl["m"], wl["k"]
    W0, H0 = synth_init(n, m, k)
    # strong scaling: the same 50000 x 10000 problem, A column-sharded (H-half) and row-sharded (W-half) over the ranks
    sess = Session(k=k, method=wl["method"], inner_max_iter=wl["inner"], inner_rel_tol=1e-9, precision=args.precision,
                   device=local_rank, synthetic=dict(n=n, m=m, na_frac=wl["na"]), timing=True, comm=comm)
    sess.set_factors(W0, H0)
    W = max(args.warmup, 3)
    sess.run(W)                                       # warm-up iterations (also moves past the cold first sweeps)
    sess.reset_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    t0 = time.perf_counter()
    dev_ms, sweeps = sess.run(args.steps)             # EXACTLY K iterations, CUDA events on the library's stream
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    dev_ms = max_over_ranks(dev_ms)                   # device time, max over ranks
    st = sess.stats()
    mse, _, _ = sess.error()
    value = args.steps / (dev_ms * 1e-3)

    s_bytes = 8 if st["precision_used"] == K.PREC_EXACT else 4
    hbm_peak, peak_src = peaks()
    cross_ms = st["cross_ms"] / max(st["cross_launches"], 1)
    # one pass over this rank's copy of A for the half (n*m/world elements) + factor in + partials out
    algo_bytes = float(n) * m * s_bytes / world + 8.0 * k * (n + m)
    achieved = algo_bytes / (cross_ms * 1e-3) / 1e9 if cross_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": 2.10e9 if (world == 1 and not args.small and args.config == 2) else None,
                "kernel": "k_cross_tc (cross-product: one pass over A per half-iteration), per GPU",
                "algorithmic_bytes_per_launch": algo_bytes, "ms_per_launch": cross_ms, "peak_source": peak_src,
                "traffic_source": ("ncu --set full dram__bytes_read+write per launch (W-half 2.05e9, H-half 2.15e9), "
                                  "profiles/r1_n_final_build.md") if (world == 1 and not args.small and args.config == 2) else None,
                "share_of_step": {"cross": st["cross_ms"] / dev_ms, "solve": st["solve_ms"] / dev_ms,
                                  "gram": st["gram_ms"] / dev_ms, "comm": st["comm_ms"] / dev_ms}}
    # second roofline, for the kernel with the largest share of the step: the SCD solve is bound by the fp64 pipe (DFMA and
    # DMMA share it: 64 FMA/clk/SM, measured 63.8 with scratch/dmma_bench.cu). Algorithmic work = k*k FMA per column sweep.
    if wl["method"] == 1 and wl["na"] == 0.0 and st["solve_ms"] > 0:
        sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
        fp64_peak = 64.0 * 148 * sm_mhz * 1e6 / 1e12                       # TFMA/s per GPU
        solve_tfma = float(sweeps) * k * k / world / (st["solve_ms"] * 1e-3) / 1e12
        roofline["solve"] = {"bound": "fp64 pipe", "achieved": solve_tfma, "peak": fp64_peak, "unit": "TFMA/s",
                             "frac": solve_tfma / fp64_peak, "kernel": "k_scd_chain (both halves), per GPU",
                             "peak_source": "64 FMA/clk/SM x 148 SMs x sampled SM clock (DMMA.8x8x4 measured at 63.8, profiles/r1_m_scd_stalls.md)"}
    launches = int(sum_over_ranks(float(st["launches"])))
    sess.close()

    metric = METRIC if args.config == 2 else f"ANLS iters/sec, BASELINE config {args.config}"
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 solver state, " + ("f64 A" if s_bytes == 8 else "fp16 hi+lo planes of A (4 B/element), fp32 TMEM accumulate drained to f64"),
            "data": "synthetic",
            "config": {"workload": wl["name"], "inner_max_iter": wl["inner"], "inner_rel_tol": 1e-9,
                       "sharding": "none" if world == 1 else f"columns (H-half) and rows (W-half) over {world} ranks; "
                                   "k x k Gram all-reduce + factor all-gather per half-iteration (NCCL)",
                       "l2_policy": "inputs larger than L2 (each half streams a %.2f GB copy of A per GPU)" % (n * m * s_bytes / world / 1e9),
                       "avg_inner_sweeps_per_column": sweeps / (args.steps * (n + m)), "mse_after": mse},
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline}

    if not args.no_e2e and args.config == 2:
        # end to end through the public API with HOST buffers: pinned A (whole matrix on 1 GPU, this rank's shards otherwise)
        # and pinned factors are copied to the device, `steps` iterations run, W and H are copied back
        pin = lambda a: np.asfortranarray(torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory().numpy().T)
        Wp, Hp = pin(W0), pin(H0)
        if world == 1:
            Ap = pin(synth_matrix(n, m, k))
            barrier()
            t0 = time.perf_counter()
            r = nnlm_b200.nnmf(Ap, k, init={"W": Wp, "H": Hp}, max_iter=args.steps, rel_tol=-1, trace=0, verbose=0,
                               show_warning=False, inner_max_iter=50, precision=args.precision, device=local_rank,
                               check_k=False)
            wall = time.perf_counter() - t0
            h2d, d2h = r.stats["h2d_bytes"], r.stats["d2h_bytes"]
            what = "nnmf(A, k, init, max.iter=steps, rel.tol=-1, trace=0) with pinned host A/W/H"
            extra = {"upload_ms": r.stats["upload_ms"], "loop_ms": r.stats["loop_ms"], "download_ms": r.stats["download_ms"],
                     "host_setup_ms": r.stats["host_setup_ms"], "host_loop_ms": r.stats["host_loop_ms"],
                     "host_finish_ms": r.stats["host_finish_ms"], "python_run_time_s": r.run_time}
            del Ap
        else:
            r0, nr = shard.shard_bounds(n, world, rank); c0, mc = shard.shard_bounds(m, world, rank)
            Acol = pin(synth_block(n, 0, n, c0, mc, k)); Arow = pin(synth_block(n, r0, nr, 0, m, k))
            barrier()
            t0 = time.perf_counter()
            s2 = Session(k=k, method=1, inner_max_iter=50, inner_rel_tol=1e-9, precision=args.precision, device=local_rank,
                         comm=comm, shards=(Acol, Arow), shape=(n, m))
            s2.set_factors(Wp, Hp)
            s2.run(args.steps)
            s2.get_factors()
            barrier()
            wall = time.perf_counter() - t0
            st2 = s2.stats()
            h2d, d2h = sum_over_ranks(float(st2["h2d_bytes"])), sum_over_ranks(float(st2["d2h_bytes"]))
            what = "sharded Session(shards=pinned host A[:,cols_g], A[rows_g,:]) + set_factors + run(steps) + get_factors, all ranks"
            extra = {"upload_ms": st2["upload_ms"]}
            s2.close()
            del Acol, Arow
        wall = max_over_ranks(wall)
        line["e2e"] = {"value": args.steps / wall, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
                       "d2h_bytes_per_step": d2h / args.steps, "seconds": wall, "what": what, **extra}
    if not args.no_cpu and world == 1:
        sec, threads, sample = cpu_iteration(wl, steps=1)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line))
    if comm is not None:
        comm.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
