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


def workload(small: bool):
    if small:
        return dict(n=5000, m=2000, k=50, name="synthetic dense 5000x2000 (smoke size), k=50, scd/mse")
    return dict(n=50000, m=10000, k=50, name="synthetic dense 50000x10000, k=50, method='scd', loss='mse'")


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
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
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
    args = ap.parse_args()
    wl = workload(args.small)
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
    from nnlm_b200 import _capi as K
    from nnlm_b200.session import Session, synth_init, synth_matrix

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != 1:
        raise SystemExit("the column-sharded multi-GPU path is not wired into bench.py yet")

    n, m, k = wl["n"], wl["m"], wl["k"]
    W0, H0 = synth_init(n, m, k)
    sess = Session(k=k, method=1, inner_max_iter=50, inner_rel_tol=1e-9, precision=args.precision, device=local_rank,
                   synthetic=dict(n=n, m=m), timing=True)
    sess.set_factors(W0, H0)
    W = max(args.warmup, 3)
    sess.run(W)                                       # warm-up iterations (also moves past the cold first sweeps)
    sess.reset_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_ms, sweeps = sess.run(args.steps)             # EXACTLY K iterations, CUDA events on the library's stream
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    st = sess.stats()
    mse, _, _ = sess.error()
    value = args.steps / (dev_ms * 1e-3)

    s_bytes = 8 if st["precision_used"] == K.PREC_EXACT else 4
    hbm_peak, peak_src = peaks()
    cross_ms = st["cross_ms"] / max(st["cross_launches"], 1)
    algo_bytes = float(n) * m * s_bytes + 8.0 * k * (n + m)           # one pass over the A copy + factor in + partials out
    achieved = algo_bytes / (cross_ms * 1e-3) / 1e9 if cross_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "kernel": "cross-product (one pass over A per half-iteration)",
                "algorithmic_bytes_per_launch": algo_bytes, "ms_per_launch": cross_ms, "peak_source": peak_src,
                "share_of_step": {"cross": st["cross_ms"] / dev_ms, "solve": st["solve_ms"] / dev_ms,
                                  "gram": st["gram_ms"] / dev_ms}}
    sess.close()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 solver state, " + ("f64" if s_bytes == 8 else "f32") + " A",
            "data": "synthetic",
            "config": {"workload": wl["name"], "inner_max_iter": 50, "inner_rel_tol": 1e-9,
                       "l2_policy": "inputs larger than L2 (each half streams a %.1f GB copy of A)" % (n * m * s_bytes / 1e9),
                       "avg_inner_sweeps_per_column": sweeps / (args.steps * (n + m)), "mse_after": mse},
            "clocks": clocks, "gpu_launches": int(st["launches"]), "roofline": roofline}

    if not args.no_e2e:
        A = synth_matrix(n, m, k)
        Ap = torch.from_numpy(A).pin_memory().numpy()   # pinned host copy (F-order is preserved through the transpose view)
        Ap = np.asfortranarray(Ap) if not Ap.flags.f_contiguous else Ap
        del A
        Wp = torch.from_numpy(np.ascontiguousarray(W0.T)).pin_memory().numpy().T
        Hp = torch.from_numpy(np.ascontiguousarray(H0.T)).pin_memory().numpy().T
        t0 = time.perf_counter()
        r = nnlm_b200.nnmf(Ap, k, init={"W": Wp, "H": Hp}, max_iter=args.steps, rel_tol=-1, trace=0, verbose=0,
                           show_warning=False, inner_max_iter=50, precision=args.precision, device=local_rank)
        wall = time.perf_counter() - t0
        line["e2e"] = {"value": args.steps / wall, "unit": UNIT,
                       "h2d_bytes_per_step": r.stats["h2d_bytes"] / args.steps,
                       "d2h_bytes_per_step": r.stats["d2h_bytes"] / args.steps,
                       "seconds": wall, "upload_ms": r.stats["upload_ms"], "loop_ms": r.stats["loop_ms"],
                       "download_ms": r.stats["download_ms"],
                       "what": "nnmf(A, k, init, max.iter=steps, rel.tol=-1, trace=0) with pinned host A/W/H"}
        del Ap
    if not args.no_cpu and world == 1:
        sec, threads, sample = cpu_iteration(wl, steps=1)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
