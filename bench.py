#!/usr/bin/env python
"""bench.py — ANLS iterations/sec on the synthetic 50000 x 10000 dense problem, k = 50, scd/mse (BASELINE.json configs[1]).

A "step" is one ANLS iteration (W-half then H-half, src/nnmf.cpp:109-133) on the resident factors.
  value      whole-job iterations/sec with A already resident in HBM (device time, CUDA events, max over ranks)
  e2e        the same metric through the public nnmf() call with HOST buffers: pinned A/W/H are copied to the device,
             `steps` iterations run (rel.tol = -1, trace = 0 -> the reference's two error evaluations, first/last),
             W and H are copied back — wall clock around the call; value = steps / seconds (median of 3 calls; the
             individual calls and a 200-step call are listed beside it)
  roofline   the DOMINANT kernel of the step (config 2/5: the SCD solve against the fp64 pipe), with the cross-product
             kernel against the measured HBM bandwidth and the whole-step HBM fraction beside it
  parity     T = 1 from the BASELINE init on the benchmarked matrix and kernels, against the oracle's iteration of the
             cpu_baseline leg (N = 1), or against a single-GPU session of the same library (N > 1)
  cpu_baseline  the oracle (C++/OpenMP restatement of the reference, all host cores) on a bounded sample of the same
             problem, including the per-iteration A.t() copy the reference makes (src/nnmf.cpp:117,131) ("faithful");
             the variant without the per-iteration copy ("fair") is listed beside it
`--impl reference` times that CPU path alone (no GPU library is loaded in that process) and prints the same JSON line
with "impl": "reference".
Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--small]
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
CPU_BUDGET_S = 120.0      # the reference arm's timed loop is capped to about this much CPU time


def workload(small: bool, config: int = 2):
    """BASELINE.json configs: 2 is the headline (the metric is quoted on it); 3, 4, 5 are the other GPU configurations.
    cpu_frac: the fraction of the columns (H-half) / rows (W-half) the CPU leg computes; the rest is extrapolated."""
    if small:
        base = {2: dict(n=5000, m=2000, k=50, method=1, inner=50, na=0.0), 3: dict(n=5000, m=2000, k=50, method=4, inner=1, na=0.0),
                4: dict(n=5000, m=2000, k=50, method=1, inner=50, na=0.2), 5: dict(n=8000, m=4000, k=128, method=1, inner=50, na=0.0)}[config]
        base.update(cpu_frac=1.0, name=f"synthetic {base['n']}x{base['m']} (smoke size of config {config}), k={base['k']}, method code {base['method']}"
                                       + (f", {int(base['na'] * 100)}% NA" if base["na"] else ""))
        return base
    if config == 3:
        # (a full oracle iteration takes ~15 s on 16 cores: affordable, and it carries the T = 1 parity check at the full size)
        return dict(n=50000, m=10000, k=50, method=4, inner=1, na=0.0, cpu_frac=1.0,
                    name="synthetic dense 50000x10000, k=50, method='lee', loss='mkl' (inner.max.iter=1)")
    if config == 4:
        return dict(n=50000, m=10000, k=50, method=1, inner=50, na=0.2, cpu_frac=1.0,
                    name="synthetic 50000x10000 with 20% NA, k=50, method='scd' (update_with_missing path)")
    if config == 5:
        return dict(n=200000, m=20000, k=128, method=1, inner=50, na=0.0, cpu_frac=1.0 / 32,
                    name="synthetic dense 200000x20000, k=128, method='scd', loss='mse'")
    return dict(n=50000, m=10000, k=50, method=1, inner=50, na=0.0, cpu_frac=1.0,
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


# ------------------------------------------------------------------------------------------------ CPU arm (oracle only)
def cpu_arm(wl, steps=1, warm=0, want_factors=False):
    """The reference's CPU path on all host cores, on the matrix of the GPU arm (host generator, bit-identical).

    cpu_frac == 1: `warm + steps` full ANLS iterations (A.t() copy + update(W) + update(H), src/nnmf.cpp:117-119), from the
    BASELINE init; the iterations continue the trajectory like the reference's loop does.
    cpu_frac < 1 (configs 3-5, seconds to minutes per iteration): one iteration on a sample — the H-half over the first
    frac*m columns (whole W), the W-half over the first frac*n rows (whole H) — and the per-column cost extrapolated:
    t_half = t_fixed + (t_sample - t_fixed) / frac, t_fixed = the same call on one column (the Gram / row sums that
    do not scale with the column count). Returns a dict."""
    import oracle
    cores = oracle.host_cores()
    oracle.set_threads(cores)                # torchrun exports OMP_NUM_THREADS=1: use the cores the process may run on
    n, m, k, method, inner, na, frac = wl["n"], wl["m"], wl["k"], wl["method"], wl["inner"], wl["na"], wl["cpu_frac"]
    W0 = 0.01 * oracle.splitmix_uniform(11, n * k).reshape((n, k), order="F")
    H0 = 0.01 * oracle.splitmix_uniform(12, k * m).reshape((k, m), order="F")
    Wt = np.asfortranarray(W0.T)
    H = H0.copy(order="F")
    miss = 1 if na > 0 else 0
    kw = dict(n_threads=0, method=method, max_iter=inner, rel_tol=1e-9, with_missing=miss)
    out = {"cores": cores}
    if frac >= 1.0:
        A = oracle.synth_matrix(n, m, k, na_frac=na)
        faithful, first = [], None
        for it in range(warm + steps):
            t0 = time.perf_counter()
            At = oracle.transpose(A)                                                     # A.t(), src/nnmf.cpp:117
            Wt, _ = oracle.update(Wt, H, At, **kw)
            del At
            H, _ = oracle.update(H, Wt, A, **kw)
            dt = time.perf_counter() - t0
            if it == 0 and want_factors:
                first = (np.asfortranarray(Wt.T), H.copy(order="F"))
            if it >= warm:
                faithful.append(dt)
            if it >= warm and sum(faithful) > CPU_BUDGET_S:
                break
        # "fair": the transposed copy is made once, outside the loop (one more iteration of the same trajectory)
        At = oracle.transpose(A)
        t0 = time.perf_counter()
        Wt, _ = oracle.update(Wt, H, At, **kw)
        H, _ = oracle.update(H, Wt, A, **kw)
        fair = time.perf_counter() - t0
        del At, A
        out.update(sec=float(np.mean(faithful)), steps=len(faithful), fair_sec=fair, first=first,
                   sample=f"{len(faithful)} full ANLS iteration(s) of {wl['name']} after {warm} warm-up iteration(s), same matrix as "
                          f"the GPU arm (host generator), incl. the per-iteration A.t() copy")
        return out
    ms, ns = max(int(round(m * frac)), 1), max(int(round(n * frac)), 1)
    Acol = oracle.synth_block(n, 0, n, 0, ms, k, na_frac=na)               # A[:, :ms]
    Arow = oracle.synth_block(n, 0, ns, 0, m, k, na_frac=na)               # A[:ns, :]
    t0 = time.perf_counter(); ArowT = oracle.transpose(Arow); t_tr = time.perf_counter() - t0
    t0 = time.perf_counter(); oracle.update(Wt[:, :1], H, ArowT[:, :1], **kw); tw_fixed = time.perf_counter() - t0
    t0 = time.perf_counter(); Wts, _ = oracle.update(Wt[:, :ns], H, ArowT, **kw); tw = time.perf_counter() - t0
    # the H-half of the reference sees the new W; the sample only has ns new rows: use them with the old rest (timing only)
    Wt2 = Wt.copy(order="F"); Wt2[:, :ns] = Wts
    t0 = time.perf_counter(); oracle.update(H[:, :1], Wt2, Acol[:, :1], **kw); th_fixed = time.perf_counter() - t0
    t0 = time.perf_counter(); oracle.update(H[:, :ms], Wt2, Acol, **kw); th = time.perf_counter() - t0
    w_full = tw_fixed + max(tw - tw_fixed, 0.0) * (n / ns)
    h_full = th_fixed + max(th - th_fixed, 0.0) * (m / ms)
    tr_full = t_tr * (n / ns)
    out.update(sec=w_full + h_full + tr_full, steps=1, fair_sec=w_full + h_full, first=None,
               sample=f"EXTRAPOLATED from a sample of {wl['name']}: W-half over the first {ns} of {n} rows ({tw:.2f} s), H-half over "
                      f"the first {ms} of {m} columns ({th:.2f} s), A.t() of the sampled rows ({t_tr:.2f} s); per-call fixed cost "
                      f"(Gram / row sums, {tw_fixed:.2f} + {th_fixed:.2f} s) counted once, the rest scaled by 1/{frac:.4g}")
    return out


def reference_line(args, wl, metric):
    r = cpu_arm(wl, steps=max(1, args.steps), warm=max(0, args.warmup) if wl["cpu_frac"] >= 1.0 else 0)
    v = 1.0 / r["sec"]
    return {"impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "steps_requested": args.steps, "warmup": args.warmup if wl["cpu_frac"] >= 1.0 else 0, "ms_per_step": r["sec"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "inner_max_iter": wl["inner"], "inner_rel_tol": 1e-9,
                       "step_cap": f"timed loop stops after {CPU_BUDGET_S:.0f} s of CPU work"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                             "variant": "faithful (per-iteration A.t() copy, as src/nnmf.cpp:117,131 executes)",
                             "fair_value": 1.0 / r["fair_sec"], "fair_variant": "transposed copy made once, outside the timed loop"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="smoke-size problem (not a bench line)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (and the oracle parity that rides on it)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--precision", type=int, default=2, help="1 exact (fp64 A), 2 fast")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configuration (default: the headline)")
    ap.add_argument("--shape", default="", help="n,m,k: override the problem size (experiments; not a bench line)")
    args = ap.parse_args()
    wl = workload(args.small, args.config)
    if args.shape:
        wl["n"], wl["m"], wl["k"] = (int(x) for x in args.shape.split(","))
        wl["name"] = f"custom {wl['n']}x{wl['m']}, k={wl['k']} (experiment)"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = METRIC if args.config == 2 else f"ANLS iters/sec, BASELINE config {args.config}"

    if args.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(reference_line(args, wl, metric)))
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
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # keeps NCCL's banner off stdout: one JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
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
    # strong scaling: the same problem, A column-sharded (H-half) and row-sharded (W-half) over the ranks
    sess = Session(k=k, method=wl["method"], inner_max_iter=wl["inner"], inner_rel_tol=1e-9, precision=args.precision,
                   device=local_rank, synthetic=dict(n=n, m=m, na_frac=wl["na"]), timing=True, comm=comm)
    # ---- parity leg 1: T = 1 from the BASELINE init through the benchmarked kernels; compared further down
    sess.set_factors(W0, H0)
    _, sweeps_T1 = sess.run(1)
    W_T1, H_T1 = sess.get_factors()
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
    mse, mkl, _ = sess.error()
    value = args.steps / (dev_ms * 1e-3)

    s_bytes = 8 if st["precision_used"] == K.PREC_EXACT else 4
    hbm_peak, peak_src = peaks()
    sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
    fp64_peak = 64.0 * 148 * sm_mhz * 1e6 / 1e12                       # TFMA/s per GPU (DFMA and DMMA share the pipe)
    fp64_src = "64 FMA/clk/SM x 148 SMs x sampled SM clock (DMMA.8x8x4 measured at 63.8 FMA/clk/SM, profiles/r1_m_scd_stalls.md)"
    share = {"cross": st["cross_ms"] / dev_ms, "solve": st["solve_ms"] / dev_ms, "gram": st["gram_ms"] / dev_ms,
             "comm": st["comm_ms"] / dev_ms}
    cross_ms = st["cross_ms"] / max(st["cross_launches"], 1)
    solve_ms = st["solve_ms"] / max(st["solve_launches"], 1)
    # one pass over this rank's copy of A for the half (n*m/world elements) + factor in + partials out
    algo_bytes = float(n) * m * s_bytes / world + 8.0 * k * (n + m)
    step_bytes = 2.0 * n * m * s_bytes / world + 3 * 8.0 * k * (n + m)
    cross = None
    if cross_ms > 0:
        achieved = algo_bytes / (cross_ms * 1e-3) / 1e9
        cross = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                 "kernel": "k_cross_tc (cross-product: one pass over A per half-iteration), per GPU",
                 "algorithmic_bytes_per_launch": algo_bytes, "ms_per_launch": cross_ms, "peak_source": peak_src}
    step_hbm = {"achieved": step_bytes / (dev_ms / args.steps * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "algorithmic_bytes_per_step": step_bytes,
                "what": "whole step against the HBM ceiling: one pass over A per half (2*n*m*s/N bytes) + factors, per GPU"}
    step_hbm["frac"] = step_hbm["achieved"] / hbm_peak
    if wl["method"] <= 2 and wl["na"] == 0.0:
        # dominant kernel: the SCD solve, bound by the fp64 pipe. Algorithmic work = k*k FMA per column sweep.
        fma = float(sweeps) * k * k / world
        tfma = fma / (st["solve_ms"] * 1e-3) / 1e12 if st["solve_ms"] > 0 else 0.0
        roofline = {"bound": "fp64", "achieved": tfma, "peak": fp64_peak, "unit": "TFMA/s", "frac": tfma / fp64_peak, "traffic": None,
                    "kernel": "k_scd_chain (SCD solve of both halves: the largest share of the step), per GPU",
                    "algorithmic_fma_per_launch": fma / max(st["solve_launches"], 1), "ms_per_launch": solve_ms,
                    "peak_source": fp64_src}
    elif wl["method"] >= 3:
        # KL updates (src/base_algorithms.cpp:119-151): per half n*m*k entries x inner sweeps, each ONE reciprocal on the MUFU
        # (special-function) pipe + 4 fp32 instructions; the kernel keeps wh / A on chip, so neither HBM nor L2 binds. The
        # roofline is the MUFU rate: 16 results / clk / SM.
        ent = 2.0 * n * m * k * wl["inner"] / world
        t = st["solve_ms"] / args.steps * 1e-3
        mufu_peak = 16.0 * 148 * sm_mhz * 1e6 / 1e12
        ach = ent / t / 1e12 if t > 0 else 0.0
        roofline = {"bound": "mufu", "achieved": ach, "peak": mufu_peak, "unit": "T reciprocals/s", "frac": ach / mufu_peak, "traffic": None,
                    "kernel": "k_solve_kl_fast (both halves: cluster kernel, wh in registers, A in shared memory), per GPU",
                    "entries_per_step": ent, "ms_per_launch": solve_ms,
                    "peak_source": "16 MUFU results/clk/SM x 148 SMs x sampled SM clock (B300_MICROARCH.md SFU rate; sm_100a has the 1x rate)"}
    else:
        # NA path: the per-column masked Grams as one tensor-core contraction (na_gram.cu): 2 halves x 2 n m (k(k+1)/2 + k)
        # multiply-adds x 4 fixed-point slices, fp16 operands; the kernel re-streams the fp16 mask plane once per 128 Z columns
        # and slice, so the HBM side is reported beside the tensor side.
        P = k * (k + 1) // 2 + k
        flops = 2.0 * 2.0 * n * m * P * 4 / world
        t = st["gram_ms"] / args.steps * 1e-3
        ach = flops / t / 1e12 if t > 0 else 0.0
        tpeak, tsrc = 1664.9, "fallback (B200_PROFILING.md dense bf16/fp16)"
        try:
            d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            tpeak, tsrc = float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops burst; fp16 runs at the same rate)"
        except Exception:
            pass
        launches_per_half = -(-P // 128) * 2          # 128 Z columns x 2 slices per pass over the mask plane
        mask_bytes = 2.0 * launches_per_half * n * m * 2 / world
        roofline = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak, "traffic": None,
                    "kernel": "k_mask_tc2 (tcgen05 cta_group::2) x %d launches per half (mask x Khatri-Rao slices, two slices per pass) + k_z_slices + k_fold, per GPU" % launches_per_half,
                    "algorithmic_flop_per_step": flops, "ms_per_step_in_these_kernels": t * 1e3, "peak_source": tsrc,
                    "hbm_side": {"achieved": mask_bytes / t / 1e9 if t > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                                 "bytes_per_step": mask_bytes, "what": "fp16 mask plane streamed once per launch (the slice planes of Z come from L2: as many bytes again on CTA pairs)"},
                    "solve_ms_per_step": st["solve_ms"] / args.steps}
    # DRAM traffic per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum from the `ncu --set full`
    # captures summarised in profiles/r2_traffic.json (a citation of those captures, valid for the full-size single-GPU
    # configurations they were taken on; null elsewhere)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if world == 1 and not args.small and not args.shape:
            pick = {2: ["k_scd_chain<13, 2>", "k_scd_chain<13, 1>"], 5: [], 3: ["k_solve_kl_fast<4, 10>", "k_solve_kl_fast<4, 13>"],
                    4: ["k_mask_tc2"]}[args.config]
            vals = [v for kname in pick for v in tj.get(f"config{args.config}", {}).get(kname, {}).get("dram_bytes_per_launch", [])]
            if vals:
                roofline["traffic"] = float(np.mean(vals))
                roofline["traffic_source"] = "mean over the captured launches of " + " / ".join(pick) + ", profiles/r2_traffic.json"
            if cross and args.config in (2, 4):
                cv = tj.get("config2", {}).get("k_cross_tc<64, 4, 0>", {}).get("dram_bytes_per_launch", [])
                if cv:
                    cross["traffic"] = float(np.mean(cv))
    except Exception:
        pass
    roofline["share_of_step"] = share
    if cross:
        roofline["cross"] = cross
    roofline["step_hbm"] = step_hbm
    launches = int(sum_over_ranks(float(st["launches"])))

    # ---- parity leg 2 (N > 1): the sharded T = 1 result against a single-GPU session of the same library on rank 0
    parity = {"T": 1, "init": "BASELINE (0.01*u(11), 0.01*u(12))", "tolerance": 1e-5}
    if world > 1:
        ok = 1
        if rank == 0:
            s1 = Session(k=k, method=wl["method"], inner_max_iter=wl["inner"], inner_rel_tol=1e-9, precision=args.precision,
                         device=local_rank, synthetic=dict(n=n, m=m, na_frac=wl["na"]))
            s1.set_factors(W0, H0)
            _, sw1 = s1.run(1)
            W1, H1 = s1.get_factors()
            s1.close()
            parity.update(against="single-GPU session of this library (rank 0), same matrix", rel_W=rel(W_T1, W1), rel_H=rel(H_T1, H1),
                          sweeps_equal=bool(sw1 == sweeps_T1), tolerance=5e-6,
                          tolerance_note="shards change the fp32 chunking of the tensor-core cross-product (1e-8 on W); the first "
                                         "H-half from the tiny init amplifies that 50x (dense) to 750x (20 % NA)")
            ok = int(parity["rel_W"] < 5e-6 and parity["rel_H"] < 5e-6)
            del W1, H1
        flag = torch.tensor([ok], device="cuda")
        dist.broadcast(flag, src=0)
        if int(flag.item()) != 1:
            if rank == 0:
                print(json.dumps({"error": "sharded result differs from the single-GPU result", "parity": parity}), file=sys.stderr)
            sys.exit(3)
    sess.close()

    dtype = "f64 solver state, " + ("f64 A" if s_bytes == 8 else
                                    ("fp16 hi+lo planes of A (4 B/element), fp32 TMEM accumulate drained to f64" if wl["method"] <= 2 and wl["na"] == 0.0
                                     else "fp16 hi+lo planes of A + fp16 0/1 mask plane; per-column Grams from exact fixed-point fp16 slices on tcgen05" if wl["method"] <= 2
                                     else "f32 A, f32 wh and ratios, sums / h / updates in f64"))
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": wl["name"], "inner_max_iter": wl["inner"], "inner_rel_tol": 1e-9,
                       "sharding": "none" if world == 1 else f"columns (H-half) and rows (W-half) over {world} ranks; "
                                   "k x k Gram all-reduce + factor all-gather per half-iteration (NCCL)",
                       "l2_policy": "inputs larger than L2 (each half streams a %.2f GB copy of A per GPU)" % (n * m * s_bytes / world / 1e9),
                       "avg_inner_sweeps_per_column": sweeps / (args.steps * (n + m)), "mse_after": mse, "mkl_after": mkl},
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline}

    if not args.no_e2e and args.config in (2, 3, 4):
        # end to end through the public API with HOST buffers: pinned A (whole matrix on 1 GPU, this rank's shards otherwise)
        # and pinned factors are copied to the device, `steps` iterations run, W and H are copied back
        pin = lambda a: np.asfortranarray(torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory().numpy().T)
        Wp, Hp = pin(W0), pin(H0)
        meth = "scd" if wl["method"] in (1, 3) else "lee"
        loss = "mse" if wl["method"] <= 2 else "mkl"
        if world == 1:
            Ap = pin(synth_matrix(n, m, k, na_frac=wl["na"]))

            def call(T):
                barrier()
                t0 = time.perf_counter()
                r = nnlm_b200.nnmf(Ap, k, method=meth, loss=loss, init={"W": Wp, "H": Hp}, max_iter=T, rel_tol=-1, trace=0,
                                   verbose=0, show_warning=False, inner_max_iter=wl["inner"], precision=args.precision,
                                   device=local_rank, check_k=False)
                return time.perf_counter() - t0, r
            runs = [call(args.steps) for _ in range(3)]
            walls = sorted(w for w, _ in runs)
            wall, r = walls[1], runs[0][1]
            h2d, d2h = r.stats["h2d_bytes"], r.stats["d2h_bytes"]
            what = (f"nnmf(A, k, '{meth}', '{loss}', init, max.iter=steps, rel.tol=-1, trace=0) with pinned host A/W/H; "
                    "median wall clock of 3 calls")
            long_T = 200 if args.config == 2 else 0
            extra = {"calls_s": [w for w, _ in runs], "upload_ms": r.stats["upload_ms"], "loop_ms": r.stats["loop_ms"],
                     "download_ms": r.stats["download_ms"], "host_setup_ms": r.stats["host_setup_ms"],
                     "host_loop_ms": r.stats["host_loop_ms"], "host_finish_ms": r.stats["host_finish_ms"],
                     "host_total_ms": r.stats["host_total_ms"], "host_alloc_ms": r.stats["host_alloc_ms"],
                     "host_teardown_ms": r.stats["host_teardown_ms"], "python_run_time_s": r.run_time,
                     "per_call": [{k2: rr.stats[k2] for k2 in ("host_setup_ms", "host_loop_ms", "host_alloc_ms", "host_teardown_ms", "host_total_ms")} for _, rr in runs]}
            if long_T:
                wl200, _ = call(long_T)
                extra["at_200_steps"] = {"value": long_T / wl200, "seconds": wl200}
            del Ap
        else:
            r0, nr = shard.shard_bounds(n, world, rank); c0, mc = shard.shard_bounds(m, world, rank)
            Acol = pin(synth_block(n, 0, n, c0, mc, k, na_frac=wl["na"])); Arow = pin(synth_block(n, r0, nr, 0, m, k, na_frac=wl["na"]))
            walls = []
            for _ in range(3):
                barrier()
                t0 = time.perf_counter()
                s2 = Session(k=k, method=wl["method"], inner_max_iter=wl["inner"], inner_rel_tol=1e-9, precision=args.precision,
                             device=local_rank, comm=comm, shards=(Acol, Arow), shape=(n, m))
                s2.set_factors(Wp, Hp)
                s2.run(args.steps)
                s2.get_factors()
                barrier()
                walls.append(max_over_ranks(time.perf_counter() - t0))
                st2 = s2.stats()
                s2.close()
            wall = sorted(walls)[1]
            h2d, d2h = sum_over_ranks(float(st2["h2d_bytes"])), sum_over_ranks(float(st2["d2h_bytes"]))
            what = ("sharded Session(shards=pinned host A[:,cols_g], A[rows_g,:]) + set_factors + run(steps) + get_factors, all ranks; "
                    "median of 3 (max over ranks each)")
            extra = {"calls_s": walls, "upload_ms": st2["upload_ms"]}
            del Acol, Arow
        wall = max_over_ranks(wall)
        line["e2e"] = {"value": args.steps / wall, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
                       "d2h_bytes_per_step": d2h / args.steps, "seconds": wall, "what": what, **extra}
    if not args.no_cpu and world == 1:
        r = cpu_arm(wl, steps=1, warm=0, want_factors=True)
        line["cpu_baseline"] = {"value": 1.0 / r["sec"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                "variant": "faithful (per-iteration A.t() copy, as src/nnmf.cpp:117,131 executes)",
                                "fair_value": 1.0 / r["fair_sec"], "fair_variant": "transposed copy made once, outside the timed loop"}
        if r["first"] is not None:
            Wc, Hc = r["first"]
            parity.update(against="oracle (C++/OpenMP restatement of the reference), the cpu_baseline iteration, same matrix",
                          rel_W=rel(W_T1, Wc), rel_H=rel(H_T1, Hc),
                          sweeps_gpu=int(sweeps_T1), max_abs_W=float(np.abs(W_T1 - Wc).max()), max_abs_H=float(np.abs(H_T1 - Hc).max()))
            if not (parity["rel_W"] < 1e-5 and parity["rel_H"] < 1e-5):
                print(json.dumps({"error": "T=1 parity against the oracle failed", "parity": parity}), file=sys.stderr)
                sys.exit(3)
    if "against" in parity:
        line["parity"] = parity
    if rank == 0:
        print(json.dumps(line))
    if comm is not None:
        comm.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
