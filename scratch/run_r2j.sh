#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_error_identity.py tests/test_gpu_parity.py -m gpu -q --maxfail=20 -s > gpurun_out/r2j_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_tests.log
grep -E "passed|failed|FAILED|mse rel" gpurun_out/r2j_tests.log | tail -12
timeout 600 python - > gpurun_out/r2j_trace.json 2> gpurun_out/r2j_trace.err <<'PY'
import json, sys, time, os
sys.path.insert(0, ".")
import numpy as np, nnlm_b200
from nnlm_b200.session import synth_matrix, synth_init, Session
n, m, k, T = 50000, 10000, 50, 40
A = synth_matrix(n, m, k); W0, H0 = synth_init(n, m, k)
out = {}
for name, kw in (("trace0", dict(trace=0)), ("trace2_mkl_all", dict(trace=2)), ("trace2_mkl_final", dict(trace=2, mkl_trace="final"))):
    best = None
    for rep in range(2):
        r = nnlm_b200.nnmf(A, k, init={"W": W0, "H": H0}, max_iter=T, rel_tol=-1, show_warning=False, check_k=False, precision=2, **kw)
        best = r if best is None or r.stats["loop_ms"] < best.stats["loop_ms"] else best
    out[name] = {"loop_ms": best.stats["loop_ms"], "iters_per_s_loop": T / (best.stats["loop_ms"] * 1e-3), "records": len(best.mse),
                 "error_ms_total": best.stats["error_ms"], "mse_last": float(best.mse[-1]), "mkl_last": float(best.mkl[-1]),
                 "mse_from_identity": best.stats["mse_from_identity"]}
# the two evaluations on the same factors
with Session(A, k=k, method=1, precision=2, timing=True) as s:
    s.set_factors(best.W, best.H)
    t0 = time.perf_counter(); e_tc = s.error(); t_tc = time.perf_counter() - t0
    t0 = time.perf_counter(); e_tc = s.error(); t_tc = time.perf_counter() - t0
    out["error_tc"] = {"mse": e_tc[0], "mkl": e_tc[1], "wall_ms": t_tc * 1e3, "error_ms": s.stats()["error_ms"]}
os.environ["NNLM_ERR_FP64"] = "1"
with Session(A, k=k, method=1, precision=2, timing=True) as s:
    s.set_factors(best.W, best.H)
    t0 = time.perf_counter(); e64 = s.error(); t64 = time.perf_counter() - t0
    t0 = time.perf_counter(); e64 = s.error(); t64 = time.perf_counter() - t0
    out["error_fp64_pass"] = {"mse": e64[0], "mkl": e64[1], "wall_ms": t64 * 1e3}
print(json.dumps(out))
PY
cat gpurun_out/r2j_trace.json; tail -3 gpurun_out/r2j_trace.err
