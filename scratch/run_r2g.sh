#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_na_path.py tests/test_gpu_scale_parity.py tests/test_gpu_parity.py tests/test_gpu_cross.py -m gpu -q --maxfail=40 -s -k "missing or na or NA or mask or cross" > gpurun_out/r2g_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_tests.log
grep -E "passed|failed|FAILED|NA |k=.*<=|Error" gpurun_out/r2g_tests.log | tail -30
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2g_c4.json 2> gpurun_out/r2g_c4.err; tail -2 gpurun_out/r2g_c4.err
NNLM_NA_FP64=1 timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2g_c4_fp64.json 2> gpurun_out/r2g_c4_fp64.err
python - <<'PY'
import json
for f in ("r2g_c4", "r2g_c4_fp64"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "it/s", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), "mse", d["config"]["mse_after"], r["share_of_step"])
    except Exception as e:
        print(f, "no line", e)
PY
