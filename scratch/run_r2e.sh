#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -m gpu -q --maxfail=40 -s -k "kl or KL or dense_parity or missing_parity" > gpurun_out/r2e_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_tests.log
grep -E "passed|failed|FAILED|KL method" gpurun_out/r2e_tests.log | tail -14
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2e_c3.json 2> gpurun_out/r2e_c3.err; tail -2 gpurun_out/r2e_c3.err
python - <<'PY'
import json
for f in ("r2e_c3",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        r = d["roofline"]
        print(f, "it/s", round(d["value"], 2), "ms", round(d["ms_per_step"], 4), "frac", round(r["frac"], 4), "solve ms/launch", r["ms_per_launch"])
    except Exception as e:
        print(f, "no line", e)
PY
