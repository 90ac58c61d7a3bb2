#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_error_identity.py tests/test_gpu_parity.py -m gpu -q --maxfail=40 -s > gpurun_out/r2b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
grep -E "passed|failed|T=1|KL method|heavy" gpurun_out/r2b_tests.log | tail -30
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err; tail -2 gpurun_out/r2b_c3.err
NNLM_KL_SLOW=1 timeout 600 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2b_c3_slow.json 2> gpurun_out/r2b_c3_slow.err
python - <<'PY'
import json
for f in ("r2b_c3", "r2b_c3_slow"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["ms_per_step"], d["config"].get("mkl_after"), d["roofline"]["share_of_step"], d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "no line", e)
PY
