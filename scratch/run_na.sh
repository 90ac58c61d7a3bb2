#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/cfg4.json 2> gpurun_out/cfg4.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/cfg4.json"))
    print("config 4", round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],3), d["roofline"]["share_of_step"], d["config"].get("mse_after"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/cfg4.err").read()[-1500:])
PY
