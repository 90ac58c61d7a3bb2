#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --config 3 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/kl.json 2> gpurun_out/kl.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/kl.json")); print("config 3", round(d["value"],2), "iters/s", round(d["ms_per_step"],1), "ms", d["config"].get("mse_after"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/kl.err").read()[-800:])
PY
