#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],4), d["roofline"]["share_of_step"], d["config"].get("mse_after"))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/s_c2.json 2> gpurun_out/s_c2.err; show gpurun_out/s_c2.json "config 2"
timeout 300 python bench.py --small --steps 40 --warmup 5 --no-cpu --no-e2e > gpurun_out/s_small.json 2> gpurun_out/s_small.err; show gpurun_out/s_small.json "small"
