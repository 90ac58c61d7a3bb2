#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --config 4 --steps 10 --warmup 5 --no-cpu > gpurun_out/r2ad_c4_n8.json 2> gpurun_out/r2ad_c4_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2ad_c4_n8.json").read().strip().splitlines()[-1])
    print("config 4 N 8 it/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), d["roofline"]["share_of_step"], "parity", {k: d.get("parity", {}).get(k) for k in ("rel_W", "rel_H", "sweeps_equal")}, "e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2ad_c4_n8.err").read()[-800:])
PY
