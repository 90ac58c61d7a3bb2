#!/bin/bash
# 8-GPU box, final build: configs 2, 5, 4, 3 at N = 8; the multi-GPU tests
mkdir -p gpurun_out
run() { # N config steps
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1 + 10 * $2)) bench.py --gpus $1 --config $2 --steps $3 --warmup 5 --no-cpu > gpurun_out/r2ab_c$2_n$1.json 2> gpurun_out/r2ab_c$2_n$1.err
}
run 8 2 20; run 8 5 10; run 8 4 10; run 8 3 10
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2ab_tests.log 2>&1
tail -3 gpurun_out/r2ab_tests.log
python - <<'PY'
import json
for c in (2, 5, 4, 3):
    try:
        d = json.loads(open(f"gpurun_out/r2ab_c{c}_n8.json").read().strip().splitlines()[-1])
        print("config", c, "N 8 it/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), d["roofline"]["share_of_step"], "parity", {k: d.get("parity", {}).get(k) for k in ("rel_W", "rel_H", "sweeps_equal")}, "e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(c, "no line", e); print(open(f"gpurun_out/r2ab_c{c}_n8.err").read()[-600:])
PY
