#!/bin/bash
# 8-GPU box: strong scaling of config 2 at N = 8, 4, 2, 1 and config 5 at N = 8; the in-call multi-GPU path at 4 GPUs
mkdir -p gpurun_out
run() { # N config steps tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1 + 10 * $2)) bench.py --gpus $1 --config $2 --steps $3 --warmup 5 --no-cpu > gpurun_out/r2l_c$2_n$1.json 2> gpurun_out/r2l_c$2_n$1.err
}
run 8 2 20; run 4 2 20; run 2 2 20
timeout 600 python bench.py --gpus 1 --config 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2l_c2_n1.json 2> gpurun_out/r2l_c2_n1.err
run 8 5 10
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/r2l_tests.log 2>&1
tail -3 gpurun_out/r2l_tests.log
python - <<'PY'
import json
for c, ns in ((2, (1, 2, 4, 8)), (5, (8,))):
    for n in ns:
        try:
            d = json.loads(open(f"gpurun_out/r2l_c{c}_n{n}.json").read().strip().splitlines()[-1])
            print("config", c, "N", n, "it/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), d["roofline"]["share_of_step"], "parity", {k: d.get("parity", {}).get(k) for k in ("rel_W", "rel_H", "sweeps_equal")}, "e2e", d.get("e2e", {}).get("value"))
        except Exception as e:
            print(c, n, "no line", e)
PY
