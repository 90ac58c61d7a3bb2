#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],4), d["roofline"]["share_of_step"], d["config"].get("mse_after"))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
}
run() {
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/$1_c2.json 2> gpurun_out/$1_c2.err; show gpurun_out/$1_c2.json "$1 config 2"
  timeout 300 python bench.py --small --steps 40 --warmup 5 --no-cpu --no-e2e > gpurun_out/$1_small.json 2> gpurun_out/$1_small.err; show gpurun_out/$1_small.json "$1 small"
}
run rem1
mv nnlm_b200/libnnlm_b200.so nnlm_b200/libnnlm_b200_rem1.so; cp nnlm_b200/libnnlm_b200_rem0.so nnlm_b200/libnnlm_b200.so
run rem0
mv nnlm_b200/libnnlm_b200_rem1.so nnlm_b200/libnnlm_b200.so
