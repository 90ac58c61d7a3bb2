#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || exit 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],4), d["roofline"]["share_of_step"], d["config"].get("avg_inner_sweeps_per_column"), d["config"].get("mse_after"))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
for cfg in "2 0" "2 2" "2 4"; do
  set -- $cfg
  NNLM_SCD_IMPL=$1 NNLM_SCD_CT=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/ab_$1_$2.json 2> gpurun_out/ab_$1_$2.err
  show gpurun_out/ab_$1_$2.json "impl=$1 ct=$2"
done
for c in 3 4; do
  NNLM_SCD_IMPL=2 timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/cfg$c.json 2> gpurun_out/cfg$c.err
  show gpurun_out/cfg$c.json "config $c"
done
