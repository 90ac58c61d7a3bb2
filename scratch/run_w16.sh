#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],4), d["roofline"]["share_of_step"])
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
}
run() {
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/$1_c2.json 2> gpurun_out/$1_c2.err; show gpurun_out/$1_c2.json "$1 config 2"
  NNLM_SCD_CT2_MIN=4736 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/$1_c2b.json 2> gpurun_out/$1_c2b.err; show gpurun_out/$1_c2b.json "$1 config 2 (ct2 both halves)"
  NNLM_SCD_CT2_MIN=100000000 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/$1_c2c.json 2> gpurun_out/$1_c2c.err; show gpurun_out/$1_c2c.json "$1 config 2 (ct1 both halves)"
  timeout 300 python bench.py --small --steps 40 --warmup 5 --no-cpu --no-e2e > gpurun_out/$1_small.json 2> gpurun_out/$1_small.err; show gpurun_out/$1_small.json "$1 small"
}
mv nnlm_b200/libnnlm_b200.so nnlm_b200/libnnlm_b200_w12.so
for w in 8 10; do cp nnlm_b200/libnnlm_b200_w$w.so nnlm_b200/libnnlm_b200.so; run w$w; done
mv nnlm_b200/libnnlm_b200_w12.so nnlm_b200/libnnlm_b200.so
