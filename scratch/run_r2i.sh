#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -rs > gpurun_out/r2i_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r2i_tests.log | tail -12
for c in 2 3 4; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2i_c$c.json 2> gpurun_out/r2i_c$c.err; done
python - <<'PY'
import json
for c in (2, 3, 4):
    try:
        d = json.loads(open(f"gpurun_out/r2i_c{c}.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print("config", c, "it/s", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), "frac", round(r["frac"], 4), r["bound"], r["share_of_step"])
    except Exception as e:
        print(c, "no line", e)
PY
