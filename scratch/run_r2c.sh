#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -rs > gpurun_out/r2c_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r2c_tests.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c_c3.json 2> gpurun_out/r2c_c3.err; tail -2 gpurun_out/r2c_c3.err
python - <<'PY'
import json
for f in ("r2c_bench", "r2c_c3"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        r = d["roofline"]
        print(f, "it/s", round(d["value"], 2), "ms", round(d["ms_per_step"], 4), "frac", round(r["frac"], 4), r["share_of_step"], "cross ms", r.get("cross", {}).get("ms_per_launch"), "solve ms/launch", r["ms_per_launch"])
    except Exception as e:
        print(f, "no line", e)
PY
