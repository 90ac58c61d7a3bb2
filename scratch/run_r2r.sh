#!/bin/bash
# warp-team SCD solver integrated: full GPU suite, sanitizer on the small cases, solve time per shard size with teams on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2r_tests.log; cat gpurun_out/r2r_tests.log
for tool in synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python scratch/sanitize.py > gpurun_out/r2r_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2r_$tool.log
  tail -3 gpurun_out/r2r_$tool.log
done
for team in 0 1; do
 for shape in 2000,500,10 6250,1250,50 12500,2500,50 25000,5000,50 50000,10000,50 25000,2500,128; do
  NNLM_SCD_TEAM=$team timeout 300 python bench.py --shape $shape --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('team', $team, 'shape', '$shape', 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'solve ms/iter', round(r['share_of_step']['solve']*d['ms_per_step'],4), 'mse', d['config'].get('mse_after'))"
 done
done 2>&1 | tee gpurun_out/r2r_times.log
