#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cross.py tests/test_gpu_scale_parity.py tests/test_gpu_parity.py tests/test_gpu_na_path.py -m gpu -q -s > gpurun_out/r2o_tests.log 2>&1
grep -E "passed|failed|FAILED|config 5|T=1" gpurun_out/r2o_tests.log | tail -12
for c in 2 5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('config', $c, 'it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],4), 'cross ms', r['cross']['ms_per_launch'], 'hbm frac', round(r['cross']['frac'],3), r['share_of_step'])"; done
