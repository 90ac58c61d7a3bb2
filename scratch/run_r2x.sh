#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_na_path.py tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
for solver in square tri; do
NNLM_NA_SOLVER=$solver timeout 300 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r2x_c4.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$solver it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), r['share_of_step'], 'solve ms', round(r['share_of_step']['solve']*d['ms_per_step'],3), 'mse', d['config'].get('mse_after'), 'sweeps', d['config'].get('avg_inner_sweeps_per_column'))"
done
tail -3 gpurun_out/r2x_c4.err
