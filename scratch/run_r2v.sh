#!/bin/bash
mkdir -p gpurun_out
NNLM_NA_PAIRS=1 timeout 300 python -m pytest tests/test_gpu_na_path.py -m gpu -x -q 2>&1 | tail -15
echo "tests rc=$?"
NNLM_NA_PAIRS=1 timeout 300 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r2v_c4.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('pairs it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), r['share_of_step'], 'mse', d['config'].get('mse_after'))"
tail -3 gpurun_out/r2v_c4.err
nvidia-smi --query-gpu=name,memory.used --format=csv
