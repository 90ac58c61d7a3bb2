#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_cross.py -m gpu -x -q -s -k "128 or config5 or cross" 2>&1 | grep -i "passed\|failed\|config5\|k=128\|rel" | tail -8
for hv in 1 0; do NNLM_TC_HALVES=$hv timeout 300 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('halves_env $hv config 5 N=1 it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), r['share_of_step'])"; done
timeout 500 python scratch/c5_t1_full.py 2>&1 | tail -2
