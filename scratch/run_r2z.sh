#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -m gpu -x -q -k "kl or KL or mkl or method or odd" 2>&1 | tail -3
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('config 3 it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'mkl', d['config'].get('mkl_after'))"
cp scratch/bin/libnnlm_b200_prof.so nnlm_b200/libnnlm_b200.so
timeout 300 python bench.py --config 3 --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 | grep "klf prof" | sort | uniq -c | sort -rn | head -4
