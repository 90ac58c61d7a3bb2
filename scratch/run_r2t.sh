#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_na_path.py -m gpu -x -q 2>&1 | tail -2
for bm in 256; do
  NNLM_NA_BM=$bm timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r2t_c4_$bm.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('bm', $bm, 'it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), r['share_of_step'], 'mse', d['config'].get('mse_after'))"
done 2>&1 | tee gpurun_out/r2t_c4.log
