#!/bin/bash
# wide (256-column) exact mask contraction + in-kernel zero fill of the partial slots: NA / cross tests, then config 4 and 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_na_path.py tests/test_gpu_cross.py tests/test_gpu_scale_parity.py tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2s_tests.log; cat gpurun_out/r2s_tests.log
for bm in 128 256; do
  NNLM_NA_BM=$bm timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r2s_c4_$bm.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('bm', $bm, 'it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), r['share_of_step'], 'mse', d['config'].get('mse_after'))"
done 2>&1 | tee gpurun_out/r2s_c4.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('config 2 it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],4), r['share_of_step'], 'cross frac', r['cross']['frac'])" | tee gpurun_out/r2s_c2.log
