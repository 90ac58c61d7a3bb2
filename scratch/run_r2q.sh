#!/bin/bash
# warp-pair SCD solver (scd_pair.cuh) vs the single-warp one: correctness under the tile tests, then solve time per shard size
mkdir -p gpurun_out
for ct in 1 2; do
  NNLM_SCD_PAIR_MAX=1000000000 NNLM_SCD_PAIR_CT=$ct timeout 600 python -m pytest tests/test_gpu_scd_tiles.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
done > gpurun_out/r2q_tests.log 2>&1
cat gpurun_out/r2q_tests.log
for mode in "0 1" "1000000000 1" "1000000000 2"; do
 set -- $mode
 for shape in 6250,1250,50 12500,2500,50 25000,5000,50 50000,10000,50 25000,2500,128; do
  NNLM_SCD_PAIR_MAX=$1 NNLM_SCD_PAIR_CT=$2 timeout 300 python bench.py --shape $shape --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('pair_max', $1, 'ct', $2, 'shape', '$shape', 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'solve ms/iter', round(r['share_of_step']['solve']*d['ms_per_step'],4), 'mse', d['config'].get('mse_after'))"
 done
done 2>&1 | tee gpurun_out/r2q_times.log
