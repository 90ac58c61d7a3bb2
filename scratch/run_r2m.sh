#!/bin/bash
mkdir -p gpurun_out
# per-GPU shard sizes of config 2 at N = 8 (6250 x 1250 columns solved per GPU) emulated on one GPU by shape: the solver sees
# ncol = n (W-half) and m (H-half); cross-product sizes differ from the sharded run but only the solve share is read here
for thr in 0 100000; do
 for shape in 6250,1250,50 12500,2500,50 25000,5000,50 50000,10000,50; do
  NNLM_SCD_WARP_MAX=$thr timeout 300 python bench.py --shape $shape --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('thr', $thr, 'shape', '$shape', 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'solve ms/iter', round(r['share_of_step']['solve']*d['ms_per_step'],4))"
 done
done
