#!/bin/bash
mkdir -p gpurun_out
for bm in 256 128; do
  kn=$([ $bm = 256 ] && echo k_mask_tc || echo k_cross_tc)
  NNLM_NA_BM=$bm timeout 900 ncu --set full --clock-control none -k regex:$kn --launch-skip 30 -c 2 -o /tmp/r2u_$bm -f python bench.py --config 4 --steps 1 --warmup 1 --no-cpu --no-e2e > /tmp/r2u_$bm.log 2>&1
  ncu -i /tmp/r2u_$bm.ncu-rep --page raw --csv > gpurun_out/r2u_full_$bm.csv 2>/dev/null
  tail -2 /tmp/r2u_$bm.log
done
ls -la gpurun_out/r2u_*
