#!/bin/bash
# round 2, call H: ncu launch lists + full captures for configs 2, 3, 4 (reports are reduced to CSV on the box: gpurun_out <= 64 MiB)
mkdir -p gpurun_out /tmp/rep
B="python bench.py --no-cpu --no-e2e"
for c in 2 3 4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches_c$c.csv $B --config $c --steps 2 --warmup 3 > gpurun_out/r2h_l$c.log 2>&1
done
timeout 900 ncu --set full --clock-control none -k regex:"k_" --launch-skip 70 --launch-count 44 -o /tmp/rep/c2 $B --config 2 --steps 2 --warmup 3 > gpurun_out/r2h_f2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_solve_kl_fast|k_factor_rows|k_rowsum" --launch-skip 6 --launch-count 8 -o /tmp/rep/c3 $B --config 3 --steps 1 --warmup 3 > gpurun_out/r2h_f3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_cross_tc|k_z_slices|k_fold|k_solve_batch_packed|k_mask_planes|k_error|k_factor_stats" --launch-skip 200 --launch-count 24 -o /tmp/rep/c4 $B --config 4 --steps 1 --warmup 3 > gpurun_out/r2h_f4.log 2>&1
for c in 2 3 4; do ncu -i /tmp/rep/c$c.ncu-rep --page raw --csv > gpurun_out/r2h_full_c$c.csv 2>/dev/null; done
ls -la /tmp/rep gpurun_out | grep -E "c[234]" | head
