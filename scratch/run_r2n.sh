#!/bin/bash
# final-build profiles: launch lists for configs 2 and 4, full captures of the kernels added / changed after run_r2h
mkdir -p gpurun_out /tmp/rep
B="python bench.py --no-cpu --no-e2e"
for c in 2 4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches_c$c.csv $B --config $c --steps 2 --warmup 3 > gpurun_out/r2n_l$c.log 2>&1
done
timeout 900 ncu --set full --clock-control none -k regex:"k_cross_tc|k_z_slices|k_fold|k_solve_batch_packed|k_factor_prep|k_error_tc|k_split_rows" --launch-skip 100 --launch-count 16 -o /tmp/rep/c4 $B --config 4 --steps 1 --warmup 3 > gpurun_out/r2n_f4.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_error_tc|k_factor_prep|k_split_rows|k_absmax_bits" --launch-count 8 -o /tmp/rep/c2 $B --config 2 --steps 1 --warmup 3 > gpurun_out/r2n_f2.log 2>&1
for c in 2 4; do ncu -i /tmp/rep/c$c.ncu-rep --page raw --csv > gpurun_out/r2n_full_c$c.csv 2>/dev/null; done
for c in 2 3 4; do timeout 900 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2n_bench_c$c.json 2> gpurun_out/r2n_bench_c$c.err; done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2n_ref_c2.json 2> gpurun_out/r2n_ref_c2.err
ls -la gpurun_out | grep r2n
