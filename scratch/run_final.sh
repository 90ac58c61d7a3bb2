#!/bin/bash
# final round-1 evidence: un-profiled bench line, ncu launch list, ncu full capture of the two hot kernels
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1_n.json 2> gpurun_out/bench_r1_n.err
tail -c 600 gpurun_out/bench_r1_n.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1_n_ref.json 2> gpurun_out/bench_r1_n_ref.err
cat gpurun_out/bench_r1_n_ref.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_n.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_r1_n.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_cross_tc|k_scd_chain" -s 8 -c 4 -o gpurun_out/full_r1_n python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/full_r1_n.log 2>&1
ls -la gpurun_out | tail -8
