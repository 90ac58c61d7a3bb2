#!/bin/bash
mkdir -p gpurun_out /tmp/rep
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_scd_chain" --launch-skip 9 --launch-count 1 -o /tmp/rep/scd1 python bench.py --config 2 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2k_ncu.log 2>&1
ncu -i /tmp/rep/scd1.ncu-rep --page source --csv --print-source sass > gpurun_out/r2k_scd_h_sass.csv 2>/dev/null
ncu -i /tmp/rep/scd1.ncu-rep --page raw --csv > gpurun_out/r2k_scd_h_raw.csv 2>/dev/null
ls -la gpurun_out/r2k*; head -c 600 gpurun_out/r2k_scd_h_raw.csv | tr ',' '\n' | grep -i "kernel name" -A1 | head
