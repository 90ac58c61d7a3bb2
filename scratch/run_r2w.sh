#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2w_tests.log; cat gpurun_out/r2w_tests.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scratch/sanitize.py > gpurun_out/r2w_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2w_memcheck.log; tail -2 gpurun_out/r2w_memcheck.log
timeout 600 python bench.py --config 4 --steps 20 --warmup 5 > gpurun_out/r2w_bench_c4.json 2> gpurun_out/r2w_bench_c4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2w_bench_c4.json').read().strip().splitlines()[-1]); r=d['roofline']; print('c4 it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), r['share_of_step'], 'parity', d.get('parity'))"
timeout 900 ncu --set full --clock-control none -k regex:k_mask_tc2 --launch-skip 30 -c 2 -o /tmp/r2w -f python bench.py --config 4 --steps 1 --warmup 1 --no-cpu --no-e2e > /tmp/r2w.log 2>&1
ncu -i /tmp/r2w.ncu-rep --page raw --csv > gpurun_out/r2w_full_c4.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches_c4.csv python bench.py --config 4 --steps 2 --warmup 3 --no-cpu --no-e2e > /tmp/r2w_l.log 2>&1
ls -la gpurun_out/r2w_*
