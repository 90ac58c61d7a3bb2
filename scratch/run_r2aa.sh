#!/bin/bash
# final build: whole GPU suite, sanitizer (memcheck) on the small cases, config 3 bench line + launch list + full capture of the product kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2aa_tests.log; cat gpurun_out/r2aa_tests.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scratch/sanitize.py > gpurun_out/r2aa_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2aa_memcheck.log; tail -2 gpurun_out/r2aa_memcheck.log
timeout 900 python bench.py --config 3 --steps 20 --warmup 5 > gpurun_out/r2aa_bench_c3.json 2> gpurun_out/r2aa_bench_c3.err; python -c "
import json; d=json.loads(open('gpurun_out/r2aa_bench_c3.json').read().strip().splitlines()[-1]); r=d['roofline']; print('config 3 it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2aa_launches_c3.csv python bench.py --no-cpu --no-e2e --config 3 --steps 2 --warmup 3 > /tmp/l3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_error_tc|k_solve_kl_fast" --launch-skip 2 -c 3 -o /tmp/r2aa -f python bench.py --config 3 --steps 1 --warmup 1 --no-cpu --no-e2e > /tmp/r2aa.log 2>&1
ncu -i /tmp/r2aa.ncu-rep --page raw --csv > gpurun_out/r2aa_full_c3.csv 2>/dev/null
ls -la gpurun_out | grep r2aa
