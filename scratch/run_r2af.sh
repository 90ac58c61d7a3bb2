#!/bin/bash
# FINAL build of round 2: whole GPU suite, smoke, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2af_tests.log; cat gpurun_out/r2af_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2af_bench_c2.json 2> gpurun_out/r2af_bench_c2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2af_bench_c2.json').read().strip().splitlines()[-1]); r=d['roofline']; print('config 2 it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value'],1), 'cpu', round(d['cpu_baseline']['value'],3), 'parity', {k: d['parity'][k] for k in ('rel_W','rel_H')}, 'launches', d['gpu_launches'], 'clocks', d['clocks'])"
