#!/bin/bash
# final build of round 2: whole GPU suite, smoke, the three single-GPU bench lines with cpu legs, the reference arm, launch lists
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2y_tests.log; cat gpurun_out/r2y_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for c in 2 3 4; do timeout 900 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2y_bench_c$c.json 2> gpurun_out/r2y_bench_c$c.err; python -c "
import json; d=json.loads(open('gpurun_out/r2y_bench_c$c.json').read().strip().splitlines()[-1]); r=d['roofline']; print('config $c it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value'],1) if d.get('e2e') else None, 'cpu', d['cpu_baseline']['value'], 'parity', d.get('parity'))"; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2y_ref_c2.json 2> gpurun_out/r2y_ref_c2.err; tail -c 600 gpurun_out/r2y_ref_c2.json
for c in 2 3 4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2y_launches_c$c.csv python bench.py --no-cpu --no-e2e --config $c --steps 2 --warmup 3 > /tmp/r2y_l$c.log 2>&1
done
ls -la gpurun_out | grep r2y
