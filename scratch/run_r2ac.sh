#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -s -k "config4 or missing" 2>&1 | grep -v "^$" | tail -6
for c in 4 3; do timeout 900 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2ac_bench_c$c.json 2> gpurun_out/r2ac_bench_c$c.err; python -c "
import json; d=json.loads(open('gpurun_out/r2ac_bench_c$c.json').read().strip().splitlines()[-1]); r=d['roofline']; print('config $c it/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['sample'][:60], 'parity', {k: d.get('parity',{}).get(k) for k in ('rel_W','rel_H')})"; tail -2 gpurun_out/r2ac_bench_c$c.err; done
