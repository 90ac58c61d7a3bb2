#!/bin/bash
# the driver's own sequence on the shipped binary
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ag_ref.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2ag_ref.json').read().strip().splitlines()[-1]); print('reference arm', round(d['value'],3), d['unit'], 'steps', d['steps'], 'cores', d['cpu_baseline']['cores'])"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ag_bench.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2ag_bench.json').read().strip().splitlines()[-1]); print('ours', round(d['value'],2), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity']['rel_W'], d['parity']['rel_H'], 'launches', d['gpu_launches'])"
