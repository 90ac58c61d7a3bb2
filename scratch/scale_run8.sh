#!/bin/bash
# one-box 8-GPU run: sharded-path parity check at 8 ranks, config 2 and config 5 at N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep "ranks\]" | tail -6
timeout 300 $TR --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/scale2_c2_n8.json
python -c "import json; d=json.load(open('gpurun_out/scale2_c2_n8.json')); print('config2 N=8', round(d['value'],1), 'iters/s', round(d['ms_per_step'],3), 'ms', d['roofline']['share_of_step'], 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))"
timeout 400 $TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --config 5 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/scale2_c5_n8.json
python -c "import json; d=json.load(open('gpurun_out/scale2_c5_n8.json')); print('config5 N=8', round(d['value'],1), 'iters/s', round(d['ms_per_step'],3), 'ms', d['roofline']['share_of_step'], 'frac', round(d['roofline']['frac'],3))"
