#!/bin/bash
# one-box scaling run: sharded-path check at 8 ranks, config 2 at N=8,4,2,1, config 5 at N=8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep "ranks\]" | tail -6
for n in 8 4 2; do
  timeout 400 $TR --nproc-per-node $n --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/scale_c2_n$n.json
  python -c "import json; d=json.load(open('gpurun_out/scale_c2_n$n.json')); print('config2 N=$n', round(d['value'],1), 'iters/s', round(d['ms_per_step'],3), 'ms', d['roofline']['share_of_step'], 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))"
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/scale_c2_n1.json
python -c "import json; d=json.load(open('gpurun_out/scale_c2_n1.json')); print('config2 N=1', round(d['value'],1), 'iters/s', round(d['ms_per_step'],3), 'ms', d['roofline']['share_of_step'], 'e2e', round(d['e2e']['value'],1))"
timeout 500 $TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --config 5 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/scale_c5_n8.json
python -c "import json; d=json.load(open('gpurun_out/scale_c5_n8.json')); print('config5 N=8', round(d['value'],1), 'iters/s', round(d['ms_per_step'],3), 'ms', d['roofline']['share_of_step'], 'frac', round(d['roofline']['frac'],3))"
