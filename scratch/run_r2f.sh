#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_smi.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_scale_parity.py -m gpu -q --maxfail=20 -s -k "multi or n_gpus or sharded or heavy or auto_precision" > gpurun_out/r2f_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_tests.log
grep -E "passed|failed|FAILED|ranks\]|heavy|Error|error" gpurun_out/r2f_tests.log | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
tail -3 gpurun_out/r2f_bench_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2f_bench_n2.json"))
    print("N=2", d["value"], d["ms_per_step"], d["roofline"]["share_of_step"], d.get("parity"), d.get("e2e", {}).get("value"))
except Exception as e:
    print("no line", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2f_ref_n2.json 2> gpurun_out/r2f_ref_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2f_ref_n2.json')); print('ref arm under torchrun: cores', d['cpu_baseline']['cores'], 'value', d['value'])"
