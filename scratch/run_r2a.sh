#!/bin/bash
# round 2, call A: full GPU test suite + headline bench (with parity / e2e) + reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -rs --durations=15 > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_bench.json | head -c 3000; tail -3 gpurun_out/r2a_bench.err
