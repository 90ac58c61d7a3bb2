#!/bin/bash
cp scratch/bin/libnnlm_b200_prof.so nnlm_b200/libnnlm_b200.so
timeout 300 python bench.py --config 3 --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 | grep "klf prof" | sort | uniq -c | sort -rn | head -4
