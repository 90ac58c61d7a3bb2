#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 python scratch/sanitize.py > gpurun_out/r2p_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2p_$tool.log
  tail -4 gpurun_out/r2p_$tool.log
done
