#!/bin/bash
# Runs every gpu_check stage in its own process (a device trap poisons the context), bounded.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/stages.log 2>&1
for st in normalize simt tc_fwd tc_bwd retrieval host; do
  echo "=== stage $st" >> gpurun_out/stages.log
  timeout 120 python scripts/gpu_check.py $st ${B:-512} ${D:-512} >> gpurun_out/stages.log 2>&1
  echo "exit $?" >> gpurun_out/stages.log
done
tail -60 gpurun_out/stages.log
