#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/cluster.log
for c in 1 2 4; do
  echo "=== VPA_CLUSTER=$c" >> gpurun_out/cluster.log
  VPA_CLUSTER=$c timeout 120 python scripts/gpu_check.py tc_bwd 512 512 >> gpurun_out/cluster.log 2>&1
  echo "exit $?" >> gpurun_out/cluster.log
  VPA_CLUSTER=$c timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/cluster.log 2>&1
  echo "exit $?" >> gpurun_out/cluster.log
done
grep -E "===|bf16\]|exit|kernel_ms" gpurun_out/cluster.log | sed -E 's/.*("ms_per_step": [0-9.]+).*("kernel_ms": \{[^}]*\}).*/\1 \2/' 
