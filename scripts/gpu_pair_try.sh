#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/pair.log
for st in "tc_fwd 512 512" "tc_fwd 2048 512" "tc_fwd 1000 256" "tc_bwd 512 512"; do
  echo "=== $st" >> gpurun_out/pair.log
  timeout 90 python scripts/gpu_check.py $st >> gpurun_out/pair.log 2>&1
  echo "exit $?" >> gpurun_out/pair.log
done
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/pair.log 2>&1
echo "exit $?" >> gpurun_out/pair.log
grep -v "^$" gpurun_out/pair.log | cut -c1-400 | tail -40
