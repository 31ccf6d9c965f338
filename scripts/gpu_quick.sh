#!/bin/bash
# quick correctness (tc stages + ragged pytest subset) and timing
mkdir -p gpurun_out
: > gpurun_out/quick.log
timeout 120 python scripts/gpu_check.py tc_bwd 512 512 >> gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log
timeout 300 python -m pytest tests/test_gpu_infonce.py -x -q -k "ragged or golden or deterministic or shard or regimes or host_buffer" >> gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_quick.json 2>> gpurun_out/quick.log; echo "exit $?" >> gpurun_out/quick.log
grep -E "bf16\]|exit|passed|failed|Error|error" gpurun_out/quick.log | cut -c1-220
grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_quick.json; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench_quick.json
