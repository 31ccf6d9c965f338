#!/bin/bash
# Final one-GPU validation of the round-2 tree: smoke, all GPU parity tests, bench (both arms).
#   gpurun --timeout 300 -- 'bash scripts/r2_final.sh'
mkdir -p gpurun_out
timeout 80 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 170 python -m pytest tests -q -x -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 110 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
