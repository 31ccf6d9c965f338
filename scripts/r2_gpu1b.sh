#!/bin/bash
# quick one-GPU re-check of the scoring / multilabel kernels + scoring bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_multilabel.py -q -m gpu --timeout 200 > gpurun_out/pytest_gpu1b.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu1b.log
timeout 300 python scripts/score_bench.py > gpurun_out/scoring.json 2> gpurun_out/scoring.err; echo "score_bench exit $?"; cat gpurun_out/scoring.json; tail -3 gpurun_out/scoring.err
