#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_composite.py tests/test_gpu_embed_cache.py tests/test_gpu_infonce.py -q -m gpu --timeout 200 -s > gpurun_out/pytest_gpu1d.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|multi-pair B=" gpurun_out/pytest_gpu1d.log | head -20
