#!/bin/bash
# single-GPU box: the peer-memory transport between processes sharing GPU 0, then the whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/p2p1_pytest.log 2>&1; echo "multi pytest exit $?"; tail -25 gpurun_out/p2p1_pytest.log
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
