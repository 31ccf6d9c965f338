#!/bin/bash
# Round-2 first visit (8 GPUs): A/B of the operand transfer variants at N = 8, B = 32768 x 512.
#   gpurun --gpus 8 --timeout 600 -- 'bash scripts/r2_ab_n8.sh'
N=8 CONFIGS="VIPANT_TRANSPORT=p2p
VIPANT_TRANSPORT=p2p VPA_P2P_MODE=stream VPA_P2P_STREAM_CTAS=64
VIPANT_TRANSPORT=p2p VPA_P2P_MODE=stream VPA_P2P_STREAM_CTAS=128
VIPANT_TRANSPORT=p2p VPA_P2P_PULL_CTAS=148 VPA_P2P_PULL_THREADS=128
VIPANT_TRANSPORT=p2p VPA_P2P_PLAN=serial
VIPANT_TRANSPORT=p2p VPA_P2P_MODE=nvls VIPANT_REQUIRE_P2P=1" bash scripts/gpu_p2p8.sh
# (bring the nvls variant up at N = 2 first: gpurun --gpus 2 -- 'VIPANT_TEST_NVLS=1 python -m pytest tests/test_gpu_multi.py -x -q -k nvls')
