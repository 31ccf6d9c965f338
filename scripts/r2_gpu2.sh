#!/bin/bash
# Round 2, two-GPU visit: multi-GPU parity over real NVLink (fused relay, NCCL, host transports, NVLS bring-up), then the
# N = 2 bench with the transport variants.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/r2_gpu2.sh'
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
VIPANT_TEST_NVLS=1 timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_infonce.py -q -m gpu --timeout 200 -k "sharded or nvls or fifty or three_pairs or two_devices" > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu2.log; tail -25 gpurun_out/pytest_gpu2.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e > gpurun_out/n2_$name.json 2> gpurun_out/n2_$name.err
  echo "== $name rc=$?"
  python - <<PY
import json
ok=False
for l in open('gpurun_out/n2_$name.json'):
    if l.startswith('{'):
        ok=True; j=json.loads(l)
        print('$name', j['config']['transport'], 'ms/step %.4f'%j['ms_per_step'], {k:round(v,4) for k,v in j['kernel_ms'].items()}, 'launches/step', j.get('gpu_launches_per_step'), 'parity', j['parity'] and (j['parity']['ok'], j['parity']['grad_rel']), j['clocks']['sm_mhz'])
if not ok: print(open('gpurun_out/n2_$name.err').read()[-1500:])
PY
}
run relay20 VIPANT_TRANSPORT=p2p VIPANT_REQUIRE_P2P=1
run relay8 VIPANT_TRANSPORT=p2p VIPANT_REQUIRE_P2P=1 VPA_P2P_RELAY_CTAS=8
run nvls VIPANT_TRANSPORT=p2p VIPANT_REQUIRE_P2P=1 VPA_P2P_MODE=nvls
run nccl VIPANT_TRANSPORT=nccl
