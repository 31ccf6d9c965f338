#!/bin/bash
# N-GPU box: transports side by side (tests at 2 ranks, bench at N ranks)
N=${N:-2}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/p2p2_pytest.log 2>&1; echo "multi pytest exit $?"; tail -4 gpurun_out/p2p2_pytest.log; fi
for tr in ${TRANSPORTS:-p2p nccl}; do
  VIPANT_TRANSPORT=$tr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_$tr.json 2> gpurun_out/bench_n${N}_$tr.err
  echo "bench n=$N $tr exit $?"; tail -3 gpurun_out/bench_n${N}_$tr.err
  python - <<PY
import json
for l in open('gpurun_out/bench_n${N}_$tr.json'):
    if l.startswith('{'):
        j=json.loads(l); print('n=$N $tr', 'ms/step %.3f'%j['ms_per_step'], 'value %.3e'%j['value'], {k:round(v,3) for k,v in j['kernel_ms'].items()}, 'host', round(j['host_enqueue_ms_per_step'],3), 'e2e', j['e2e'] and round(j['e2e']['ms_per_step'],3), 'loss', j['loss'])
PY
done
