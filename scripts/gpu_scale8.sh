#!/bin/bash
# 8-GPU box: N = 8, 4, 2 with the default (peer-memory) transport
mkdir -p gpurun_out
for n in ${NS:-8 4 2}; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
    bench.py --gpus $n --steps 40 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  rc=$?
  python - <<PY
import json
ok=False
for l in open('gpurun_out/scale_n$n.json'):
    if l.startswith('{'):
        ok=True
        j=json.loads(l); print('n=$n', j['config']['transport'], 'ms/step %.3f'%j['ms_per_step'], 'value %.4e'%j['value'], {k:round(v,3) for k,v in j['kernel_ms'].items()}, 'host', round(j['host_enqueue_ms_per_step'],3), 'e2e', j['e2e'] and round(j['e2e']['ms_per_step'],3), j['clocks'])
if not ok: print('n=$n FAILED rc=$rc'); print(open('gpurun_out/scale_n$n.err').read()[-1500:])
PY
done
