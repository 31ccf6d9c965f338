#!/bin/bash
# N-GPU box: A/B of the operand transport variants (bench only)
N=${N:-8}
mkdir -p gpurun_out
i=0
while IFS= read -r cfg; do
  [ -z "$cfg" ] && continue
  i=$((i+1))
  env $cfg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$i \
    bench.py --gpus $N --steps 40 --warmup 5 --no-e2e > gpurun_out/ab_n${N}_$i.json 2> gpurun_out/ab_n${N}_$i.err
  rc=$?
  python - <<PY
import json
ok=False
for l in open('gpurun_out/ab_n${N}_$i.json'):
    if l.startswith('{'):
        ok=True
        j=json.loads(l); print('n=$N [$cfg]', 'ms/step %.3f'%j['ms_per_step'], {k:round(v,3) for k,v in j['kernel_ms'].items()}, 'host', round(j['host_enqueue_ms_per_step'],3))
if not ok: print('n=$N [$cfg] FAILED rc=$rc'); print(open('gpurun_out/ab_n${N}_$i.err').read()[-1500:])
PY
done <<< "$CONFIGS"
