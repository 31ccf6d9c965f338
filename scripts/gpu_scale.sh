#!/bin/bash
# scaling sweep on one box: N = 1, 2, 4, 8 (as available)
mkdir -p gpurun_out
for n in ${NS:-1 2 4 8}; do
  if [ $n = 1 ]; then
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "bench n=$n exit $?"
  python - <<PY
import json
for l in open('gpurun_out/scale_n$n.json'):
    if l.startswith('{'):
        j=json.loads(l); km=j['kernel_ms']; print('n=$n', 'ms/step %.3f'%j['ms_per_step'], 'value %.3e'%j['value'], 'kernels_sum %.3f'%sum(km.values()), {k:round(v,3) for k,v in km.items()}, 'e2e', j['e2e'] and round(j['e2e']['ms_per_step'],3))
PY
done
