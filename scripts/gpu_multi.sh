#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/multi_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/multi_pytest.log
for n in 1 $N; do
  if [ $n = 1 ]; then
    timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
  echo "bench n=$n exit $?"; tail -3 gpurun_out/bench_n$n.err
  python - <<PY
import json
for l in open('gpurun_out/bench_n$n.json'):
    if l.startswith('{'):
        j=json.loads(l); print('n=$n', 'ms/step %.3f'%j['ms_per_step'], 'value %.3e'%j['value'], j['kernel_ms'], 'e2e', j['e2e'] and round(j['e2e']['ms_per_step'],3))
PY
done
