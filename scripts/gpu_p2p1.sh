#!/bin/bash
# single-GPU box: the peer-memory transport between processes sharing GPU 0, then the whole GPU suite (+ optional bench)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/p2p1_pytest.log 2>&1; echo "multi pytest exit $?"; tail -3 gpurun_out/p2p1_pytest.log
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
if [ -n "$BENCH" ]; then
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  python - <<PY
import json
for l in open('gpurun_out/bench.json'):
    if l.startswith('{'):
        j=json.loads(l); print('n=1', 'ms/step %.3f'%j['ms_per_step'], {k:round(v,3) for k,v in j['kernel_ms'].items()}, 'host', round(j['host_enqueue_ms_per_step'],3), 'e2e', j['e2e'] and round(j['e2e']['ms_per_step'],3), 'loss', j['loss'])
PY
  for R in 2 4 8; do python scripts/shard_bench.py $R 32768 30 2>&1 | tail -1; done
fi
