#!/bin/bash
# one-GPU check of programmatic dependent launch / carve-out knobs: parity tests + N=1 bench with each
mkdir -p gpurun_out
VPA_PDL=1 VPA_CARVEOUT=1 timeout 600 python -m pytest tests/test_gpu_infonce.py tests/test_gpu_multi.py -q -m gpu --timeout 200 -x > gpurun_out/pytest_gpu1c.log 2>&1; echo "pytest(pdl) exit $?"; tail -4 gpurun_out/pytest_gpu1c.log
for v in "VPA_PDL=0 VPA_CARVEOUT=0" "VPA_PDL=1 VPA_CARVEOUT=0" "VPA_PDL=0 VPA_CARVEOUT=1" "VPA_PDL=1 VPA_CARVEOUT=1"; do
  env $v timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-eager --no-cpu-baseline > gpurun_out/n1_knobs.json 2> gpurun_out/n1_knobs.err
  python - <<PY
import json
for l in open('gpurun_out/n1_knobs.json'):
    if l.startswith('{'):
        j=json.loads(l); print('$v', 'ms/step %.4f'%j['ms_per_step'], {k:round(v,4) for k,v in j['kernel_ms'].items()}, 'parity', j['parity']['ok'], j['clocks']['sm_mhz'])
PY
done
