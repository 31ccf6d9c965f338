#!/bin/bash
# A/B timing of library variants within ONE box visit: VARIANTS="name1 name2" (files vipant_b200/_lib/libvipant_b200_<name>.so; "cur" = in-tree)
mkdir -p gpurun_out; : > gpurun_out/ab.log
for rep in 1 2; do
for v in $VARIANTS; do
  if [ "$v" = "cur" ]; then lib=""; else lib="$PWD/vipant_b200/_lib/libvipant_b200_$v.so"; fi
  echo "=== $v rep $rep" >> gpurun_out/ab.log
  VIPANT_B200_LIB=$lib timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e >> gpurun_out/ab.log 2>&1
done; done
python - <<'PY'
import json,re
name=None
for line in open('gpurun_out/ab.log'):
    if line.startswith('==='): name=line.strip()
    elif line.startswith('{'):
        j=json.loads(line); print(name, 'ms/step %.3f'%j['ms_per_step'], {k:round(v,3) for k,v in j['kernel_ms'].items()}, j['clocks'])
PY
