#!/bin/bash
# Round 2, eight-GPU visit: the N = 8 / 4 / 2 lines of the final build (default transport; the driver's SCALE run does the same).
#   gpurun --gpus 8 --timeout 500 -- 'bash scripts/r2_gpu8c.sh'
mkdir -p gpurun_out
run() {  # n, name, extra bench args, env...
  n=$1; name=$2; extra=$3; shift 3
  env "$@" timeout ${TMO:-100} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n \
    bench.py --gpus $n --steps 40 --warmup 5 $extra > gpurun_out/n${n}_$name.json 2> gpurun_out/n${n}_$name.err
  echo "== n=$n $name rc=$?"
  python - <<PY
import json
ok=False
for l in open('gpurun_out/n${n}_$name.json'):
    if l.startswith('{'):
        ok=True; j=json.loads(l)
        print('n=$n $name', 'ms/step %.4f'%j['ms_per_step'], {k:round(v,4) for k,v in j['kernel_ms'].items()}, 'launches/step', j.get('gpu_launches_per_step'), 'host', round(j['host_enqueue_ms_per_step'],3), 'parity', j['parity'] and (j['parity']['ok'], round(j['parity']['grad_rel'],5)), j['clocks']['sm_mhz'], 'e2e', j.get('e2e') and round(j['e2e']['ms_per_step'],3))
if not ok: print(open('gpurun_out/n${n}_$name.err').read()[-1200:])
PY
}
for n in ${NS:-8 4 2}; do run $n final "${EXTRA:---no-e2e}" X=1; done
