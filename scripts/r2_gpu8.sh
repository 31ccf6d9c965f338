#!/bin/bash
# Round 2, eight-GPU visit: transport A/B at N = 8 (relay CTA count, NVLS, NCCL), then N = 4 and the full N = 8 line.
#   gpurun --gpus 8 --timeout 600 -- 'bash scripts/r2_gpu8.sh'
mkdir -p gpurun_out
run() {  # n, name, extra bench args, env...
  n=$1; name=$2; extra=$3; shift 3
  env "$@" timeout ${TMO:-120} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n \
    bench.py --gpus $n --steps 40 --warmup 5 $extra > gpurun_out/n${n}_$name.json 2> gpurun_out/n${n}_$name.err
  echo "== n=$n $name rc=$?"
  python - <<PY
import json
ok=False
for l in open('gpurun_out/n${n}_$name.json'):
    if l.startswith('{'):
        ok=True; j=json.loads(l)
        print('n=$n $name', j['config']['transport'], 'ms/step %.4f'%j['ms_per_step'], {k:round(v,4) for k,v in j['kernel_ms'].items()}, 'launches/step', j.get('gpu_launches_per_step'), 'host', round(j['host_enqueue_ms_per_step'],3), 'parity', j['parity'] and (j['parity']['ok'], round(j['parity']['grad_rel'],5)), j['clocks']['sm_mhz'], 'e2e', j.get('e2e') and round(j['e2e']['ms_per_step'],3))
if not ok: print(open('gpurun_out/n${n}_$name.err').read()[-1200:])
PY
}
for v in ${VARIANTS:-relay20 relay12 relay28 nccl nvls}; do
  case $v in
    relay*) run 8 $v --no-e2e VIPANT_TRANSPORT=p2p VIPANT_REQUIRE_P2P=1 VPA_P2P_RELAY_CTAS=${v#relay} ;;
    nccl) run 8 nccl --no-e2e VIPANT_TRANSPORT=nccl ;;
    nvls) TMO=90 run 8 nvls --no-e2e VIPANT_TRANSPORT=p2p VIPANT_REQUIRE_P2P=1 VPA_P2P_MODE=nvls ;;
  esac
done
run 4 default --no-e2e VIPANT_TRANSPORT=p2p
run 8 full "" VIPANT_TRANSPORT=p2p
