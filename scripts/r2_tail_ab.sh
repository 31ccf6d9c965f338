#!/bin/bash
# Encoder tail A/B on one B200: single-CTA projection kernel vs the CTA-pair variant (VPA_TAIL_PAIR=1): parity tests, timings,
# ncu launch list (LayerNorm vs projection split) and one full capture of the projection kernel.
#   gpurun --timeout 200 -- 'bash scripts/r2_tail_ab.sh'
mkdir -p gpurun_out
VPA_TAIL_PAIR=1 timeout 60 python -m pytest tests/test_gpu_encoder_tail.py -q -x 2>&1 | tail -4; echo "pair tests exit ${PIPESTATUS[0]}"
VPA_TAIL_PAIR=0 timeout 40 python scripts/encoder_tail_bench.py --rows 32768 4096 2>&1 | tail -3; cp gpurun_out/encoder_tail.json gpurun_out/encoder_tail_single.json
VPA_TAIL_PAIR=1 timeout 40 python scripts/encoder_tail_bench.py --rows 32768 4096 2>&1 | tail -3; cp gpurun_out/encoder_tail.json gpurun_out/encoder_tail_pair.json
for v in 0 1; do
  VPA_TAIL_PAIR=$v timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ln_cast|proj_norm" -c 12 --csv \
      --log-file gpurun_out/tail_launches_$v.csv python scripts/encoder_tail_bench.py --rows 32768 > /dev/null 2>&1
  grep -c proj_norm gpurun_out/tail_launches_$v.csv; tail -4 gpurun_out/tail_launches_$v.csv | cut -d, -f5,12-
done
VPA_TAIL_PAIR=${FULL_VARIANT:-1} timeout 60 ncu --set full --clock-control none --import-source on -k regex:"ln_cast|proj_norm" -s 6 -c 2 -f -o gpurun_out/prof_tail \
    python scripts/encoder_tail_bench.py --rows 32768 > gpurun_out/ncu_tail.log 2>&1; echo "ncu full exit $?"
