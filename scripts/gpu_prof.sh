#!/bin/bash
# full ncu capture of the backward pair kernel (2nd launch) + source-level stall table
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"^pair_kernel" -s ${SKIP:-7} -c 1 -f -o gpurun_out/prof_sweep \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"
