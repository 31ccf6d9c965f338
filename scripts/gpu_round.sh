#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list (+ full capture of the top kernel).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
cat gpurun_out/bench.json gpurun_out/bench_ref.json
if [ -n "$NCU" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"pair_kernel" -s 9 -c 3 -f -o gpurun_out/prof_pair \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
  ncu --set full --clock-control none -k regex:"normalize_pair_kernel|finalize_bwd_kernel" -s 6 -c 2 -f -o gpurun_out/prof_hbm \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_hbm.log 2>&1
fi
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
python scripts/score_bench.py > gpurun_out/scoring.json 2> gpurun_out/scoring.err; echo "score_bench exit $?"
