#!/bin/bash
# Round 2, one-GPU visit: smoke, parity tests (incl. the peer-memory transport between processes sharing the GPU), bench,
# reference arm, scoring bench, ncu launch list + full captures of the scoring / normalise kernels.
#   gpurun --timeout 1500 -- 'bash scripts/r2_gpu1.sh'
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 1100 python -m pytest tests -q -m gpu --timeout 300 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 500 python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
cat gpurun_out/bench.json gpurun_out/bench_ref.json; tail -3 gpurun_out/bench.err
timeout 300 python scripts/score_bench.py > gpurun_out/scoring.json 2> gpurun_out/scoring.err; echo "score_bench exit $?"; cat gpurun_out/scoring.json
if [ -n "$NCU" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-eager > gpurun_out/ncu_bench.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sim_rank_tile|sim_gt_ref|normalize_pair|ap_class|micro_ap_kernel|finalize_multi|rank_topk" -c 14 -f -o gpurun_out/prof_scoring \
      python scripts/score_bench.py > gpurun_out/ncu_scoring.log 2>&1
fi
