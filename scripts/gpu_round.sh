#!/bin/bash
# One GPU-box visit: tests, smoke, bench (both paths), ncu launch list + full capture of the main kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 500 --warmup 20 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -c 3000 gpurun_out/bench_fused.json; tail -3 gpurun_out/bench_fused.err
timeout 300 python bench.py --steps 200 --warmup 10 --path materialised > gpurun_out/bench_mat.json 2> gpurun_out/bench_mat.err; tail -c 1500 gpurun_out/bench_mat.json; tail -3 gpurun_out/bench_mat.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fused.csv python bench.py --steps 3 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_mat.csv python bench.py --steps 3 --warmup 3 --path materialised > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'assign_main|assign_lq|pairwise_kernel|match_colmax|match_lq' -s 8 -c 8 -o gpurun_out/prof_targets -f python scripts/profile_targets.py > gpurun_out/ncu_targets.log 2>&1; tail -3 gpurun_out/ncu_targets.log
ls -la gpurun_out | tail -12
