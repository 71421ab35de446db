#!/bin/bash
# One GPU-box visit: tests, smoke, bench (both paths + reference arm), bandwidth probe, per-kernel sweep,
# ncu launch list of the bench command and full captures of the main kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 2000 --warmup 20 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -c 4000 gpurun_out/bench_fused.json; tail -3 gpurun_out/bench_fused.err
timeout 300 python bench.py --steps 300 --warmup 10 --path materialised > gpurun_out/bench_mat.json 2> gpurun_out/bench_mat.err; tail -3 gpurun_out/bench_mat.err
timeout 300 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
timeout 200 python scripts/bw_probe.py > gpurun_out/bw_probe.log 2>&1
timeout 600 python scripts/perf_all.py > gpurun_out/perf_all.log 2>&1; tail -3 gpurun_out/perf_all.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_fused.csv python bench.py --steps 5 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_materialised.csv python bench.py --steps 3 --warmup 3 --path materialised > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'assign_main|assign_lq|pairwise_kernel|match_colmax|match_lq' -s 8 -c 8 -o gpurun_out/prof_targets -f python scripts/profile_targets.py > gpurun_out/ncu_targets.log 2>&1; tail -1 gpurun_out/ncu_targets.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_sort|nms_sort_small|select_decode|nms_fused|roi_align|score_filter|nms_chunk|nms_sweep" -s 28 -c 28 -o gpurun_out/prof_post -f python scripts/profile_post.py > gpurun_out/ncu_post.log 2>&1; tail -1 gpurun_out/ncu_post.log
ls -la gpurun_out | tail -20
