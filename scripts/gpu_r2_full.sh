#!/bin/bash
# round 2 evidence visit: whole GPU suite, smoke, both bench arms, ncu launch list of the bench command, full captures
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > /dev/null 2> gpurun_out/r02_launches_bench.err; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'roi_align_fwd_tma|roi_align_bwd_kernel|assign_main|select_sort|sample_labels|nms_chunk|nms_sweep' -c 24 -o gpurun_out/r02_prof_c3 -f python bench.py --steps 1 --warmup 3 --only c3 --eager > gpurun_out/r02_prof_c3.log 2>&1; echo "prof c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'score_filter|filter_sample|select_sort|nms_fused|nms_sort_small' -c 24 -o gpurun_out/r02_prof_c4 -f python bench.py --steps 1 --warmup 3 --only c4 --eager > gpurun_out/r02_prof_c4.log 2>&1; echo "prof c4 rc=$?"
ls -la gpurun_out | grep r02_
