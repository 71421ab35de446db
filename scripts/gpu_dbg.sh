mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'roi_align_fwd_tma' -s 1 -c 1 -o gpurun_out/prof_roi_fwd_tma2 -f python scripts/perf_roi.py > gpurun_out/ncu_roi_fwd_tma2.log 2>&1; tail -1 gpurun_out/ncu_roi_fwd_tma2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'roi_align_bwd_tma' -s 1 -c 1 -o gpurun_out/prof_roi_bwd_tma2 -f python scripts/perf_roi.py > gpurun_out/ncu_roi_bwd_tma2.log 2>&1; tail -1 gpurun_out/ncu_roi_bwd_tma2.log
