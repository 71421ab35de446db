mkdir -p gpurun_out
for w in fwd bwd; do DBG_WHAT=$w timeout 300 python scripts/dbg_tma.py 2>&1 | tail -2; done
bash scripts/gpu_r2b.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'roi_align_(fwd|bwd)_tma' -s 2 -c 2 -o gpurun_out/prof_roi_tma -f python scripts/perf_roi.py > gpurun_out/ncu_roi_tma.log 2>&1; tail -2 gpurun_out/ncu_roi_tma.log
