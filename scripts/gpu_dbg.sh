mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'roi_align_fwd_tma|roi_align_bwd_kernel' -c 2 -o /tmp/prof_roi_chain -f python bench.py --steps 1 --warmup 3 --only c3 --eager > gpurun_out/prof_roi_chain.log 2>&1; echo "prof rc=$?"
ncu -i /tmp/prof_roi_chain.ncu-rep --page raw --csv > gpurun_out/prof_roi_chain_raw.csv 2>/dev/null
ncu -i /tmp/prof_roi_chain.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_roi_chain_sass.csv 2>/dev/null
ls -la gpurun_out/prof_roi_chain*
