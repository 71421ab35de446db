#!/bin/bash
# source-level ncu capture of one ROI kernel inside scripts/perf_roi.py; exports CSV pages, the .ncu-rep stays on the box
# usage: gpu_ncu_src.sh <kernel regex> <tag> [skip]
mkdir -p gpurun_out
K=${1:-roi_align_fwd_tma_kernel}; TAG=${2:-fwd}; SKIP=${3:-4}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o /tmp/prof_$TAG python scripts/perf_roi.py > gpurun_out/ncu_src_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_${TAG}_sass.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source cuda > gpurun_out/ncu_${TAG}_cuda.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page details --csv > gpurun_out/ncu_${TAG}_details.csv 2>/dev/null
ls -la gpurun_out/ncu_${TAG}_*
