#!/bin/bash
# backward ROIAlign: gather (tile-owner) kernel vs scatter
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_roi.py -q -m gpu -x -k "backward" > gpurun_out/pytest_roi.log 2>&1; rc=$?; echo "pytest_roi rc=$rc"
tail -3 gpurun_out/pytest_roi.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_roi.log; exit 1; fi
timeout 200 python scripts/perf_roi.py 2>&1 | grep -E "bwd" | cut -c1-220
