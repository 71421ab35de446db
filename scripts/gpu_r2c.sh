#!/bin/bash
# forward ROIAlign tuning visit: parity, then configuration A/B (short timeouts: a hang must not eat the budget)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_roi_tma.py tests/test_gpu_roi.py -q -m gpu -x > gpurun_out/pytest_roi.log 2>&1; rc=$?; echo "pytest_roi rc=$rc"
tail -3 gpurun_out/pytest_roi.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_roi.log; exit 1; fi
for cfg in ${CFGS:-1}; do
  BDET_ROI_FWD_CFG=$cfg timeout 120 python scripts/perf_roi.py 2>&1 | grep -E "fwd" | sed "s/^/cfg=$cfg /"
done > gpurun_out/perf_roi_sweep.log
cut -c1-150 gpurun_out/perf_roi_sweep.log
timeout 240 python bench.py --steps 20 --warmup 5 --only c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
