#!/bin/bash
# forward ROIAlign tuning visit: parity, then chunk-size sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi_tma.py tests/test_gpu_roi.py -q -m gpu -x > gpurun_out/pytest_roi.log 2>&1; echo "pytest_roi rc=$?"
tail -3 gpurun_out/pytest_roi.log
for kb in 40 26 20 13; do
  BDET_ROI_CHUNK_KB=$kb timeout 300 python scripts/perf_roi.py 2>&1 | grep -E "fwd" | sed "s/^/kb=$kb /"
done > gpurun_out/perf_roi_sweep.log
cat gpurun_out/perf_roi_sweep.log
timeout 600 python bench.py --steps 20 --warmup 5 --only c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
