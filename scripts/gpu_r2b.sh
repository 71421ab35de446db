#!/bin/bash
# round 2, visit B: TMA ROIAlign kernels: parity, A/B timing, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi_tma.py tests/test_gpu_roi.py -q -m gpu -x > gpurun_out/pytest_roi.log 2>&1; echo "pytest_roi rc=$?"
tail -15 gpurun_out/pytest_roi.log
BDET_ROI_TMA=0 timeout 300 python scripts/perf_roi.py > gpurun_out/perf_roi0.log 2>&1; echo "perf0 rc=$?"
BDET_ROI_TMA=1 timeout 300 python scripts/perf_roi.py > gpurun_out/perf_roi1.log 2>&1; echo "perf1 rc=$?"
cat gpurun_out/perf_roi0.log gpurun_out/perf_roi1.log | tail -12
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_pipelines.py -q -m gpu -x > gpurun_out/pytest_full.log 2>&1; echo "pytest_full rc=$?"
tail -5 gpurun_out/pytest_full.log
timeout 600 python bench.py --steps 20 --warmup 5 --only c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_c3.err
