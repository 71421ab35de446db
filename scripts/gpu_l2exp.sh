#!/bin/bash
# where is the ROI forward bound?  same 8192 rois on ONE image whose pyramid fits L2, without the flush
mkdir -p gpurun_out
PERF_B=1 PERF_NOFLUSH=1 timeout 120 python scripts/perf_roi.py 2>&1 | grep -E "fwd|bwd" | cut -c1-160 | sed "s/^/B=1 noflush /"
PERF_B=1 timeout 120 python scripts/perf_roi.py 2>&1 | grep -E "fwd|bwd" | cut -c1-160 | sed "s/^/B=1 flush /"
PERF_B=2 PERF_NOFLUSH=1 timeout 120 python scripts/perf_roi.py 2>&1 | grep -E "fwd|bwd" | cut -c1-160 | sed "s/^/B=2 noflush /"
