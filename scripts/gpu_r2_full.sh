#!/bin/bash
# round 2 evidence visit: whole GPU suite, smoke, both bench arms, ncu launch list of the bench command, full captures
# (exported to CSV on the box: gpurun brings back <= 64 MiB)
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $O/r02_smi.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench.json 2> $O/r02_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --simulate-world 8 > $O/r02_bench_shard_of_8.json 2> $O/r02_bench_shard_of_8.err; echo "shard rc=$?"
BDET_ROI_TMA=0 timeout 300 python scripts/perf_roi.py > $O/r02_perf_roi_direct.log 2>&1
BDET_ROI_TMA=1 timeout 300 python scripts/perf_roi.py > $O/r02_perf_roi_tma.log 2>&1
BDET_ROI_TMA=1 BDET_ROI_BWD_TMA_CLS=6 timeout 300 python scripts/perf_roi.py > $O/r02_perf_roi_tma_bwd.log 2>&1
# is it HBM?  the same 8192 rois on ONE image whose 91 MB pyramid stays in L2, no flush between iterations
PERF_B=1 PERF_NOFLUSH=1 timeout 300 python scripts/perf_roi.py > $O/r02_perf_roi_l2resident.log 2>&1
timeout 120 scripts/tma_probe/red_probe > $O/r02_red_probe.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > /dev/null 2> $O/r02_launches_bench.err; echo "launch list rc=$?"
cap() {  # name, kernel regex, config, count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -c $4 -o /tmp/$1 -f python bench.py --steps 1 --warmup 3 --only $3 --eager > $O/r02_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $O/r02_ncu_$1_raw.csv 2>/dev/null
  echo "cap $1 rc=$? $(wc -l < $O/r02_ncu_$1_raw.csv) rows"
}
cap roi 'roi_align_fwd_tma|roi_align_bwd_kernel' c3 2
cap c3 'assign_main|select_sort|sample_labels|nms_chunk|nms_sweep|rcnn_match' c3 16
cap c4 'score_filter|filter_sample|select_sort|nms_fused|nms_sort_small|select_decode' c4 12
cap c2 'assign_main|assign_lq|anchors_grid|count_labels' c2 8
cap c5 'pairwise|match_colmax|nms_chunk|nms_sweep|nms_tile_sort' c5 10
du -sh $O; ls $O | grep r02_ | tr '\n' ' '
