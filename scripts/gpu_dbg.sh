mkdir -p gpurun_out
for cls in -1 0 1 2 3 6; do
  BDET_ROI_BWD_TMA_CLS=$cls timeout 300 python bench.py --steps 10 --warmup 3 --only c3 > gpurun_out/bench_c3_cls$cls.json 2> gpurun_out/bench_c3_cls$cls.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c3_cls$cls.json').read().strip().splitlines()[-1])
k=d['configs']['c3']['kernels']
print('cls $cls ms/step %.3f'%d['ms_per_step'], {n:round(v['avg_us']) for n,v in k.items() if 'roi_align' in n})
PY
done
BDET_ROI_BWD_TMA_CLS=1 timeout 300 python scripts/perf_roi.py 2>&1 | grep bwd
BDET_ROI_BWD_TMA_CLS=2 timeout 300 python scripts/perf_roi.py 2>&1 | grep bwd
