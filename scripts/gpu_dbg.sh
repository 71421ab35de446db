mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pipelines.py tests/test_gpu_postprocess.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_gpu_dropin.py -q -m gpu -x 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 --only c1,c4 > gpurun_out/bench_c14.json 2> gpurun_out/bench_c14.err; tail -2 gpurun_out/bench_c14.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c14.json').read().strip().splitlines()[-1])
for n,c in d['configs'].items():
    print('==',n,c['images_total'],'img/s %.1f'%c['value'],'ms %.4f'%c['ms_per_step'],'launches',c['gpu_launches_per_step'], {k:round(v['avg_us'],1) for k,v in c['kernels'].items()})
PY
