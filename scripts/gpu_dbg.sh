mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_targets.py tests/test_gpu_fullsize.py tests/test_gpu_pipelines.py -q -m gpu -x 2>&1 | tail -5
timeout 600 python bench.py --steps 50 --warmup 5 --only c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err
timeout 600 python bench.py --steps 50 --warmup 5 --only c2 --simulate-world 8 > gpurun_out/bench_c2s8.json 2> gpurun_out/bench_c2s8.err
python - <<PY
import json
for f in ('bench_c2','bench_c2s8'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    c=d['configs']['c2']
    print(f, c['images_total'],'img/s %.1f'%c['value'],'ms %.4f'%c['ms_per_step'],'e2e %.1f'%c['e2e']['value'], {k:round(v['avg_us'],1) for k,v in c['kernels'].items()})
PY
