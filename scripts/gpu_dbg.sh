mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -x -k "config3" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --only c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err
timeout 600 python bench.py --steps 20 --warmup 5 --only c3 --simulate-world 8 > gpurun_out/bench_sim8.json 2> gpurun_out/bench_sim8.err; tail -2 gpurun_out/bench_sim8.err
python - <<PY
import json
for f in ('bench_c3','bench_sim8'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    c=d['configs']['c3']
    print(f, c['images_total'],'img/s %.1f'%c['value'],'ms %.4f'%c['ms_per_step'],'e2e %.1f'%c['e2e']['value'])
PY
