mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_postprocess.py tests/test_gpu_pipelines.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -q -m gpu -x 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 --only c1,c4 > gpurun_out/bench_c14.json 2> gpurun_out/bench_c14.err; tail -2 gpurun_out/bench_c14.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c14.json').read().strip().splitlines()[-1])
for n,c in d['configs'].items():
    print('==',n,c['images_total'],'img/s %.1f'%c['value'],'ms %.4f'%c['ms_per_step'],'graph',c['cuda_graph'],'launches',c['gpu_launches_per_step'])
    for k,v in c['kernels'].items():
        print('   %-28s %8.1f us x%5.1f share %.2f %s'%(k,v['avg_us'],v['launches_per_step'],v['share_of_kernel_time'], ('GB/s %.0f frac %.3f'%(v['GBps'],v['frac_of_peak'])) if 'GBps' in v else ''))
PY
