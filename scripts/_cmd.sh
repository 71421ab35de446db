python -m pytest tests/test_gpu_postprocess.py tests/test_gpu_pipelines.py tests/test_gpu_golden.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -5
python scripts/perf_all.py c3_rpn_train_b16 c3_rpn_test_b16 c5_nms_100k_b2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    sp=l.rstrip().split(' ',1)
    if len(sp)==2 and sp[1].startswith('{'):
        d=json.loads(sp[1]); print(sp[0], round(d['total_ms_median'],3), {k.replace('_kernel',''):(round(v['avg_us'],1),v['launches_per_iter']) for k,v in d['kernels'].items()})
    else: print(l.rstrip())
"
