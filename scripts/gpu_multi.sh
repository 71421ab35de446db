#!/bin/bash
# N-GPU visit: NCCL identity test + strong / weak scaling bench lines (N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multirank_identity.py > gpurun_out/multirank_n$N.log 2>&1; echo "identity rc=$?"
tail -8 gpurun_out/multirank_n$N.log
timeout 900 python -m pytest tests/test_gpu_multirank.py -q -m gpu > gpurun_out/pytest_multirank_n$N.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_multirank_n$N.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.loads(open('gpurun_out/bench_n%d.json'%n).read().strip().splitlines()[-1])
        print(n, 'value %.0f ms %.3f e2e %.0f'%(d['value'], d['ms_per_step'], d['e2e']['value']), {k:(round(v['value']), round(v['ms_per_step'],3)) for k,v in d['configs'].items()}, d.get('weak_scaling'))
    except Exception as e: print(n, 'fail', e)
PY
