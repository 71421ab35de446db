#!/bin/bash
# round 2, visit A: new full-size parity tests, multi-rank identity, the reworked bench (both arms)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1800 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_multirank.py -q -m gpu -x -s > gpurun_out/pytest_new.log 2>&1; echo "pytest_new rc=$?"
tail -5 gpurun_out/pytest_new.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
head -c 1500 gpurun_out/bench.json
