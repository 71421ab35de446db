#!/bin/bash
# NCCL identity check only (N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multirank_identity.py > gpurun_out/multirank_n$N.log 2>&1; echo "identity rc=$?"
grep -E "gather_detections|byte-identical|dfeat|IDENTITY|differs" gpurun_out/multirank_n$N.log
