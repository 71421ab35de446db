"""N-rank image sharding on real GPUs: per-image outputs of a 2-rank run are byte-identical
to the 1-rank run, and distributed.gather_detections returns the 1-rank detections.  With >= 2 GPUs the ranks own one
GPU each and the gather runs over NCCL; on a 1-GPU box both ranks share cuda:0 and the group is gloo."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_outputs_byte_identical_to_one_rank(cuda):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "multirank_identity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "IDENTITY OK" in r.stdout, r.stdout[-3000:]
