"""CPU, world_size 2 on gloo: the image-sharding / detection-gather host logic of the N>1 path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from basedet_b200 import distributed as D


def test_shard_range_matches_inference_sampler():
    # basedet/data/samplers/inference_sampler.py:26-28: begin = ceil(n/w)*rank, end = min(ceil(n/w)*(rank+1), n)
    for n, w in ((16, 8), (16, 1), (64, 8), (9, 4), (3, 8), (0, 2)):
        got = [D.shard_range(n, r, w) for r in range(w)]
        per = -(-n // w) if n else 0
        assert got == [(min(per * r, n), min(per * (r + 1), n)) for r in range(w)]
        covered = [i for lo, hi in got for i in range(lo, hi)]
        assert covered == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = D.init("gloo")
    assert (r, w) == (rank, world)
    rng = np.random.default_rng(0)
    all_dets = torch.from_numpy(rng.normal(size=(num_images, 5, 6)).astype(np.float32))
    all_cnt = torch.from_numpy(rng.integers(0, 6, num_images).astype(np.int32))
    batch = {"dets": all_dets, "cnt": all_cnt}
    mine, (lo, hi) = D.shard_batch(batch, rank, world)
    assert (lo, hi) == D.shard_range(num_images, rank, world)
    g_d, g_c = D.gather_detections(mine["dets"], mine["cnt"], num_images)
    ok = torch.equal(g_d, all_dets) and torch.equal(g_c, all_cnt)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_gather_detections_world2_gloo():
    ctx = mp.get_context("spawn")
    for num_images in (7, 16):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, num_images, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in procs)
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        assert res == [(0, True), (1, True)]
