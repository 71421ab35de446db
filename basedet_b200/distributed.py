"""Image sharding across GPUs: one process per GPU, no collective on the data path.

Every box op is per image, so rank r simply owns a contiguous slice of the batch -- the same split the reference's
``InferenceSampler`` uses (basedet/data/samplers/inference_sampler.py:26-28).  The only collective is the optional
final all-gather of the (fixed-size, padded) detections for rank-0 evaluation; targets stay on the GPU that owns
the image.  torch.distributed is plumbing here (NCCL on GPUs, gloo in the CPU tests)."""
import math
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local


def shard_range(num_items, rank, world):
    """Contiguous [begin, end) of rank's items: per = ceil(n / world) (inference_sampler.py:26-27)."""
    per = int(math.ceil(num_items / float(world)))
    begin = min(per * rank, num_items)
    return begin, min(per * (rank + 1), num_items)


def shard_batch(batch, rank, world):
    """Slice every tensor / array of a batch dict (``data``, ``gt_boxes``, ``im_info`` ...) along dim 0."""
    n = len(next(iter(batch.values())))
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}, (lo, hi)


def gather_detections(dets, counts, num_images):
    """All-gather fixed-size padded detections.

    dets (b, K, 6) fp32 rows [x1, y1, x2, y2, score, label], counts (b,) int32 valid rows per image, for this rank's
    ``b`` images (shard_range of num_images).  Returns (dets (num_images, K, 6), counts (num_images,)) on every rank,
    in global image order.  With world_size 1 this is the identity."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return dets, counts
    world, rank = dist.get_world_size(), dist.get_rank()
    per = int(math.ceil(num_images / float(world)))
    K = dets.shape[1]
    pad_d = torch.zeros((per, K, 6), dtype=dets.dtype, device=dets.device)
    pad_c = torch.zeros((per,), dtype=counts.dtype, device=counts.device)
    pad_d[: dets.shape[0]] = dets
    pad_c[: counts.shape[0]] = counts
    out_d = [torch.empty_like(pad_d) for _ in range(world)]
    out_c = [torch.empty_like(pad_c) for _ in range(world)]
    dist.all_gather(out_d, pad_d)
    dist.all_gather(out_c, pad_c)
    return torch.cat(out_d)[:num_images], torch.cat(out_c)[:num_images]
