"""torchrun worker (tests/test_gpu_multirank.py, `gpurun --gpus N`): the N-rank image-sharded run produces, image for
image, the bytes the 1-rank run produces, and `distributed.gather_detections` returns the 1-rank
detections on every rank.  One GPU per rank over NCCL when the box has enough GPUs; otherwise all ranks share cuda:0
and the process group is gloo (the arithmetic is the same; only the gather transport differs).

ROIAlign backward accumulates with fp32 atomics, so dfeat is compared at the 1e-5 gate instead of byte for byte."""
import hashlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from basedet_b200 import benchmarks as BM  # noqa: E402
from basedet_b200 import distributed as D  # noqa: E402


def digest(t):
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()


def main():
    rank, world, local = D.env_world()
    own_gpu = torch.cuda.device_count() >= world
    dev = torch.device("cuda", local if own_gpu else 0)
    torch.cuda.set_device(dev)
    if own_gpu:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    ok = True
    report = []

    # ---- inference side (config 4 shape, 16 images): detections gathered over the process group
    n_img = 16
    lo, hi = D.shard_range(n_img, rank, world)
    dets, cnt = BM.FCOSPostprocess(list(range(lo, hi)), dev).eager()
    if own_gpu:
        g_d, g_c = D.gather_detections(dets, cnt, n_img)           # NCCL all-gather of device tensors
    else:
        g_d, g_c = D.gather_detections(dets.cpu(), cnt.cpu(), n_img)
    full_d, full_c = BM.FCOSPostprocess(list(range(n_img)), dev).eager()
    same = torch.equal(g_d.cpu(), full_d.cpu()) and torch.equal(g_c.cpu(), full_c.cpu())
    report.append("gather_detections[%s] == 1-rank detections: %s" % (dist.get_backend(), same))
    ok &= same

    # ---- training side: per-image digests of every output of config 2 and config 3 (targets stay on the owning GPU)
    def image_digests(arm_cls, images, keys_of):
        arm = arm_cls(images, dev)
        out = arm.eager()
        torch.cuda.synchronize()
        per = {}
        for j, img in enumerate(images):
            per[img] = {k: digest(v) for k, v in keys_of(out, j, arm).items()}
        return per, arm, out

    def c2_keys(out, j, arm):
        lab, idx, off = out
        return {"labels": lab[j], "match_idx": idx[j], "offsets": off[j]}

    def c3_keys(out, j, arm):
        n = 512
        d = {k: out[k][j] for k in ("rois", "n_rois", "rpn_labels", "rpn_targets", "rcnn_rois", "rcnn_labels", "rcnn_targets",
                                     "rcnn_count")}
        d["levels"] = out["levels"][j * n:(j + 1) * n]
        d["pooled"] = out["pooled"][j * n:(j + 1) * n]
        d["rcnn_rois"] = out["rcnn_rois"][j, :, 1:]      # column 0 is the index within the rank's batch
        d["rois"] = out["rois"][j, :, 1:]
        return d

    # at least one image per rank (8 ranks: 8 images)
    for name, arm_cls, n_img, keys in (("c2", BM.RetinaNetTargets, max(8, world), c2_keys),
                                       ("c3", BM.FasterRCNNTrainBoxOps, max(4, world), c3_keys)):
        lo, hi = D.shard_range(n_img, rank, world)
        mine, arm, out = image_digests(arm_cls, list(range(lo, hi)), keys)
        dfe = None
        if name == "c3":
            dfe = [[f[j].cpu() for f in out["dfeats"]] for j in range(hi - lo)]
        del arm, out
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        gd = [None] * world
        dist.all_gather_object(gd, dfe)
        if rank == 0:
            full, arm, out = image_digests(arm_cls, list(range(n_img)), keys)
            merged = {}
            for g in gathered:
                merged.update(g)
            same = merged == full
            report.append("%s: %d images x %d tensors byte-identical across %d ranks: %s" % (name, n_img, len(full[0]), world, same))
            if not same:
                for img in full:
                    for k in full[img]:
                        if merged.get(img, {}).get(k) != full[img][k]:
                            report.append("  differs: image %d %s" % (img, k))
            ok &= same
            if name == "c3":
                flat = [x for g in gd for x in g]
                err = 0.0
                for j in range(n_img):
                    for l, f in enumerate(out["dfeats"]):
                        ref = f[j].cpu()
                        err = max(err, float((flat[j][l] - ref).abs().max() / max(float(ref.abs().max()), 1.0)))
                report.append("c3: dfeat max rel err across ranks %.2e (fp32 atomics: 1e-5 gate)" % err)
                ok &= err <= 1e-5
            del arm, out
        torch.cuda.empty_cache()

    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag.to(dev) if own_gpu else flag, 0)
    if rank == 0:
        print("\n".join(report))
        print("IDENTITY OK" if ok else "IDENTITY FAILED", "world=%d backend=%s gpus=%d" % (world, dist.get_backend(),
                                                                                        torch.cuda.device_count()))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
