"""F.vision.nms (ASSUMED-3 / ASSUMED-5)."""
import numpy as np

from ..tensor import Tensor


def nms(boxes, scores, iou_thresh, max_output=None):
    b = boxes._a.astype(np.float32)
    s = scores._a.astype(np.float32)
    n = b.shape[0]
    order = np.argsort(-s, kind="stable")
    sb = b[order]
    area = ((sb[:, 2] - sb[:, 0]) * (sb[:, 3] - sb[:, 1])).astype(np.float32)
    removed = np.zeros(n, dtype=bool)
    thr = np.float32(iou_thresh)
    keep = []
    zero = np.float32(0)
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        if max_output is not None and len(keep) >= max_output:
            break
        r = sb[i + 1:]
        w = np.maximum((np.minimum(sb[i, 2], r[:, 2]) - np.maximum(sb[i, 0], r[:, 0])).astype(np.float32), zero)
        h = np.maximum((np.minimum(sb[i, 3], r[:, 3]) - np.maximum(sb[i, 1], r[:, 1])).astype(np.float32), zero)
        inter = (w * h).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = (inter / ((area[i] + area[i + 1:]).astype(np.float32) - inter).astype(np.float32)).astype(np.float32)
        removed[i + 1:] |= iou > thr
    return Tensor(order[np.asarray(keep, dtype=np.int64)].astype(np.int32))


def interpolate(inp, size=None, scale_factor=None, mode="bilinear", align_corners=None):
    import torch
    import torch.nn.functional as TF

    out = TF.interpolate(torch.from_numpy(np.ascontiguousarray(inp._a)), size=size, scale_factor=scale_factor, mode=mode,
                         align_corners=bool(align_corners) if mode != "nearest" else None)
    return Tensor(out.numpy())
