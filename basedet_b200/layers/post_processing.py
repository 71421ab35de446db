"""basedet/layers/common/post_processing.py:17-132."""
from typing import Optional

import numpy as np
import torch

from .. import ops
from ..structures import Boxes, Container

__all__ = ["batched_nms", "post_processing", "py_cpu_nms", "post_process_with_empty_input"]


def batched_nms(boxes, scores, idxs, iou_thresh: float, max_output: Optional[int] = None):
    """Class-aware NMS, post_processing.py:17-47.  Returns int32 indices of the kept boxes, score-descending."""
    assert boxes.ndim == 2 and boxes.shape[1] == 4, "the expected shape of boxes is (N, 4)"
    assert scores.ndim == 1, "the expected shape of scores is (N,)"
    assert idxs.ndim == 1, "the expected shape of idxs is (N,)"
    assert boxes.shape[0] == scores.shape[0] == idxs.shape[0], "number of boxes, scores and idxs are not matched"
    boxes = boxes.as_subclass(torch.Tensor)
    if boxes.shape[0] == 0:
        return torch.zeros((0,), dtype=torch.int32, device=boxes.device)
    keep, cnt = ops.nms_batched(boxes[None], scores[None], idxs.detach()[None], iou_thresh, max_output)
    return keep[0, : int(cnt.item())]  # variable-length result: one D2H read of the count


def post_process_with_empty_input(boxes, box_scores, box_labels, img_info, iou_threshold: float = 0.5,
                                  max_detections_per_image: int = 100):
    """post_processing.py:50-74: per-level lists -> concatenated -> post_processing; [] -> empty Container."""
    if not boxes:
        empty = torch.zeros((0,), dtype=torch.float32, device=img_info.device)
        return Container(boxes=empty, box_scores=empty, box_labels=empty)
    boxes_container = Container(
        boxes=Boxes(torch.cat([b.as_subclass(torch.Tensor) for b in boxes], dim=0)),
        box_scores=torch.cat(box_scores, dim=0),
        box_labels=torch.cat(box_labels, dim=0),
    )
    return post_processing(boxes_container, img_info, iou_threshold=iou_threshold,
                           max_detections_per_image=max_detections_per_image)


def post_processing(boxes_container, img_info, iou_threshold, process_method="nms", max_detections_per_image=None):
    """post_processing.py:78-103: NMS -> gather -> scale to the original image -> clip."""
    keep_idx = batched_nms(boxes_container.boxes, boxes_container.box_scores, boxes_container.box_labels,
                           iou_thresh=iou_threshold, max_output=max_detections_per_image)
    keeped_boxes = boxes_container[keep_idx.long()]
    info = img_info.detach().float().cpu().numpy().astype(np.float32)  # 5 scalars; the reference reads them too
    scale_ratios = (np.float32(info[0, 2] / info[0, 0]), np.float32(info[0, 3] / info[0, 1]))
    kb = keeped_boxes.boxes
    kb = Boxes(kb.as_subclass(torch.Tensor).contiguous())
    kb.scale(scale_ratios).clip((float(info[0, 2]), float(info[0, 3])))
    keeped_boxes.boxes = kb
    return keeped_boxes


def py_cpu_nms(dets: np.ndarray, thresh: float):
    """Host-side greedy NMS on ``dets`` rows [x1, y1, x2, y2, score] (the reference ships the same utility,
    post_processing.py:106-132): highest score first, a box is dropped when its IoU with a kept box exceeds ``thresh``
    (union clamped at 1e-5).  Returns the kept row indices in score order.  Not used by any GPU path of this package."""
    dets = np.asarray(dets)
    corners, scores = dets[:, :4], dets[:, 4]
    area = (corners[:, 2] - corners[:, 0]) * (corners[:, 3] - corners[:, 1])
    order = scores.argsort()[::-1]
    alive = np.ones(order.shape[0], dtype=bool)
    keep = []
    for pos, i in enumerate(order):
        if not alive[pos]:
            continue
        keep.append(i)
        rest = order[pos + 1:]
        wh = np.maximum(np.minimum(corners[i, 2:], corners[rest, 2:]) - np.maximum(corners[i, :2], corners[rest, :2]), 0)
        inter = wh[:, 0] * wh[:, 1]
        iou = inter / np.maximum(area[i] + area[rest] - inter, 1e-5)
        alive[pos + 1:] &= iou <= thresh
    return keep
