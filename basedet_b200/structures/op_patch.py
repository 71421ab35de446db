"""basedet/structures/op_patch.py:81,116,152,211 -- same four public functions, one CUDA kernel each."""
from .. import _lib, ops

__all__ = ["box_iou", "box_ioa", "box_center", "point_distance"]


def box_iou(boxes1, boxes2):
    """(m,4) x (n,4) -> (m,n) IoU, op order of op_patch.py:33-78."""
    return ops.pairwise(boxes1, boxes2, _lib.PAIR_IOU)


def box_ioa(boxes1, boxes2):
    """(m,4) x (n,4) -> (m,n) intersection over area(boxes2), op_patch.py:169-208."""
    return ops.pairwise(boxes1, boxes2, _lib.PAIR_IOA)


def box_center(boxes):
    """(m,4) -> (m,2) centers, op_patch.py:100-113."""
    return ops.box_center(boxes)


def point_distance(points1, points2):
    """(m,2) x (n,2) -> (m,n) euclidean distance, op_patch.py:133-149."""
    return ops.point_distance(points1, points2)
