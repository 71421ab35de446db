"""Drop-in for ``basedet.structures`` (same names, arguments and result ordering); arithmetic runs in libbdet.so."""
from .box_convert import BoxConverter, BoxMode
from .boxcoder import BoxCoder, BoxCoderBase, PointCoder, SumBoxCoder
from .boxes import Boxes
from .container import Container
from .op_patch import box_center, box_ioa, box_iou, point_distance

__all__ = [
    "BoxConverter", "BoxMode", "BoxCoder", "BoxCoderBase", "PointCoder", "SumBoxCoder", "Boxes", "Container",
    "box_center", "box_ioa", "box_iou", "point_distance",
]
