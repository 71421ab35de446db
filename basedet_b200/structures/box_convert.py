"""basedet/structures/box_convert.py:10-112."""
from enum import IntEnum, unique

import numpy as np
import torch

from .. import ops


@unique
class BoxMode(IntEnum):
    XYXY = 0
    XYWH = 1
    XcYcWH = 2


class BoxConverter:
    @classmethod
    def convert(cls, boxes, mode="xywh2xyxy"):
        from_mode, to_mode = cls.get_from_mode_and_to_mode(mode)
        if from_mode == to_mode:
            return boxes
        return ops.box_convert(boxes.as_subclass(torch.Tensor), int(from_mode), int(to_mode))

    @classmethod
    def numpy_convert(cls, boxes, mode="xywh2xyxy"):
        t = torch.as_tensor(np.asarray(boxes, dtype=np.float32), device="cuda")
        return np.array(cls.convert(t, mode).cpu().numpy())

    @classmethod
    def get_from_mode_and_to_mode(cls, mode):
        from_mode, to_mode = mode.split("2")
        return cls.parse_mode(from_mode), cls.parse_mode(to_mode)

    @classmethod
    def parse_mode(cls, mode):
        return {"xyxy": BoxMode.XYXY, "xywh": BoxMode.XYWH, "xcycwh": BoxMode.XcYcWH}[mode.lower()]
