"""basedet/structures/boxcoder.py:14-141 -- BoxCoder / SumBoxCoder / PointCoder with the reference signatures."""
from abc import ABCMeta, abstractmethod

import torch

from .. import ops

__all__ = ["BoxCoderBase", "BoxCoder", "PointCoder", "SumBoxCoder"]


class BoxCoderBase(metaclass=ABCMeta):
    def __init__(self):
        pass

    @abstractmethod
    def encode(self):
        pass

    @abstractmethod
    def decode(self):
        pass


def _plain(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) else t


class BoxCoder(BoxCoderBase, metaclass=ABCMeta):
    def __init__(self, reg_mean=(0.0, 0.0, 0.0, 0.0), reg_std=(1.0, 1.0, 1.0, 1.0)):
        self.reg_mean = [float(v) for v in reg_mean]
        self.reg_std = [float(v) for v in reg_std]
        assert len(self.reg_mean) == 4 and len(self.reg_std) == 4
        super().__init__()

    def encode(self, bbox, gt):
        """boxcoder.py:61-73: (N,4),(N,4) -> (N,4)."""
        return ops.box_encode(_plain(bbox), _plain(gt), self.reg_mean, self.reg_std)

    def decode(self, anchors, deltas):
        """boxcoder.py:75-98: (N,4),(N,4k) -> (N,4k).  Like the reference, ``deltas`` is rescaled IN PLACE
        (deltas *= std; deltas += mean) when it is a contiguous fp32 tensor."""
        return ops.box_decode(_plain(anchors), _plain(deltas), self.reg_mean, self.reg_std, writeback=True)


class SumBoxCoder(BoxCoderBase, metaclass=ABCMeta):
    def __init__(self, reg_mean=(0.0, 0.0, 0.0, 0.0), reg_std=(1.0, 1.0, 1.0, 1.0)):
        self.reg_mean = [float(v) for v in reg_mean]
        self.reg_std = [float(v) for v in reg_std]
        super().__init__()

    def encode(self, anchors, gt):
        return ops.sum_encode(_plain(anchors), _plain(gt), self.reg_mean, self.reg_std)

    def decode(self, anchors, deltas):
        return ops.sum_decode(_plain(anchors), _plain(deltas), self.reg_mean, self.reg_std, writeback=True)


class PointCoder(BoxCoderBase, metaclass=ABCMeta):
    def encode(self, point, gt):
        """boxcoder.py:132-133.  point (A,2); gt (A,4) -> (A,4), or gt (G,1,4) -> (G,A,4) (fcos.py:231)."""
        point, gt = _plain(point), _plain(gt)
        if gt.ndim == 3:
            assert gt.shape[1] == 1
            return ops.point_encode(point, gt[:, 0, :])
        assert gt.ndim == 2
        if gt.shape[0] == point.shape[0]:
            # element-wise form used after matching (fcos.py:268): row a against gt row a
            enc = ops.point_encode_rows(point, gt)
            return enc
        raise ValueError("PointCoder.encode: gt must be (A,4) or (G,1,4)")

    def decode(self, anchors, deltas):
        return ops.point_decode(_plain(anchors), _plain(deltas))
