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


def _rescale_inplace(d, mean, std):
    """deltas *= std; deltas += mean on the caller's tensor (two roundings, like the reference's two statements)."""
    if not isinstance(d, torch.Tensor) or d.requires_grad or not d.is_floating_point():
        return
    if all(m == 0.0 for m in mean) and all(v == 1.0 for v in std):
        return
    k = d.shape[-1] // 4
    with torch.no_grad():
        d.mul_(torch.tensor(list(std) * k, dtype=d.dtype, device=d.device))
        d.add_(torch.tensor(list(mean) * k, dtype=d.dtype, device=d.device))


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
        """boxcoder.py:75-98: (N,4),(N,4k) -> (N,4k).  Like the reference (:76-77), ``deltas`` is rescaled IN
        PLACE (deltas *= std; deltas += mean) -- for any float dtype and layout, through torch in-place ops so that
        autograd's version counter sees the write.  A tensor that is part of an autograd graph is left untouched (the
        in-place write would corrupt the activations saved for backward); with the identity mean / std nothing is written."""
        d = _plain(deltas)
        out = ops.box_decode(_plain(anchors), d, self.reg_mean, self.reg_std, writeback=False)
        _rescale_inplace(d, self.reg_mean, self.reg_std)
        return out


class SumBoxCoder(BoxCoderBase, metaclass=ABCMeta):
    def __init__(self, reg_mean=(0.0, 0.0, 0.0, 0.0), reg_std=(1.0, 1.0, 1.0, 1.0)):
        self.reg_mean = [float(v) for v in reg_mean]
        self.reg_std = [float(v) for v in reg_std]
        super().__init__()

    def encode(self, anchors, gt):
        return ops.sum_encode(_plain(anchors), _plain(gt), self.reg_mean, self.reg_std)

    def decode(self, anchors, deltas):
        d = _plain(deltas)
        out = ops.sum_decode(_plain(anchors), d, self.reg_mean, self.reg_std, writeback=False)
        _rescale_inplace(d, self.reg_mean, self.reg_std)   # boxcoder.py:123-124, same rule as BoxCoder.decode
        return out


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
