"""basedet/structures/boxes.py:10-219 -- ``Boxes`` as a tensor subclass (torch.Tensor stands in for megengine.Tensor)."""
import torch

from .. import _lib, ops
from .op_patch import box_center, box_ioa, box_iou


def _pair(v):
    if isinstance(v, torch.Tensor):
        v = v.tolist()
    if isinstance(v, (int, float)):
        v = (v, v)
    assert len(v) == 2
    return [float(x) if not isinstance(x, torch.Tensor) else float(x.item()) for x in v]


class Boxes(torch.Tensor):
    """(N, 4) xyxy boxes; results of arithmetic are plain tensors, (N, 4) slices stay ``Boxes`` (boxes.py:214-219).

    Value semantics like the reference's: MegEngine tensors are values, ``Boxes(t)`` only shares ``__dict__`` (boxes.py:26),
    and the in-place ``clip`` / ``scale`` (``self[:, 0::2] = ...``, boxes.py:170-177, 206-212) rebind the Boxes object --
    the tensor it was built from keeps its values (``Boxes(proposals).clip(...)`` in rpn.py:168 leaves
    ``proposals`` un-clipped, which is also what the fused ``pipelines.rpn_proposals`` does).  Construction is a
    zero-copy alias; the first in-place operation detaches the Boxes onto private storage (copy on write)."""

    __torch_function__ = torch._C._disabled_torch_function_impl

    @staticmethod
    def __new__(cls, boxes):
        assert isinstance(boxes, torch.Tensor)
        assert boxes.ndim == 2
        assert boxes.shape[1] == 4
        return boxes.as_subclass(cls)

    def _plain(self):
        return self.as_subclass(torch.Tensor)

    @property
    def centers(self):
        return box_center(self._plain())

    @property
    def area(self):
        return ops.box_props(self._plain(), 2)

    @property
    def width(self):
        return ops.box_props(self._plain(), 0)

    @property
    def height(self):
        return ops.box_props(self._plain(), 1)

    def iou(self, boxes):
        return box_iou(self._plain(), torch.Tensor.as_subclass(boxes, torch.Tensor))

    def giou(self, boxes):
        return ops.pairwise(self._plain(), torch.Tensor.as_subclass(boxes, torch.Tensor), _lib.PAIR_GIOU)

    def ioa(self, boxes):
        return box_ioa(self._plain(), torch.Tensor.as_subclass(boxes, torch.Tensor))

    def intersection(self, boxes):
        return ops.pairwise(self._plain(), torch.Tensor.as_subclass(boxes, torch.Tensor), _lib.PAIR_INTER)

    def filter_by_size(self, sizes=0):
        s = _pair(sizes)
        return ops.boxes_filter_by_size(self._plain(), s[0], s[1])

    def _apply(self, sw, sh, cw, ch, inplace):
        src = self._plain()
        if inplace and getattr(self, "_owned", False) and src.is_contiguous() and src.dtype == torch.float32:
            ops.boxes_scale_clip(src, sw, sh, cw, ch)  # already on private storage
            return self
        out = src.float().contiguous().clone()
        ops.boxes_scale_clip(out, sw, sh, cw, ch)
        if inplace:
            self.set_(out)  # rebinds THIS object only; the tensor it was built from is untouched
            self._owned = True
            return self
        return out

    def clip(self, sizes, inplace=True):
        """sizes = (height, width); boxes.py:152-177."""
        h, w = _pair(sizes)
        return self._apply(1.0, 1.0, w, h, inplace)

    def cat(self, boxes, inplace=True):
        out = torch.cat([self._plain(), torch.Tensor.as_subclass(boxes, torch.Tensor)])
        if inplace:  # boxes.py:187-190 assigns into self (shape-changing in MegEngine); return the result instead
            return Boxes(out)
        return out

    def scale(self, scale_ratios, inplace=True):
        """scale_ratios = (scale_h, scale_w); boxes.py:193-212."""
        sh, sw = _pair(scale_ratios)
        return self._apply(sw, sh, -1.0, -1.0, inplace)

    def __getitem__(self, idx):
        out = self._plain()[idx]
        if out.ndim == 2 and out.shape[1] == 4:
            out = Boxes(out)
        return out
