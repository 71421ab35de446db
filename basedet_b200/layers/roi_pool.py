"""basedet/layers/common/roi_pool.py:12-78 -- level assignment + multi-level ROIAlign in one launch each."""
import math
from typing import List

import torch

from .. import ops

__all__ = ["roi_pool", "assign_rois"]


def assign_rois(rois, strides):
    """roi_pool.py:12-32 -> (rois_with_dummies, assigned_level) exactly as the reference returns them
    (one zero dummy ROI per level appended).  roi_pool() below does not need the dummies."""
    rois = rois.detach()
    min_level, max_level = int(math.log2(strides[0])), int(math.log2(strides[-1]))
    num_fms = len(strides)
    level = ops.roi_assign_levels(rois, min_level, max_level)
    level = torch.cat([level, torch.arange(num_fms, dtype=torch.int32, device=level.device)])
    rois = torch.cat([rois.float(), torch.zeros((num_fms, rois.shape[-1]), dtype=torch.float32, device=rois.device)])
    return rois, level


class _RoiAlignFn(torch.autograd.Function):
    """MegEngine autodiff of F.nn.roi_align <-> torch.autograd.Function: gradient flows to the features only
    (rois are detached, roi_pool.py:56)."""

    @staticmethod
    def forward(ctx, rois, levels, scales, pool_shape, *features):
        ctx.save_for_backward(rois, levels)
        ctx.scales, ctx.pool_shape = scales, pool_shape
        ctx.shapes = [tuple(f.shape) for f in features]
        return ops.roi_align_fwd(list(features), rois, levels, scales, pool_shape, (2, 2), True)

    @staticmethod
    def backward(ctx, dout):
        rois, levels = ctx.saved_tensors
        grads = ops.roi_align_bwd(dout.contiguous(), ctx.shapes, rois, levels, ctx.scales, ctx.pool_shape, (2, 2), True)
        return (None, None, None, None) + tuple(grads)


class _RoiMaxPoolFn(torch.autograd.Function):
    """F.nn.roi_pooling(mode="max") (roi_pool.py:62-63): the gradient goes to the argmax pixel of every bin."""

    @staticmethod
    def forward(ctx, rois, levels, scales, pool_shape, *features):
        out, argmax = ops.roi_maxpool_fwd(list(features), rois, levels, scales, pool_shape)
        ctx.save_for_backward(rois, levels, argmax)
        ctx.pool_shape = pool_shape
        ctx.shapes = [tuple(f.shape) for f in features]
        return out

    @staticmethod
    def backward(ctx, dout):
        rois, levels, argmax = ctx.saved_tensors
        grads = ops.roi_maxpool_bwd(dout.contiguous(), argmax, ctx.shapes, rois, levels, ctx.pool_shape)
        return (None, None, None, None) + tuple(grads)


def roi_pool(features: List[torch.Tensor], rois: torch.Tensor, strides: List[int], pool_shape,
             pooler_type: str = "roi_align") -> torch.Tensor:
    """features: list of (B, C, H_l, W_l); rois (K, 5) [batch, x1, y1, x2, y2] -> (K, C, PH, PW) in roi order."""
    assert pooler_type in ("roi_align", "roi_pool")
    assert len(strides) == len(features)
    if isinstance(pool_shape, int):
        pool_shape = (pool_shape, pool_shape)
    rois = rois.detach().float().contiguous()
    levels = None  # a single level needs no assignment (every roi clamps to it)
    if len(strides) > 1:
        levels = ops.roi_assign_levels(rois, int(math.log2(strides[0])), int(math.log2(strides[-1])))
    scales = tuple(1.0 / s for s in strides)
    feats = [f.float().contiguous() for f in features]
    fn = _RoiMaxPoolFn if pooler_type == "roi_pool" else _RoiAlignFn
    return fn.apply(rois, levels, scales, tuple(pool_shape), *feats)
