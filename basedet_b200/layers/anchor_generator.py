"""basedet/layers/common/anchor_generator.py:23-182 -- anchors for every level in one kernel launch."""
import math
from abc import ABCMeta, abstractmethod
from functools import cached_property
from typing import List

import numpy as np
import torch

from .. import ops

__all__ = ["AnchorPointGenerator", "BaseAnchorGenerator", "DefaultAnchorGenerator", "FastPointGenerator",
           "create_anchor_grid"]


def create_anchor_grid(featmap_size, offsets, stride, device):
    """anchor_generator.py:23-30 -> (grids_x, grids_y), flattened (H*W,), x varying along W."""
    pts = ops.points_grid([tuple(featmap_size)], [stride], [offsets * stride], 1, 0, device)[0]
    return pts[:, 0].contiguous(), pts[:, 1].contiguous()


class BaseAnchorGenerator(metaclass=ABCMeta):
    def __init__(self):
        pass

    @abstractmethod
    def generate_anchors_by_features(self, sizes, device) -> List[torch.Tensor]:
        pass

    def __call__(self, featmaps):
        feat_sizes = [tuple(fmap.shape[-2:]) for fmap in featmaps]
        return self.generate_anchors_by_features(feat_sizes, featmaps[0].device)

    @property
    def anchor_dim(self):
        return 4


class DefaultAnchorGenerator(BaseAnchorGenerator):
    def __init__(self, anchor_scales: list = [[32], [64], [128], [256], [512]],
                 anchor_ratios: list = [[0.5, 1, 2]], strides: list = [4, 8, 16, 32, 64], offset: float = 0):
        super().__init__()
        self.anchor_scales = np.array(anchor_scales, dtype=np.float32)
        self.anchor_ratios = np.array(anchor_ratios, dtype=np.float32)
        self.strides = strides
        self.offset = offset
        self.num_features = len(strides)

    @cached_property
    def base_anchors(self):
        """Per-level (n, 4) fp32 base anchors (host numpy; float64 math then fp32, anchor_generator.py:95-109)."""
        return self._different_level_anchors(self.anchor_scales.tolist(), self.anchor_ratios.tolist())

    def _different_level_anchors(self, scales, ratios):
        if len(scales) == 1:
            scales = scales * self.num_features
        assert len(scales) == self.num_features
        if len(ratios) == 1:
            ratios = ratios * self.num_features
        assert len(ratios) == self.num_features
        return [np.array(self.generate_base_anchors(s, r), dtype=np.float32).reshape(-1, 4)
                for s, r in zip(scales, ratios)]

    def generate_base_anchors(self, scales, ratios):
        """Cell anchors centred on the origin, scales outer / ratios inner, in float64 like the reference
        (anchor_generator.py:99-109): width = sqrt(scale^2 / ratio), height = ratio * width."""
        def centred(scale, ratio):
            width = math.sqrt(scale ** 2.0 / ratio)
            height = ratio * width
            return [-width / 2.0, -height / 2.0, width / 2.0, height / 2.0]

        return [centred(scale, ratio) for scale in scales for ratio in ratios]

    def _plan(self, sizes):
        key = tuple(tuple(int(v) for v in s) for s in sizes)
        plans = self.__dict__.setdefault("_plans", {})
        if key not in plans:
            shifts = [self.offset * s for s in self.strides]
            plans[key] = ops.AnchorPlan(key, self.strides, shifts, self.base_anchors)
        return plans[key]

    def generate_anchors_by_features(self, sizes, device):
        assert len(sizes) == self.num_features, (
            "input features expected {}, got {}".format(self.num_features, len(sizes))
        )
        return ops.anchors_grid(None, None, None, None, device, plan=self._plan(sizes))

    def generate_all_level_anchors(self, sizes, device):
        """Extension: the level-concatenated (sum, 4) anchors (``F.concat(anchors_list)``, retinanet.py:128) with
        no extra copy -- the per-level tensors are views of this buffer."""
        assert len(sizes) == self.num_features
        return ops.anchors_grid(None, None, None, None, device, flat=True, plan=self._plan(sizes))


class AnchorPointGenerator(BaseAnchorGenerator):
    def __init__(self, num_anchors: int = 1, strides: tuple = (4, 8, 16, 32, 64), offset: float = 0.5):
        super().__init__()
        self.num_anchors = num_anchors
        self.strides = strides
        self.offset = offset
        self.num_features = len(strides)

    @property
    def anchor_dim(self):
        return 2

    def generate_anchors_by_features(self, sizes, device):
        assert len(sizes) == self.num_features, (
            "input features expected {}, got {}".format(self.num_features, len(sizes))
        )
        shifts = [self.offset * s for s in self.strides]
        return ops.points_grid(sizes, list(self.strides), shifts, self.num_anchors, 0, device)


class FastPointGenerator:
    """anchor_generator.py:169-182 (the reference's (w, h)-mesh ordering is kept)."""

    def __init__(self, strides: tuple = (8, 16, 32)):
        self.strides = strides

    def __call__(self, featmaps):
        feats = list(featmaps)[: len(self.strides)]
        sizes = [tuple(f.shape[-2:]) for f in feats]
        strides = list(self.strides)[: len(sizes)]
        return ops.points_grid(sizes, strides, [0.0] * len(sizes), 1, 1, feats[0].device)
