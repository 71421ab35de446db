"""Drop-in for the box-op part of ``basedet.layers`` (layers/common/{anchor_generator,matcher,post_processing,
roi_pool,function}.py).  Same names, arguments, result ordering and assertions; arithmetic runs in libbdet.so."""
from .anchor_generator import (AnchorPointGenerator, BaseAnchorGenerator, DefaultAnchorGenerator,
                               FastPointGenerator, create_anchor_grid)
from .function import is_empty_tensor, meshgrid, non_zeros, permute_to_N_Any_K, safelog
from .matcher import Matcher, OTATopkMatcher
from .post_processing import batched_nms, post_process_with_empty_input, post_processing, py_cpu_nms
from .roi_pool import assign_rois, roi_pool
from .sampling import sample_labels

__all__ = [
    "AnchorPointGenerator", "BaseAnchorGenerator", "DefaultAnchorGenerator", "FastPointGenerator",
    "create_anchor_grid", "is_empty_tensor", "meshgrid", "non_zeros", "permute_to_N_Any_K", "safelog",
    "Matcher", "OTATopkMatcher", "batched_nms", "post_process_with_empty_input", "post_processing", "py_cpu_nms",
    "assign_rois", "roi_pool", "sample_labels",
]
