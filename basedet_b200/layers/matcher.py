"""basedet/layers/common/matcher.py:19-51 -- Matcher only (Hungarian / Sinkhorn / OTA matchers are out of scope)."""
from .. import ops

__all__ = ["Matcher"]


class Matcher:
    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        assert len(thresholds) + 1 == len(labels), "thresholds and labels are not matched"
        assert all(low <= high for (low, high) in zip(thresholds[:-1], thresholds[1:]))
        # the reference mutates the caller's list (matcher.py:24-25); kept
        thresholds.append(float("inf"))
        thresholds.insert(0, -float("inf"))
        self.thresholds = thresholds
        self.labels = labels
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, matrix):
        """matrix (G, A) -> (match_indices (A,) int32, labels (A,) int32)."""
        assert len(matrix.shape) == 2
        assert matrix.shape[0] > 0, "Matcher needs at least one row (the reference's max over an empty axis raises)"
        return ops.match(matrix, self.thresholds[1:-1], self.labels, self.allow_low_quality_matches)


class OTATopkMatcher:
    """Drop-in for ``basedet.layers.OTATopkMatcher`` (layers/common/matcher.py:129-161): dynamic-k matching."""

    def __init__(self, candidate_k=10):
        self.candidate_k = candidate_k

    def __call__(self, cost, ious):
        """cost: (#boxes, #anchors) cost matrix; ious: pairwise IoU of gt boxes and anchors, same shape.
        Returns the matched gt index per anchor, ``#boxes`` for unmatched anchors."""
        from .. import ops
        return ops.ota_topk_match(cost, ious, self.candidate_k)
