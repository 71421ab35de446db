"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of BaseDet's dense box-op hot path.  Nothing under
``basedet_b200/`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` do, and there only as the checker / reported baseline.

Parity status (see DESIGN.md "Oracle"):
  * pinned by the reference's own known-answer tests: IoU, IoA, intersection,
    centers, scale, class-aware NMS keep list, ROIAlign interior values,
    ROIAlign scale equivariance;
  * pinned against the reference's own Python source executed under the numpy
    ``megengine`` shim in ``oracle/mge_shim`` (op ORDER comes from the reference
    files, leaf-op semantics from the shim): Matcher, BoxCoder, PointCoder,
    anchor generators, batched_nms wrapper, roi_pool glue, top-k glue;
  * PARITY UNPINNED (MegEngine itself is absent, leaf semantics are stated
    assumptions): argmax/sort/top-k tie-breaking, NaN handling of MAX/MIN,
    NMS IoU exactly at threshold, ROIAlign border taps, ROIAlign backward.
"""
from . import ref_ops  # noqa: F401
