"""Functional host layer: torch tensors in, raw device pointers into libbdet.so, torch tensors out.

PyTorch is used for device memory and streams only; every arithmetic result comes from the CUDA
library.  CPU tensors are rejected -- there is no fallback path.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, darr, farr, iarr, larr

_VP = ctypes.c_void_p


def as_tensor(x):
    """Tensor adapter: torch.Tensor, anything with __dlpack__ (MegEngine >= 1.9 tensors, cupy), or
    __cuda_array_interface__ (numba / cupy / MegEngine through a one-line shim)."""
    if isinstance(x, torch.Tensor):
        return x
    if hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    if hasattr(x, "__cuda_array_interface__"):
        return torch.as_tensor(x, device="cuda")
    raise TypeError("expected a device tensor (torch / DLPack / __cuda_array_interface__), got %r" % type(x))


def _dev(t, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError("%s must live on a CUDA device: basedet_b200 has no CPU path" % name)
    return t


def _f32(t, name="tensor"):
    t = _dev(as_tensor(t), name)
    if t.dtype != torch.float32:
        t = t.float()
    return t


def _f32c(t, name="tensor"):
    return _f32(t, name).contiguous()


def _i32c(t, name="tensor"):
    t = _dev(as_tensor(t), name)
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()


def _rows(t, name="boxes"):
    """(N, >=4) fp32 view -> (tensor, ld): keeps row-strided views such as gt[:, :4] without a copy."""
    t = _f32(t, name)
    assert t.ndim == 2 and t.shape[1] >= 4
    if t.stride(1) == 1 and (t.shape[0] <= 1 or t.stride(0) >= t.shape[1]):
        return t, (t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 4))
    t = t.contiguous()
    return t, t.shape[1]


def _p(t):
    return _VP(t.data_ptr()) if t is not None else None


def _stream(t=None):
    dev = t.device if t is not None else None
    return _VP(torch.cuda.current_stream(dev).cuda_stream)


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _guard(t):
    """Make t's device current for the call; a no-op (no cudaSetDevice round trip) when it already is."""
    if t.device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(t.device)


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------- anchors
class AnchorPlan:
    """Host-side arguments of bdet_anchors_grid, built once per (feature sizes, generator) and reused every step
    (the reference regenerates anchors every forward, retinanet.py:116; only the launch should cost)."""

    def __init__(self, sizes, strides, shifts, base_anchors):
        self.n = len(sizes)
        counts = [int(h) * int(w) * len(b) for (h, w), b in zip(sizes, base_anchors)]
        self.offs = [0]
        for c in counts:
            self.offs.append(self.offs[-1] + c)
        self.hw = iarr([int(v) for hw_ in sizes for v in hw_])
        self.strides = darr(strides)
        self.shifts = darr(shifts)
        self.n_base = iarr([len(b) for b in base_anchors])
        self.base = farr([float(v) for b in base_anchors for row in b for v in row])
        self.out_off = larr(self.offs[:-1])


def anchors_grid(sizes, strides, shifts, base_anchors, device, flat=False, plan=None):
    """All levels in one launch.  base_anchors: list (per level) of (n_base, 4) float arrays (host).
    Returns the per-level list (views of one buffer), or that (sum, 4) buffer itself when ``flat``."""
    lib = _lib.load()
    if plan is None:
        plan = AnchorPlan(sizes, strides, shifts, base_anchors)
    out = torch.empty((plan.offs[-1], 4), dtype=torch.float32, device=device)
    with _guard(out):
        check(lib.bdet_anchors_grid(_p(out), plan.n, plan.hw, plan.strides, plan.shifts, plan.n_base, plan.base,
                                    plan.out_off, _stream(out)))
    if flat:
        return out
    return [out[plan.offs[i]:plan.offs[i + 1]] for i in range(plan.n)]


def points_grid(sizes, strides, shifts, num_anchors, mode, device):
    lib = _lib.load()
    n = len(sizes)
    rep = num_anchors if mode == 0 else 1
    counts = [int(h) * int(w) * rep for (h, w) in sizes]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    out = torch.empty((offs[-1], 2), dtype=torch.float32, device=device)
    hw = [int(v) for hw_ in sizes for v in hw_]
    with _guard(out):
        check(lib.bdet_points_grid(_p(out), n, iarr(hw), darr(strides), darr(shifts), int(num_anchors), int(mode),
                                   larr(offs[:-1]), _stream(out)))
    return [out[offs[i]:offs[i + 1]] for i in range(n)]


# ----------------------------------------------------------------------------- pairwise
def _padded_rows(shape_prefix, M, device):
    """(..., M) fp32 view whose rows start 16-byte aligned (row stride rounded up to 4 floats): lets the kernels
    use 128-bit stores / loads.  Returns (view, row stride)."""
    ldo = (M + 3) // 4 * 4
    buf = torch.empty(tuple(shape_prefix) + (ldo,), dtype=torch.float32, device=device)
    return buf[..., :M], ldo


def pairwise(boxes1, boxes2, mode=_lib.PAIR_IOU):
    """(N, >=4) x (M, >=4) -> (N, M).  The result is a view of a row-padded buffer when M % 4 != 0."""
    lib = _lib.load()
    b1, ld1 = _rows(boxes1, "boxes1")
    b2, ld2 = _rows(boxes2, "boxes2")
    N, M = b1.shape[0], b2.shape[0]
    out, ldo = _padded_rows((N,), M, b1.device)
    with _guard(out):
        check(lib.bdet_pairwise(_p(b1), ld1, N, _p(b2), ld2, M, _p(out), ldo, mode, _stream(out)))
    return out


def pairwise_batched(gt, num_gt, anchors, mode=_lib.PAIR_IOU, out=None):
    """gt (B, Gmax, >=4) contiguous, num_gt (B,) int32 or None, anchors (A, 4) shared by the batch or (B, A, 4) per
    image -> (B, Gmax, A).  Rows >= num_gt[b] are left untouched.  ``out`` may be a view made by ``_padded_rows``."""
    lib = _lib.load()
    gt = _f32c(gt, "gt")
    anchors = _f32c(anchors, "anchors")
    B, Gmax, ld = gt.shape
    A = anchors.shape[-2]
    assert anchors.shape[-1] == 4 and (anchors.ndim == 2 or anchors.shape[0] == B)
    bs2 = A * 4 if anchors.ndim == 3 else 0
    if out is None:
        out, _ = _padded_rows((B, Gmax), A, gt.device)
    assert out.shape == (B, Gmax, A) and out.stride(2) == 1 and out.stride(0) == Gmax * out.stride(1)
    n1 = _i32c(num_gt, "num_gt") if num_gt is not None else None
    with _guard(out):
        check(lib.bdet_pairwise_batched(_p(gt), ld, Gmax * ld, _p(n1), Gmax, _p(anchors), 4, bs2, A, _p(out),
                                        out.stride(1), out.stride(0), B, mode, _stream(out)))
    return out


def box_center(boxes):
    lib = _lib.load()
    b, ld = _rows(boxes)
    out = torch.empty((b.shape[0], 2), dtype=torch.float32, device=b.device)
    with _guard(out):
        check(lib.bdet_box_center(_p(b), ld, b.shape[0], _p(out), _stream(out)))
    return out


def point_distance(p1, p2):
    lib = _lib.load()
    p1, p2 = _f32c(p1), _f32c(p2)
    out = torch.empty((p1.shape[0], p2.shape[0]), dtype=torch.float32, device=p1.device)
    with _guard(out):
        check(lib.bdet_point_distance(_p(p1), p1.shape[0], _p(p2), p2.shape[0], _p(out), _stream(out)))
    return out


# ----------------------------------------------------------------------------- matcher
def match(matrix, thresholds, labels, allow_low_quality=False, num_g=None):
    """matrix (G, A) or (B, Gmax, A), rows may be padded (stride >= A).  Returns (match_idx, labels) int32 of shape
    (A,) / (B, A)."""
    lib = _lib.load()
    m = _f32(matrix, "matrix")
    squeeze = m.ndim == 2
    if squeeze:
        m = m.unsqueeze(0)
    B, G, A = m.shape
    if not (m.stride(2) == 1 and m.stride(1) >= A and (B == 1 or m.stride(0) >= G * m.stride(1))) and m.numel():
        m = m.contiguous()
    ld = m.stride(1) if G > 1 else max(A, 1)
    bs = m.stride(0) if B > 1 else G * ld
    idx = torch.empty((B, A), dtype=torch.int32, device=m.device)
    lab = torch.empty((B, A), dtype=torch.int32, device=m.device)
    ws_bytes = lib.bdet_match_workspace(G, A, B)
    ws = _workspace(ws_bytes, m.device)
    gd = _i32c(num_g) if num_g is not None else None
    with _guard(m):
        check(lib.bdet_match(_p(m), ld, bs, _p(gd), G, A, B, farr(thresholds), iarr(labels), len(labels),
                             int(bool(allow_low_quality)), _p(idx), _p(lab), _p(ws), ws.numel(), _stream(m)))
    if squeeze:
        return idx[0], lab[0]
    return idx, lab


def match_rows(matrix):
    lib = _lib.load()
    m = _f32c(matrix, "matrix")
    R, G = m.shape
    mx = torch.empty((R,), dtype=torch.float32, device=m.device)
    am = torch.empty((R,), dtype=torch.int32, device=m.device)
    with _guard(m):
        check(lib.bdet_match_rows(_p(m), R, G, _p(mx), _p(am), _stream(m)))
    return mx, am


# ----------------------------------------------------------------------------- coders
def box_encode(bbox, gt, mean, std, gather_idx=None):
    lib = _lib.load()
    bbox = _f32c(bbox, "bbox")
    N = bbox.shape[0]
    if gather_idx is not None:
        g, ld = _rows(gt, "gt")
        gi = _i32c(gather_idx)
    else:
        g, ld, gi = _f32c(gt, "gt"), 4, None
        assert g.shape == bbox.shape
    out = torch.empty((N, 4), dtype=torch.float32, device=bbox.device)
    with _guard(out):
        check(lib.bdet_box_encode(_p(bbox), _p(g), ld, _p(gi), N, farr(mean), farr(std), _p(out), _stream(out)))
    return out


def box_decode(anchors, deltas, mean, std, writeback=False, sel_idx=None, sel_div=1):
    """deltas must be fp32 contiguous for the in-place write-back to reach the caller's tensor."""
    lib = _lib.load()
    anchors = _f32c(anchors, "anchors")
    d = _f32(deltas, "deltas")
    if not d.is_contiguous():
        d = d.contiguous()
    N = anchors.shape[0]
    k = d.shape[1] // 4
    assert d.shape[0] == N and d.shape[1] == 4 * k
    if sel_idx is not None:
        si = _i32c(sel_idx)
        out = torch.empty((si.numel(), 4), dtype=torch.float32, device=anchors.device)
        nsel = si.numel()
    else:
        si, nsel = None, 0
        out = torch.empty((N, 4 * k), dtype=torch.float32, device=anchors.device)
    with _guard(out):
        check(lib.bdet_box_decode(_p(anchors), _p(d), N, k, farr(mean), farr(std), _p(out), int(bool(writeback)),
                                  _p(si), nsel, int(sel_div), _stream(out)))
    return out


def sum_encode(anchors, gt, mean, std):
    lib = _lib.load()
    a, g = _f32c(anchors), _f32c(gt)
    out = torch.empty_like(a)
    with _guard(out):
        check(lib.bdet_sum_encode(_p(a), _p(g), a.shape[0], farr(mean), farr(std), _p(out), _stream(out)))
    return out


def sum_decode(anchors, deltas, mean, std, writeback=False):
    lib = _lib.load()
    a = _f32c(anchors)
    d = _f32(deltas)
    if not d.is_contiguous():
        d = d.contiguous()
    out = torch.empty_like(a)
    with _guard(out):
        check(lib.bdet_sum_decode(_p(a), _p(d), a.shape[0], farr(mean), farr(std), _p(out), int(bool(writeback)),
                                  _stream(out)))
    return out


def point_encode(points, gt):
    """points (A, 2), gt (G, >=4) -> (G, A, 4)."""
    lib = _lib.load()
    p = _f32c(points)
    g, ld = _rows(gt, "gt")
    out = torch.empty((g.shape[0], p.shape[0], 4), dtype=torch.float32, device=p.device)
    with _guard(out):
        check(lib.bdet_point_encode(_p(p), p.shape[0], _p(g), ld, g.shape[0], _p(out), _stream(out)))
    return out


def point_encode_rows(points, gt):
    """points (N, 2), gt (N, >=4) -> (N, 4): row a against gt row a."""
    lib = _lib.load()
    p = _f32c(points)
    g, ld = _rows(gt, "gt")
    assert g.shape[0] == p.shape[0]
    out = torch.empty((p.shape[0], 4), dtype=torch.float32, device=p.device)
    with _guard(out):
        check(lib.bdet_point_encode_rows(_p(p), _p(g), ld, p.shape[0], _p(out), _stream(out)))
    return out


def point_decode(points, deltas, sel_idx=None, sel_div=1):
    lib = _lib.load()
    p, d = _f32c(points), _f32c(deltas)
    N, k = p.shape[0], d.shape[1] // 4
    if sel_idx is not None:
        si = _i32c(sel_idx)
        out = torch.empty((si.numel(), 4), dtype=torch.float32, device=p.device)
        nsel = si.numel()
    else:
        si, nsel = None, 0
        out = torch.empty_like(d)
    with _guard(out):
        check(lib.bdet_point_decode(_p(p), _p(d), N, k, _p(out), _p(si), nsel, int(sel_div), _stream(out)))
    return out


def boxes_scale_clip(boxes, scale_w, scale_h, clip_w=-1.0, clip_h=-1.0):
    """In place on a contiguous fp32 (N, 4) tensor."""
    lib = _lib.load()
    assert boxes.is_cuda and boxes.dtype == torch.float32 and boxes.is_contiguous()
    with _guard(boxes):
        check(lib.bdet_boxes_scale_clip(_p(boxes), boxes.shape[0], float(scale_w), float(scale_h), float(clip_w),
                                        float(clip_h), _stream(boxes)))
    return boxes


def boxes_filter_by_size(boxes, size0=0.0, size1=0.0):
    lib = _lib.load()
    b = _f32c(boxes)
    keep = torch.empty((b.shape[0],), dtype=torch.uint8, device=b.device)
    with _guard(b):
        check(lib.bdet_boxes_filter_by_size(_p(b), b.shape[0], float(size0), float(size1), _p(keep), _stream(b)))
    return keep.bool()


# ----------------------------------------------------------------------------- fused target assignment
class AssignPlan:
    """Pre-allocated outputs + workspace for bdet_assign_targets (reused across steps)."""

    def __init__(self, A, Gmax, B, device):
        lib = _lib.load()
        self.A, self.Gmax, self.B = A, Gmax, B
        self.labels = torch.empty((B, A), dtype=torch.int32, device=device)
        self.idx = torch.empty((B, A), dtype=torch.int32, device=device)
        self.offsets = torch.empty((B, A, 4), dtype=torch.float32, device=device)
        self.ws = _workspace(lib.bdet_assign_targets_workspace(Gmax, A, B), device)


def assign_targets(anchors, gt_boxes, num_gt, thresholds, labels, allow_low_quality=True, apply_class=True,
                   mean=(0, 0, 0, 0), std=(1, 1, 1, 1), plan=None):
    """anchors (A,4); gt_boxes (B,Gmax,5); num_gt (B,) -> (labels (B,A), match_idx (B,A), offsets (B,A,4))."""
    lib = _lib.load()
    anchors = _f32c(anchors, "anchors")
    gt = _f32c(gt_boxes, "gt_boxes")
    assert gt.ndim == 3 and gt.shape[2] == 5, "gt_boxes must be (B, Gmax, 5)"
    B, Gmax, _ = gt.shape
    A = anchors.shape[0]
    ng = _i32c(num_gt, "num_gt")
    if plan is None:
        plan = AssignPlan(A, Gmax, B, anchors.device)
    assert (plan.A, plan.Gmax, plan.B) == (A, Gmax, B)
    with _guard(anchors):
        check(lib.bdet_assign_targets(_p(anchors), A, _p(gt), Gmax, _p(ng), B, farr(thresholds), iarr(labels),
                                      len(labels), int(bool(allow_low_quality)), int(bool(apply_class)),
                                      farr(mean), farr(std), _p(plan.labels), _p(plan.idx), _p(plan.offsets),
                                      _p(plan.ws), plan.ws.numel(), _stream(anchors)))
    return plan.labels, plan.idx, plan.offsets


def assign_targets_grid(anchor_plan, gt_boxes, num_gt, thresholds, labels, allow_low_quality=True, apply_class=True,
                        mean=(0, 0, 0, 0), std=(1, 1, 1, 1), plan=None, counts=None):
    """``assign_targets`` with the anchors generated inside the kernels from an ``AnchorPlan`` (the generator's grid
    description) instead of read from a tensor; ``counts`` (B,3) int32, optional, receives the label census.
    -> (labels (B,A), match_idx (B,A), offsets (B,A,4))."""
    lib = _lib.load()
    gt = _f32c(gt_boxes, "gt_boxes")
    assert gt.ndim == 3 and gt.shape[2] == 5, "gt_boxes must be (B, Gmax, 5)"
    B, Gmax, _ = gt.shape
    A = anchor_plan.offs[-1]
    ng = _i32c(num_gt, "num_gt")
    if plan is None:
        plan = AssignPlan(A, Gmax, B, gt.device)
    assert (plan.A, plan.Gmax, plan.B) == (A, Gmax, B)
    assert counts is None or (counts.shape == (B, 3) and counts.dtype == torch.int32 and counts.is_contiguous())
    with _guard(gt):
        check(lib.bdet_assign_targets_grid(anchor_plan.n, anchor_plan.hw, anchor_plan.strides, anchor_plan.shifts,
                                           anchor_plan.n_base, anchor_plan.base, _p(gt), Gmax, _p(ng), B, farr(thresholds),
                                           iarr(labels), len(labels), int(bool(allow_low_quality)), int(bool(apply_class)),
                                           farr(mean), farr(std), _p(plan.labels), _p(plan.idx), _p(plan.offsets), _p(counts),
                                           _p(plan.ws), plan.ws.numel(), _stream(gt)))
    return plan.labels, plan.idx, plan.offsets


class DensePlan:
    """Pre-allocated outputs (+ ATSS workspace) of the anchor-free target assignment, reused across steps."""

    def __init__(self, A, B, device, atss=False):
        lib = _lib.load()
        self.A, self.B = A, B
        self.labels = torch.empty((B, A), dtype=torch.int32, device=device)
        self.offsets = torch.empty((B, A, 4), dtype=torch.float32, device=device)
        self.ctrness = torch.empty((B, A), dtype=torch.float32, device=device)
        self.idx = torch.empty((B, A), dtype=torch.int32, device=device)
        self.ws = _workspace(lib.bdet_atss_targets_workspace(A, B), device) if atss else None


def _dense_inputs(points_list, gt_boxes, num_gt):
    pts = [_f32c(p, "points") for p in points_list]
    assert all(p.ndim == 2 and p.shape[1] == 2 for p in pts), "points must be (n_l, 2)"
    starts = [0]
    for p in pts:
        starts.append(starts[-1] + p.shape[0])
    flat = pts[0] if len(pts) == 1 else torch.cat(pts, dim=0)
    gt = _f32c(gt_boxes, "gt_boxes")
    assert gt.ndim == 3 and gt.shape[2] == 5, "gt_boxes must be (B, Gmax, 5)"
    return flat, starts, gt, _i32c(num_gt, "num_gt")


def fcos_targets(points_list, gt_boxes, num_gt, strides, sizes_of_interest, center_sampling_radius, plan=None):
    """FCOS.get_ground_truth (models/det/fcos.py:222-293) for the whole batch, fused.
    points_list: L tensors (n_l, 2); gt_boxes (B, Gmax, 5); num_gt (B,).
    -> labels (B, A) int32, offsets (B, A, 4), ctrness (B, A), match_idx (B, A)."""
    lib = _lib.load()
    flat, starts, gt, ng = _dense_inputs(points_list, gt_boxes, num_gt)
    A, (B, Gmax, _) = flat.shape[0], gt.shape
    L = len(points_list)
    assert len(strides) == L and len(sizes_of_interest) == L
    if plan is None:
        plan = DensePlan(A, B, flat.device)
    assert (plan.A, plan.B) == (A, B)
    radius = [float(s * center_sampling_radius) for s in strides] if center_sampling_radius > 0 else [0.0] * L
    with _guard(flat):
        check(lib.bdet_fcos_targets(_p(flat), A, iarr(starts), L, farr(radius), farr([s[0] for s in sizes_of_interest]),
                                    farr([s[1] for s in sizes_of_interest]), _p(gt), Gmax, _p(ng), B, _p(plan.labels),
                                    _p(plan.offsets), _p(plan.ctrness), _p(plan.idx), _stream(flat)))
    return plan.labels, plan.offsets, plan.ctrness, plan.idx


def atss_targets(points_list, gt_boxes, num_gt, strides, anchor_scale=8, topk=9, plan=None):
    """ATSS.get_ground_truth (models/det/atss.py:17-86) for the whole batch, fused.  Same returns as fcos_targets."""
    lib = _lib.load()
    flat, starts, gt, ng = _dense_inputs(points_list, gt_boxes, num_gt)
    A, (B, Gmax, _) = flat.shape[0], gt.shape
    L = len(points_list)
    assert len(strides) == L
    if plan is None:
        plan = DensePlan(A, B, flat.device, atss=True)
    assert (plan.A, plan.B) == (A, B) and plan.ws is not None
    half = [float(s * anchor_scale / 2) for s in strides]  # atss.py:33-34, evaluated in Python floats
    with _guard(flat):
        check(lib.bdet_atss_targets(_p(flat), A, iarr(starts), L, farr(half), int(topk), _p(gt), Gmax, _p(ng), B,
                                    _p(plan.labels), _p(plan.offsets), _p(plan.ctrness), _p(plan.idx), _p(plan.ws),
                                    plan.ws.numel(), _stream(flat)))
    return plan.labels, plan.offsets, plan.ctrness, plan.idx


# ----------------------------------------------------------------------------- score filter + top-k
def _segments(tensors, per_image_len=None):
    """Describe (image, level) segments that live in several tensors as element offsets from one base pointer.

    tensors: list of L fp32 CUDA tensors, each (B, n_l) contiguous (flattened trailing dims).  Segment s = b*L + l
    (image-major) covers tensors[l][b].  Returns (base_tensor, starts, lens)."""
    base = min(tensors, key=lambda t: t.data_ptr())
    starts, lens = [], []
    B = tensors[0].shape[0]
    for b in range(B):
        for t in tensors:
            n = t[0].numel()
            diff = t.data_ptr() - base.data_ptr()
            assert diff % 4 == 0
            starts.append(diff // 4 + b * n)
            lens.append(n)
    return base, starts, lens


def topk_segments(scores, seg_lengths, k):
    """scores: flat fp32 tensor holding the segments back to back.  Returns (vals (S,k), idx (S,k), count (S,)):
    per segment the min(k, n_s) largest scores sorted by (score desc, index asc); idx is within the segment."""
    s = _f32c(scores, "scores").reshape(-1)
    starts = [0]
    for n in seg_lengths[:-1]:
        starts.append(starts[-1] + int(n))
    assert sum(int(n) for n in seg_lengths) == s.numel()
    return topk_raw(s, starts if seg_lengths else [], list(seg_lengths), k)


def topk_raw(base, starts, lens, k, out=None, workspace=None):
    """Low-level form: segments given as element offsets from ``base`` (see ``_segments``)."""
    lib = _lib.load()
    S = len(lens)
    dev = base.device
    if out is None:
        out = (torch.empty((S, k), dtype=torch.float32, device=dev), torch.empty((S, k), dtype=torch.int32, device=dev),
               torch.empty((S,), dtype=torch.int32, device=dev))
    vals, idx, cnt = out
    need = lib.bdet_topk_workspace(sum(int(n) for n in lens), S, k)
    ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, dev)
    with _guard(base):
        check(lib.bdet_topk(_p(base), larr(starts), larr(lens), S, int(k), _p(vals), _p(idx), _p(cnt), _p(ws), ws.numel(),
                            _stream(base)))
    return vals, idx, cnt


def score_filter_topk(logits, seg_lengths, threshold, k, mode=_lib.SCORE_SIGMOID, ctrness=None, num_classes=1,
                      workspace=None):
    """Fused score -> (score > threshold) -> top-k per segment (retinanet.py:181-191 / fcos.py:194-202).
    logits flat fp32 (segments back to back); ctrness (FCOS) has one value per `num_classes` logits.
    Returns (scores (S,k), idx (S,k) flat index within the segment, count (S,))."""
    lg = _f32c(logits, "logits").reshape(-1)
    starts = [0]
    for n in seg_lengths[:-1]:
        starts.append(starts[-1] + int(n))
    assert sum(int(n) for n in seg_lengths) == lg.numel()
    ct = _f32c(ctrness, "ctrness").reshape(-1) if ctrness is not None else None
    return score_filter_topk_raw(lg, starts if seg_lengths else [], list(seg_lengths), threshold, k, mode, ct, None,
                                 num_classes, workspace=workspace)


def score_filter_topk_raw(base, starts, lens, threshold, k, mode, ctr_base=None, ctr_starts=None, num_classes=1,
                          out=None, workspace=None):
    lib = _lib.load()
    S = len(lens)
    dev = base.device
    if out is None:
        out = (torch.empty((S, k), dtype=torch.float32, device=dev), torch.empty((S, k), dtype=torch.int32, device=dev),
               torch.empty((S,), dtype=torch.int32, device=dev))
    vals, idx, cnt = out
    need = lib.bdet_score_filter_topk_workspace(sum(int(n) for n in lens), S, k)
    ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, dev)
    with _guard(base):
        check(lib.bdet_score_filter_topk(_p(base), _p(ctr_base), int(num_classes), larr(starts), larr(lens),
                                         larr(ctr_starts) if ctr_starts is not None else None, S, float(threshold), int(k),
                                         int(mode), _p(vals), _p(idx), _p(cnt), _p(ws), ws.numel(), _stream(base)))
    return vals, idx, cnt


def score_filter_topk_nchw(head_list, threshold, k, num_classes, mode=_lib.SCORE_SIGMOID, ctr_list=None, workspace=None):
    """Score filter + per-level top-k straight from the head outputs: head_list[l] (B, A*C, H, W) as the network
    writes them (ctr_list[l] (B, A, H, W) for FCOS).  Segment s = b*L + l; indices are those of the reference's
    permuted layout (function.py:26-32), so the result equals score_filter_topk on the permuted tensors."""
    lib = _lib.load()
    hs = [_f32c(h, "logits") for h in head_list]
    C = int(num_classes)
    A = hs[0].shape[1] // C
    assert all(h.ndim == 4 and h.shape[1] == A * C for h in hs)
    base, starts, lens = _segments([h.reshape(h.shape[0], -1) for h in hs])
    hw = []
    for b in range(hs[0].shape[0]):
        hw += [h.shape[2] * h.shape[3] for h in hs]
    cbase, cstarts = None, None
    if ctr_list is not None:
        cs = [_f32c(c, "ctrness") for c in ctr_list]
        assert all(c.shape[1] == A for c in cs)
        cbase, cstarts, _ = _segments([c.reshape(c.shape[0], -1) for c in cs])
    S, dev = len(lens), base.device
    vals = torch.empty((S, k), dtype=torch.float32, device=dev)
    idx = torch.empty((S, k), dtype=torch.int32, device=dev)
    cnt = torch.empty((S,), dtype=torch.int32, device=dev)
    need = lib.bdet_score_filter_topk_workspace(sum(int(n) for n in lens), S, k)
    ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, dev)
    with _guard(base):
        check(lib.bdet_score_filter_topk_nchw(_p(base), _p(cbase), C, A, larr(starts), iarr(hw),
                                              larr(cstarts) if cstarts is not None else None, S, float(threshold), int(k),
                                              int(mode), _p(vals), _p(idx), _p(cnt), _p(ws), ws.numel(), _stream(base)))
    return vals, idx, cnt


def scores(logits, mode=_lib.SCORE_SIGMOID, ctrness=None, num_classes=1):
    lib = _lib.load()
    lg = _f32c(logits, "logits")
    ct = _f32c(ctrness, "ctrness").reshape(-1) if ctrness is not None else None
    out = torch.empty_like(lg)
    with _guard(lg):
        check(lib.bdet_scores(_p(lg), _p(ct), int(num_classes), lg.numel(), int(mode), _p(out), _stream(lg)))
    return out


def select_decode(anchors, deltas, topk, k, div, coder=0, label_mode=0, mean=(0, 0, 0, 0), std=(1, 1, 1, 1), im_info=None,
                  with_runs=False, nchw=False):
    """anchors: L tensors (n_l, 4|2); deltas: L tensors (B, n_l, 4) -- or, with nchw=True, the head outputs
    (B, A*4, H, W); topk = (vals, idx, cnt) with segment s = b*L + l.
    Returns boxes (B, L*k, 4), scores (B, L*k), labels (B, L*k) [int32 or fp32 level ids], count (B,)
    [, run_end (B, L): end of every level's (already score-sorted) run, for ``nms_batched(runs=...)``]."""
    lib = _lib.load()
    L = len(anchors)
    anc = [_f32c(a) for a in anchors]
    dl = [_f32c(d) for d in deltas]
    B = dl[0].shape[0]
    vals, idx, cnt = topk
    dev = dl[0].device
    boxes = torch.empty((B, L * k, 4), dtype=torch.float32, device=dev)
    sc = torch.empty((B, L * k), dtype=torch.float32, device=dev)
    labels = torch.empty((B, L * k), dtype=torch.float32 if label_mode == 1 else torch.int32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    run_end = torch.empty((B, L), dtype=torch.int32, device=dev) if with_runs else None
    ap = (ctypes.c_void_p * L)(*[a.data_ptr() for a in anc])
    dp_ = (ctypes.c_void_p * L)(*[d.data_ptr() for d in dl])
    info = _f32c(im_info) if im_info is not None else None
    hw = None
    if nchw:
        assert all(d.ndim == 4 and d.shape[1] * d.shape[2] * d.shape[3] == 4 * a.shape[0] for d, a in zip(dl, anc))
        hw = iarr([d.shape[2] * d.shape[3] for d in dl])
    ws = _workspace(lib.bdet_select_decode_workspace(L, B, int(k)), dev) if info is not None else None
    with _guard(boxes):
        check(lib.bdet_select_decode_ws(ap, dp_, iarr([a.shape[0] for a in anc]), hw, L, B, int(k), int(div), int(coder),
                                        int(label_mode), _p(idx), _p(vals), _p(cnt), farr(mean), farr(std), _p(info),
                                        info.shape[1] if info is not None else 0, _p(boxes), _p(sc), _p(labels),
                                        _p(count), _p(run_end), _p(ws), ws.numel() if ws is not None else 0,
                                        _stream(boxes)))
    if with_runs:
        return boxes, sc, labels, count, run_end
    return boxes, sc, labels, count


def dense_tail(anchors, deltas, topk, k, div, coder, iou_thresh, max_out, im_info, mean=(0, 0, 0, 0), std=(1, 1, 1, 1), nchw=False):
    """select_decode -> NMS -> finalize of a dense head in one kernel (one CTA per image); returns None when the candidate
    set does not fit shared memory (callers then use the separate kernels).  -> dets (B, max_out, 6), count (B,)."""
    lib = _lib.load()
    L = len(anchors)
    if lib.bdet_dense_tail_smem(L, int(k), int(max_out)) > 200 * 1024:
        return None
    anc = [_f32c(a) for a in anchors]
    dl = [_f32c(d) for d in deltas]
    B = dl[0].shape[0]
    vals, idx, cnt = topk
    dev = dl[0].device
    dets = torch.empty((B, int(max_out), 6), dtype=torch.float32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    ap = (ctypes.c_void_p * L)(*[a.data_ptr() for a in anc])
    dp_ = (ctypes.c_void_p * L)(*[d.data_ptr() for d in dl])
    info = _f32c(im_info) if im_info is not None else None
    hw = iarr([d.shape[2] * d.shape[3] for d in dl]) if nchw else None
    with _guard(dets):
        check(lib.bdet_dense_tail(ap, dp_, iarr([a.shape[0] for a in anc]), hw, L, B, int(k), int(div), int(coder), _p(idx),
                                  _p(vals), _p(cnt), farr(mean), farr(std), _p(info), info.shape[1] if info is not None else 0,
                                  float(iou_thresh), int(max_out), _p(dets), _p(count), _stream(dets)))
    return dets, count


def finalize_detections(boxes, scores_, labels, keep, keep_count, max_out, im_info=None, mode=0):
    """mode 0 -> (B, max_out, 6) [x1,y1,x2,y2,score,label] scaled/clipped by im_info; mode 1 -> (B, max_out, 5) rois."""
    lib = _lib.load()
    B, N = boxes.shape[0], boxes.shape[1]
    out = torch.empty((B, max_out, 6 if mode == 0 else 5), dtype=torch.float32, device=boxes.device)
    info = _f32c(im_info) if im_info is not None else None
    lf = labels is not None and labels.dtype.is_floating_point
    with _guard(boxes):
        check(lib.bdet_finalize_detections(_p(boxes), _p(scores_), _p(labels), int(lf), N, _p(keep), keep.shape[1],
                                           _p(keep_count), _p(info), info.shape[1] if info is not None else 0, B,
                                           int(max_out), int(mode), _p(out), _stream(boxes)))
    return out


# ----------------------------------------------------------------------------- NMS
def nms_batched(boxes, scores_, idxs, iou_thresh, max_output=None, num=None, workspace=None, runs=None):
    """boxes (B,Nmax,4), scores (B,Nmax), idxs (B,Nmax) int32/fp32 or None, num (B,) int32 or None.
    runs (B, R) int32, optional: each image's list is R back-to-back runs already in (score desc, index asc) order,
    ending at runs[b, r] (``select_decode(with_runs=True)``); only a speed hint, the result is the same.
    Returns (keep (B,cap) int32 original indices in score-descending order, keep_count (B,) int32)."""
    lib = _lib.load()
    b = _f32c(boxes, "boxes")
    s = _f32c(scores_, "scores")
    assert b.ndim == 3 and b.shape[2] == 4 and s.shape == b.shape[:2]
    B, Nmax = s.shape
    if idxs is not None:
        ix = _dev(as_tensor(idxs), "idxs")
        is_float = ix.dtype.is_floating_point
        ix = ix.float().contiguous() if is_float else ix.to(torch.int32).contiguous()
        assert ix.shape == s.shape
    else:
        ix, is_float = None, False
    cap = Nmax if not max_output or max_output <= 0 else min(int(max_output), Nmax)
    keep = torch.empty((B, max(cap, 1)), dtype=torch.int32, device=b.device)
    cnt = torch.empty((B,), dtype=torch.int32, device=b.device)
    nd = _i32c(num) if num is not None else None
    need = lib.bdet_nms_workspace(Nmax, B)
    ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, b.device)
    rn = _i32c(runs) if runs is not None else None
    assert rn is None or (rn.ndim == 2 and rn.shape[0] == B)
    with _guard(b):
        check(lib.bdet_nms_runs(_p(b), _p(s), _p(ix), int(is_float), _p(nd), _p(rn), rn.shape[1] if rn is not None else 0,
                                Nmax, B, float(iou_thresh), int(max_output) if max_output else 0, _p(keep), cap, _p(cnt),
                                _p(ws), ws.numel(), _stream(b)))
    return keep, cnt


# ----------------------------------------------------------------------------- ROI pooling
def roi_assign_levels(rois, min_level, max_level):
    lib = _lib.load()
    r = _f32c(rois, "rois")
    out = torch.empty((r.shape[0],), dtype=torch.int32, device=r.device)
    with _guard(r):
        check(lib.bdet_roi_assign_levels(_p(r), r.shape[0], int(min_level), int(max_level), _p(out), _stream(r)))
    return out


def _level_args(features):
    n = len(features)
    ptrs = (ctypes.c_void_p * n)(*[f.data_ptr() for f in features])
    hw = [int(v) for f in features for v in f.shape[-2:]]
    return n, ptrs, iarr(hw)


ROI_ORDER_MIN, ROI_ORDER_MAX = 1024, 16384


def roi_order(feature_shapes, rois, levels, scales, pool_shape, aligned=True):
    """Processing order of the ROI kernels (bdet_roi_order): rois sorted by (image, level, tile) so that concurrently
    processed rois share feature planes in L2.  -> perm (K,) int32, or None when K is outside [1024, 16384]."""
    lib = _lib.load()
    r = _f32c(rois, "rois")
    K = r.shape[0]
    if K < ROI_ORDER_MIN or K > ROI_ORDER_MAX:
        return None
    lv = _i32c(levels) if levels is not None else None
    hw = iarr([int(v) for s in feature_shapes for v in s[-2:]])
    perm = torch.empty((K,), dtype=torch.int32, device=r.device)
    with _guard(r):
        check(lib.bdet_roi_order(len(feature_shapes), hw, farr(scales), int(feature_shapes[0][0]), _p(r), _p(lv), K,
                                 pool_shape[0], pool_shape[1], int(bool(aligned)), _p(perm), _stream(r)))
    return perm


def roi_align_fwd(features, rois, levels, scales, pool_shape, sample_points=(2, 2), aligned=True, perm=None):
    """features: list of (B,C,H_l,W_l) fp32 contiguous; rois (K,5); levels (K,) int32 or None -> (K,C,PH,PW).
    perm: optional processing order from ``roi_order`` (results are in roi order either way)."""
    lib = _lib.load()
    feats = [_f32c(f, "feature") for f in features]
    B, C = feats[0].shape[:2]
    for f in feats:
        assert f.ndim == 4 and f.shape[0] == B and f.shape[1] == C
    r = _f32c(rois, "rois")
    assert r.ndim == 2 and r.shape[1] == 5
    K = r.shape[0]
    lv = _i32c(levels) if levels is not None else None
    PH, PW = pool_shape
    out = torch.empty((K, C, PH, PW), dtype=torch.float32, device=r.device)
    n, ptrs, hw = _level_args(feats)
    with _guard(r):
        check(lib.bdet_roi_align_fwd_perm(ptrs, n, hw, farr(scales), B, C, _p(r), _p(lv), K, PH, PW, int(sample_points[0]),
                                          int(sample_points[1]), int(bool(aligned)), _p(out), _p(perm), _stream(r)))
    return out


def roi_align_bwd(dout, feature_shapes, rois, levels, scales, pool_shape, sample_points=(2, 2), aligned=True,
                  dfeats=None, accumulate=False, gather=False, workspace=None, perm=None):
    """Gradient w.r.t. the features.  ``dfeats`` (list) are written (accumulate=False) or added to (True);
    allocated when None.  gather=False (default, currently the faster one) is the shared-memory scatter kernel;
    gather=True the atomics-free, run-to-run deterministic tile-gather kernel."""
    lib = _lib.load()
    d = _f32c(dout, "dout")
    r = _f32c(rois, "rois")
    K = r.shape[0]
    lv = _i32c(levels) if levels is not None else None
    if dfeats is None:
        assert not accumulate
        dfeats = [torch.empty(tuple(s), dtype=torch.float32, device=d.device) for s in feature_shapes]
    B, C = dfeats[0].shape[:2]
    PH, PW = pool_shape
    assert d.shape == (K, C, PH, PW)
    n, ptrs, hw = _level_args(dfeats)
    ws = None
    if gather:
        need = lib.bdet_roi_align_bwd_workspace(n, hw, B, max(K, 1))
        ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, d.device)
    with _guard(d):
        check(lib.bdet_roi_align_bwd_perm(ptrs, n, hw, farr(scales), B, C, _p(r), _p(lv), K, PH, PW, int(sample_points[0]),
                                          int(sample_points[1]), int(bool(aligned)), _p(d), int(bool(accumulate)), _p(perm),
                                          _p(ws), ws.numel() if ws is not None else 0, _stream(d)))
    return dfeats


def roi_maxpool_fwd(features, rois, levels, scales, pool_shape):
    """F.nn.roi_pooling(mode="max") over all levels: -> (out (K,C,PH,PW), argmax (K,C,PH,PW) int32)."""
    lib = _lib.load()
    feats = [_f32c(f, "feature") for f in features]
    B, C = feats[0].shape[:2]
    r = _f32c(rois, "rois")
    K = r.shape[0]
    lv = _i32c(levels) if levels is not None else None
    PH, PW = pool_shape
    out = torch.empty((K, C, PH, PW), dtype=torch.float32, device=r.device)
    am = torch.empty((K, C, PH, PW), dtype=torch.int32, device=r.device)
    n, ptrs, hw = _level_args(feats)
    with _guard(r):
        check(lib.bdet_roi_maxpool_fwd(ptrs, n, hw, farr(scales), B, C, _p(r), _p(lv), K, PH, PW, _p(out), _p(am), _stream(r)))
    return out, am


def roi_maxpool_bwd(dout, argmax, feature_shapes, rois, levels, pool_shape):
    lib = _lib.load()
    d = _f32c(dout, "dout")
    r = _f32c(rois, "rois")
    lv = _i32c(levels) if levels is not None else None
    dfeats = [torch.empty(tuple(s), dtype=torch.float32, device=d.device) for s in feature_shapes]
    B, C = dfeats[0].shape[:2]
    n, ptrs, hw = _level_args(dfeats)
    with _guard(d):
        check(lib.bdet_roi_maxpool_bwd(ptrs, n, hw, B, C, _p(r), _p(lv), r.shape[0], pool_shape[0], pool_shape[1], _p(d),
                                       _p(_i32c(argmax)), 0, _stream(d)))
    return dfeats


# ----------------------------------------------------------------------------- small Boxes / glue ops
def box_props(boxes, mode):
    """mode 0 width, 1 height, 2 area (structures/boxes.py:36-52)."""
    lib = _lib.load()
    b, ld = _rows(boxes)
    out = torch.empty((b.shape[0],), dtype=torch.float32, device=b.device)
    with _guard(b):
        check(lib.bdet_box_props(_p(b), ld, b.shape[0], int(mode), _p(out), _stream(b)))
    return out


def box_convert(boxes, from_mode, to_mode):
    lib = _lib.load()
    b = _f32c(boxes)
    out = torch.empty_like(b)
    with _guard(b):
        check(lib.bdet_box_convert(_p(b), b.shape[0], int(from_mode), int(to_mode), _p(out), _stream(b)))
    return out


def cond_take(x, mask=None):
    """F.cond_take(mask, x) -> (values, ascending int32 flat indices).  mask None means x != 0.
    The variable-length result needs the count on the host: one D2H read, as in the reference API."""
    lib = _lib.load()
    xf = _f32c(x, "x").reshape(-1)
    n = xf.numel()
    m = None
    if mask is not None:
        m = _dev(as_tensor(mask), "mask").reshape(-1).to(torch.uint8).contiguous()
        assert m.numel() == n
    vals = torch.empty((n,), dtype=torch.float32, device=xf.device)
    idx = torch.empty((n,), dtype=torch.int32, device=xf.device)
    cnt = torch.empty((1,), dtype=torch.int32, device=xf.device)
    ws = _workspace(lib.bdet_cond_take_workspace(n), xf.device)
    with _guard(xf):
        check(lib.bdet_cond_take(_p(xf), _p(m), n, _p(vals), _p(idx), _p(cnt), _p(ws), ws.numel(), _stream(xf)))
    k = int(cnt.item())
    return vals[:k], idx[:k]


def count_labels(labels, out=None):
    """labels (B, A) int32 -> (B, 3) int32 counts of (label < 0, label == 0, label > 0)."""
    lib = _lib.load()
    lab = _i32c(labels, "labels")
    if lab.ndim == 1:
        lab = lab.unsqueeze(0)
    B, A = lab.shape
    if out is None:
        out = torch.empty((B, 3), dtype=torch.int32, device=lab.device)
    assert out.shape == (B, 3) and out.dtype == torch.int32 and out.is_contiguous()
    with _guard(lab):
        check(lib.bdet_count_labels(_p(lab), A, B, _p(out), _stream(lab)))
    return out


def sample_labels(labels, noise, num_samples, label_value, ignore_label=-1):
    """sample_labels (layers/common/sampling.py:7-30) for a batch, IN PLACE on ``labels`` (B, A) int32.
    noise (B, A) fp32: one uniform variate per element (the explicit RNG contract, see include/bdet.h);
    num_samples: int, or a (B,) int32 device tensor (per-image budget).  Returns ``labels``."""
    lib = _lib.load()
    lab = _dev(as_tensor(labels), "labels")
    assert lab.dtype == torch.int32 and lab.is_contiguous(), "labels must be a contiguous int32 tensor (modified in place)"
    if lab.ndim == 1:
        lab = lab.unsqueeze(0)
    nz = _f32c(noise, "noise").reshape(lab.shape)
    B, A = lab.shape
    ns_dev = _i32c(num_samples, "num_samples") if torch.is_tensor(num_samples) else None
    assert ns_dev is None or ns_dev.numel() == B
    with _guard(lab):
        check(lib.bdet_sample_labels(_p(lab), _p(nz), A, B, int(label_value), int(ignore_label),
                                     0 if ns_dev is not None else int(num_samples), _p(ns_dev), _stream(lab)))
    return labels


def rcnn_match(rois, n_rois, gt_boxes, num_gt, fg_thresh=0.5, bg_thresh_low=0.0, bg_thresh_high=0.5):
    """First step of RCNN.get_ground_truth (layers/head/rcnn.py:105-123) for the batch.  rois (B, Rmax, 5) padded,
    n_rois (B,).  -> dict(all_rois (B,N,5), n_all (B,), assign (B,N), cls (B,N) fp32, fg (B,N) int32, bg (B,N) int32)."""
    lib = _lib.load()
    r = _f32c(rois, "rois")
    gt = _f32c(gt_boxes, "gt_boxes")
    assert r.ndim == 3 and r.shape[2] == 5 and gt.ndim == 3 and gt.shape[2] == 5 and r.shape[0] == gt.shape[0]
    B, Rmax, Gmax = r.shape[0], r.shape[1], gt.shape[1]
    N, dev = Rmax + Gmax, r.device
    out = dict(all_rois=torch.empty((B, N, 5), dtype=torch.float32, device=dev),
               n_all=torch.empty((B,), dtype=torch.int32, device=dev),
               assign=torch.empty((B, N), dtype=torch.int32, device=dev),
               cls=torch.empty((B, N), dtype=torch.float32, device=dev),
               fg=torch.empty((B, N), dtype=torch.int32, device=dev), bg=torch.empty((B, N), dtype=torch.int32, device=dev))
    with _guard(r):
        check(lib.bdet_rcnn_match(_p(r), _p(_i32c(n_rois, "n_rois")), Rmax, _p(gt), _p(_i32c(num_gt, "num_gt")), Gmax, B,
                                  float(fg_thresh), float(bg_thresh_low), float(bg_thresh_high), _p(out["all_rois"]),
                                  _p(out["n_all"]), _p(out["assign"]), _p(out["cls"]), _p(out["fg"]), _p(out["bg"]), _stream(r)))
    return out


def rcnn_collect(m, gt_boxes, num_out, mean=(0, 0, 0, 0), std=(0.1, 0.1, 0.2, 0.2)):
    """Last step of RCNN.get_ground_truth (rcnn.py:130-137) from the (sampled) masks of ``rcnn_match``.
    -> rois (B, num_out, 5), labels (B, num_out) int32, bbox_targets (B, num_out, 4), count (B,)."""
    lib = _lib.load()
    gt = _f32c(gt_boxes, "gt_boxes")
    B, N = m["fg"].shape
    dev = gt.device
    rois = torch.empty((B, num_out, 5), dtype=torch.float32, device=dev)
    labels = torch.empty((B, num_out), dtype=torch.int32, device=dev)
    targets = torch.empty((B, num_out, 4), dtype=torch.float32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    with _guard(gt):
        check(lib.bdet_rcnn_collect(_p(m["all_rois"]), _p(m["n_all"]), _p(m["assign"]), _p(m["cls"]), _p(m["fg"]), _p(m["bg"]), N,
                                    _p(gt), gt.shape[1], B, farr(mean), farr(std), int(num_out), _p(rois), _p(labels),
                                    _p(targets), _p(count), _stream(gt)))
    return rois, labels, targets, count


def ota_topk_match(cost, ious, candidate_k=10):
    """OTATopkMatcher (layers/common/matcher.py:134-161): cost, ious (G, A) fp32 -> matched GT per anchor (A,) int32,
    G = background."""
    lib = _lib.load()
    c = _f32c(cost, "cost")
    u = _f32c(ious, "ious")
    assert c.ndim == 2 and c.shape == u.shape
    G, A = c.shape
    out = torch.empty((A,), dtype=torch.int32, device=c.device)
    ws = _workspace(lib.bdet_ota_topk_match_workspace(A), c.device)
    with _guard(c):
        check(lib.bdet_ota_topk_match(_p(c), A, _p(u), A, G, A, int(candidate_k), _p(out), _p(ws), ws.numel(), _stream(c)))
    return out


def ota_cost(points, radius, gt5, cls_logits, pred_deltas, alpha=0.25, gamma=2.0, reg_weight=1.5):
    """OTA cost construction (models/det/ota.py:91-152) for one image: points (A,2), radius (A) = stride * 2.5 of each
    point's level, gt5 (G,5), cls_logits (A,C), pred_deltas (A,4) ltrb -> cost (G,A), ious (G,A)."""
    lib = _lib.load()
    pts, rad, gt = _f32c(points, "points"), _f32c(radius, "radius"), _f32c(gt5, "gt")
    lg, dl = _f32c(cls_logits, "cls_logits"), _f32c(pred_deltas, "pred_deltas")
    A, G, C = pts.shape[0], gt.shape[0], lg.shape[1]
    assert pts.shape == (A, 2) and rad.shape == (A,) and lg.shape[0] == A and dl.shape == (A, 4) and gt.shape[1] == 5
    cost = torch.empty((G, A), dtype=torch.float32, device=pts.device)
    ious = torch.empty((G, A), dtype=torch.float32, device=pts.device)
    ws = _workspace(lib.bdet_ota_cost_workspace(A), pts.device)
    with _guard(pts):
        check(lib.bdet_ota_cost(_p(pts), _p(rad), A, _p(gt), G, _p(lg), C, _p(dl), float(alpha), float(gamma), float(reg_weight),
                                _p(cost), _p(ious), _p(ws), ws.numel(), _stream(pts)))
    return cost, ious


def ota_collect(matched, points, gt5, ious):
    """Targets of the matched points (ota.py:160-175): -> gt_classes (A,) fp32, box_targets (A,4), iou_targets (A,)."""
    lib = _lib.load()
    m, pts, gt, io = _i32c(matched, "matched"), _f32c(points, "points"), _f32c(gt5, "gt"), _f32c(ious, "ious")
    A, G = pts.shape[0], gt.shape[0]
    cls_t = torch.empty((A,), dtype=torch.float32, device=pts.device)
    box_t = torch.empty((A, 4), dtype=torch.float32, device=pts.device)
    iou_t = torch.empty((A,), dtype=torch.float32, device=pts.device)
    with _guard(pts):
        check(lib.bdet_ota_collect(_p(m), _p(pts), A, _p(gt), G, _p(io), _p(cls_t), _p(box_t), _p(iou_t), _stream(pts)))
    return cls_t, box_t, iou_t


def free_anchor_box_prob(pred_boxes, gt5, num_classes, box_iou_thresh=0.6, clamp_eps=1e-7):
    """FreeAnchor box-probability scatter (free_anchor.py:54-86): pred_boxes (A,4) decoded, gt5 (G,5) -> (A,C)."""
    import numpy as np

    lib = _lib.load()
    pb, gt = _f32c(pred_boxes, "pred_boxes"), _f32c(gt5, "gt")
    A, G = pb.shape[0], gt.shape[0]
    out = torch.empty((A, int(num_classes)), dtype=torch.float32, device=pb.device)
    ws = _workspace(lib.bdet_free_anchor_box_prob_workspace(G), pb.device)
    lower = float(np.float32(float(box_iou_thresh) + float(clamp_eps)))  # thresh1 + clamp_eps in Python floats, then fp32
    with _guard(pb):
        check(lib.bdet_free_anchor_box_prob(_p(pb), A, _p(gt), G, int(num_classes), float(box_iou_thresh), lower,
                                            float(clamp_eps), _p(out), _p(ws), ws.numel(), _stream(pb)))
    return out


def free_anchor_bags(matched_idx, anchors, gt5, pred_scores, mean=(0, 0, 0, 0), std=(0.1, 0.1, 0.2, 0.2)):
    """Bag scores / BoxCoder targets (free_anchor.py:95-113): matched_idx (G,K) int32 -> (G,K), (G*K,4)."""
    lib = _lib.load()
    mi, an, gt, sc = _i32c(matched_idx, "matched_idx"), _f32c(anchors, "anchors"), _f32c(gt5, "gt"), _f32c(pred_scores, "scores")
    G, K = mi.shape
    ms = torch.empty((G, K), dtype=torch.float32, device=an.device)
    mo = torch.empty((G * K, 4), dtype=torch.float32, device=an.device)
    with _guard(an):
        check(lib.bdet_free_anchor_bags(_p(mi), G, K, _p(an), _p(gt), _p(sc), sc.shape[1], farr(mean), farr(std), _p(ms), _p(mo),
                                        _stream(an)))
    return ms, mo


def coco_format(dets, counts, image_ids, category_ids=None):
    """COCOEvaluator.format on the device (coco_eval.py:111-138): dets (B,K,6) + counts (B,) + image_ids (B,) int32
    [+ category_ids (C,) int32 = classes_originID] -> (image_id (N,), bbox_xywh (N,4) f64, score (N,) f64, category_id (N,)),
    N = counts.sum() (one D2H read of the total: the result is variable-length)."""
    lib = _lib.load()
    d, c, ids = _f32c(dets, "dets"), _i32c(counts, "counts"), _i32c(image_ids, "image_ids")
    B, K = d.shape[0], d.shape[1]
    cat = _i32c(category_ids, "category_ids") if category_ids is not None else None
    dev = d.device
    ri = torch.empty((B * K,), dtype=torch.int32, device=dev)
    rb = torch.empty((B * K, 4), dtype=torch.float64, device=dev)
    rs = torch.empty((B * K,), dtype=torch.float64, device=dev)
    rc = torch.empty((B * K,), dtype=torch.int32, device=dev)
    tot = torch.zeros((1,), dtype=torch.int32, device=dev)
    with _guard(d):
        check(lib.bdet_coco_format(_p(d), _p(c), B, K, _p(ids), _p(cat), cat.numel() if cat is not None else 0, _p(ri), _p(rb),
                                   _p(rs), _p(rc), _p(tot), _stream(d)))
    n = int(tot.item())
    return ri[:n], rb[:n], rs[:n], rc[:n]


# ----------------------------------------------------------------------------- measurement hooks
def profile_begin(only=None):
    """Bracket kernel launches with CUDA events; ``only`` = time just that kernel (the rest are counted)."""
    lib = _lib.load()
    check(lib.bdet_profile_select(only.encode() if only else None))
    check(lib.bdet_profile_begin())


def profile_collect(name=None):
    """(total_ms, launches) of the named kernel since profile_begin() (None = all kernels)."""
    ms, n = ctypes.c_float(0), ctypes.c_int(0)
    check(_lib.load().bdet_profile_collect(name.encode() if name else None, ctypes.byref(ms), ctypes.byref(n)))
    return ms.value, n.value


def profile_report():
    """{kernel name: (total_ms, launches)} of every bracketed kernel since profile_begin()."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(_lib.load().bdet_profile_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, n = line.rsplit(" ", 2)
        out[name] = (float(ms), int(n))
    return out


def profile_end():
    check(_lib.load().bdet_profile_end())
