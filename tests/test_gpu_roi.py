"""GPU parity: FPN level assignment, multi-level ROIAlign forward / backward vs the oracle (<= 1e-5 relative)."""
import numpy as np
import pytest
import torch

from basedet_b200 import ops
from basedet_b200 import workloads as W
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def T(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


def rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0))) if ref.size else 0.0


def test_roi_align_reference_kat(cuda):
    from tests.test_oracle_kat import ROI, ROI_ALIGN_EXPECTED, ROI_FEAT

    out = ops.roi_align_fwd([T(ROI_FEAT, cuda)], T(ROI, cuda), None, [1.0], (4, 4))
    assert np.allclose(out.cpu().numpy()[0, 0], ROI_ALIGN_EXPECTED)
    # border rule (oracle ASSUMED-6): zero padding, not clamping
    out = ops.roi_align_fwd([T(ROI_FEAT, cuda)], T(np.array([[0, 0, 0, 2, 2]], np.float32), cuda), None, [1.0], (2, 2))
    assert np.allclose(out.cpu().numpy()[0, 0], [[0.65625, 1.5], [4.5, 6.0]])


def test_assign_levels(cuda):
    rng = np.random.default_rng(1)
    rois = W.make_rois(rng, 2000, 1, 800, 1344, 4, 900)
    rois[0, 1:] = [10, 10, 10, 10]      # zero area -> -inf -> lowest level
    rois[1, 1:] = [10, 10, 5, 20]       # negative area -> NaN -> lowest level
    got = ops.roi_assign_levels(T(rois, cuda), 2, 5).cpu().numpy()
    ref = R.assign_levels(rois, W.FRCNN_RCNN_STRIDES)
    # logf ulp differences can only matter when sqrt(area)/224 is within 1e-6 of a power of two
    s = np.sqrt(np.maximum((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]), 1e-9)) / 224
    safe = np.abs(np.log2(s) - np.round(np.log2(s))) > 1e-5
    assert np.array_equal(got[safe], ref[safe])
    assert set(np.unique(got)) <= {0, 1, 2, 3}


def _pyramid(rng, B, C, hw):
    sizes = [(-(-hw[0] // s), -(-hw[1] // s)) for s in W.FRCNN_RCNN_STRIDES]
    return [rng.normal(0, 1, (B, C, h, w)).astype(np.float32) for h, w in sizes]


def test_roi_pool_multilevel_forward(cuda):
    rng = np.random.default_rng(2)
    B, C = 2, 8
    feats = _pyramid(rng, B, C, (160, 224))
    rois = W.make_rois(rng, 40, B, 160, 224, 6, 250)
    rois[3, 1:] = [-30, -20, 40, 50]        # sticks out of the image: zero-padded taps
    rois[4, 1:] = [200, 140, 260, 200]      # partly beyond the far border
    rois[5, 1:] = [50, 50, 50, 50]          # empty roi
    levels = R.assign_levels(rois, W.FRCNN_RCNN_STRIDES)
    out = ops.roi_align_fwd([T(f, cuda) for f in feats], T(rois, cuda), T(levels, cuda),
                            [1.0 / s for s in W.FRCNN_RCNN_STRIDES], (7, 7)).cpu().numpy()
    ref = R.roi_pool(feats, rois, W.FRCNN_RCNN_STRIDES, (7, 7))
    assert out.shape == ref.shape == (80, C, 7, 7)
    assert rel(out, ref) <= 1e-5
    assert np.array_equal(out, ref)  # same op order and no FMA contraction: expected bit-exact


@pytest.mark.parametrize("pool,samples", [((7, 7), (2, 2)), ((4, 4), (2, 2)), ((3, 5), (1, 3)), ((14, 14), (2, 2))])
def test_roi_align_generic_shapes(cuda, pool, samples):
    rng = np.random.default_rng(3)
    feat = rng.normal(0, 1, (2, 5, 33, 47)).astype(np.float32)
    rois = W.make_rois(rng, 12, 2, 33 * 8, 47 * 8, 8, 200)
    out = ops.roi_align_fwd([T(feat, cuda)], T(rois, cuda), None, [1 / 8], pool, samples).cpu().numpy()
    ref = R.roi_align(feat, rois, pool, 1 / 8, samples, True)
    assert rel(out, ref) <= 1e-5


@pytest.mark.parametrize("gather", [True, False])
def test_roi_align_backward(cuda, gather):
    """gather=True: atomics-free tile-gather kernel; False: shared-memory scatter + red.global.add."""
    rng = np.random.default_rng(4)
    B, C = 2, 38   # 38 channels: one full 32-channel chunk + a ragged one
    feats = _pyramid(rng, B, C, (160, 224))
    rois = W.make_rois(rng, 30, B, 160, 224, 6, 250)
    rois[0, 1:] = [-30, -20, 40, 50]
    rois[1, 1:] = [0, 0, 224, 160]           # whole image on the coarsest level
    rois[2, 1:] = [300, 300, 340, 330]       # entirely outside the image: contributes nothing
    levels = R.assign_levels(rois, W.FRCNN_RCNN_STRIDES)
    dout = rng.normal(0, 1, (rois.shape[0], C, 7, 7)).astype(np.float32)
    scales = [1.0 / s for s in W.FRCNN_RCNN_STRIDES]
    got = ops.roi_align_bwd(T(dout, cuda), [f.shape for f in feats], T(rois, cuda), T(levels, cuda), scales, (7, 7),
                            gather=gather)
    for l, f in enumerate(feats):
        sel = levels == l
        ref = R.roi_align_backward(dout[sel], f.shape, rois[sel], (7, 7), scales[l])
        g = got[l].cpu().numpy()
        denom = max(np.abs(ref).max(), 1.0)
        assert np.max(np.abs(g - ref)) / denom <= 1e-5, l
    # accumulate=True adds to what is already there
    again = ops.roi_align_bwd(T(dout, cuda), None, T(rois, cuda), T(levels, cuda), scales, (7, 7), dfeats=[g.clone() for g in got],
                              accumulate=True, gather=gather)
    for a, g in zip(again, got):
        assert torch.allclose(a, 2 * g, rtol=1e-5, atol=1e-5)
    # adjointness: <roi_align(x), dout> == <x, roi_align_bwd(dout)>
    out = ops.roi_align_fwd([T(f, cuda) for f in feats], T(rois, cuda), T(levels, cuda), scales, (7, 7))
    lhs = float((out.double() * T(dout, cuda).double()).sum())
    rhs = sum(float((T(f, cuda).double() * g.double()).sum()) for f, g in zip(feats, got))
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)


def test_roi_align_backward_gather_is_deterministic_and_generic(cuda):
    rng = np.random.default_rng(7)
    feat_shape = (2, 5, 45, 70)
    rois = W.make_rois(rng, 40, 2, 45 * 4, 70 * 4, 6, 200)
    dout = rng.normal(0, 1, (80, 5, 3, 5)).astype(np.float32)
    a = ops.roi_align_bwd(T(dout, cuda), [feat_shape], T(rois, cuda), None, [0.25], (3, 5), (1, 3), gather=True)[0]
    b = ops.roi_align_bwd(T(dout, cuda), [feat_shape], T(rois, cuda), None, [0.25], (3, 5), (1, 3), gather=True)[0]
    assert torch.equal(a, b)
    ref = R.roi_align_backward(dout, feat_shape, rois, (3, 5), 0.25, (1, 3))
    assert np.max(np.abs(a.cpu().numpy() - ref)) / max(np.abs(ref).max(), 1.0) <= 1e-5
    empty = ops.roi_align_bwd(T(dout[:0], cuda), [feat_shape], T(rois[:0], cuda), None, [0.25], (3, 5), (1, 3), gather=True)[0]
    assert float(empty.abs().max()) == 0.0   # K = 0: the gradient is all zeros, still fully written


def test_roi_align_backward_large_footprint_fallback(cuda):
    """A ROI whose footprint exceeds the shared accumulation buffer takes the direct-atomics path."""
    rng = np.random.default_rng(5)
    feat_shape = (1, 2, 160, 200)
    rois = np.array([[0, 0, 0, 199, 159], [0, 20, 30, 60, 90]], np.float32)
    dout = rng.normal(0, 1, (2, 2, 7, 7)).astype(np.float32)
    ref = R.roi_align_backward(dout, feat_shape, rois, (7, 7), 1.0)
    for gather in (False, True):
        got = ops.roi_align_bwd(T(dout, cuda), [feat_shape], T(rois, cuda), None, [1.0], (7, 7), gather=gather)[0].cpu().numpy()
        assert np.max(np.abs(got - ref)) / max(np.abs(ref).max(), 1.0) <= 1e-5


@pytest.mark.parametrize("pool,samples", [((3, 5), (1, 3)), ((4, 4), (2, 2)), ((14, 14), (2, 2)), ((7, 7), (3, 3))])
def test_roi_align_backward_scatter_generic_shapes(cuda, pool, samples):
    """The scatter kernel's generic path (anything but 7 x 7 bins takes the table loops, padded table rows) and 7 x 7 bins
    with another sample count (separable register path) against the oracle."""
    rng = np.random.default_rng(pool[0] * 10 + samples[1])
    C, feat_shape = 5, (2, 5, 40, 56)
    rois = W.make_rois(rng, 17, 2, 160, 224, 6, 200)
    rois[0, 1:] = [-20, -10, 60, 50]
    dout = rng.normal(0, 1, (rois.shape[0], C, pool[0], pool[1])).astype(np.float32)
    got = ops.roi_align_bwd(T(dout, cuda), [feat_shape], T(rois, cuda), None, [0.25], pool, samples, gather=False)[0].cpu().numpy()
    ref = R.roi_align_backward(dout, feat_shape, rois, pool, 0.25, samples)
    assert np.max(np.abs(got - ref)) / max(np.abs(ref).max(), 1.0) <= 1e-5


def test_roi_max_pool_kat_and_oracle(cuda):
    """pooler_type="roi_pool" (roi_pool.py:62-63): the reference's known answer, then random multi-level pyramids
    against the oracle (bit-exact: a maximum), and the argmax backward through autograd."""
    from basedet_b200.layers import roi_pool
    from tests.test_oracle_kat import ROI, ROI_FEAT

    expected = np.array([[6.0, 7.0, 8.0, 8.0], [11.0, 12.0, 13.0, 13.0], [16.0, 17.0, 18.0, 18.0], [16.0, 17.0, 18.0, 18.0]])
    out = roi_pool([T(ROI_FEAT, cuda)], T(ROI, cuda), strides=[1], pool_shape=4, pooler_type="roi_pool")
    assert np.array_equal(out.cpu().numpy()[0, 0], expected)
    rng = np.random.default_rng(12)
    B, C = 2, 6
    feats = _pyramid(rng, B, C, (160, 224))
    rois = W.make_rois(rng, 50, B, 160, 224, 4, 250)
    rois[3, 1:] = [-30, -20, 40, 50]
    rois[4, 1:] = [200, 140, 260, 200]
    rois[5, 1:] = [50, 50, 50, 50]
    rois[6, 1:] = [500, 500, 600, 600]      # outside: empty bins -> 0
    tf = [T(f, cuda).requires_grad_(True) for f in feats]
    out = roi_pool(tf, T(rois, cuda), W.FRCNN_RCNN_STRIDES, (7, 7), "roi_pool")
    ref = R.roi_pool(feats, rois, W.FRCNN_RCNN_STRIDES, (7, 7), "roi_pool")
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    # backward: d(sum(out * w)) / d feat = w scattered to the argmax pixels; check against a finite structure: the gradient
    # of sum(out) counts how many bins selected each pixel, and its total equals the number of non-empty bins
    out.sum().backward()
    total = sum(float(f.grad.sum()) for f in tf)
    levels = R.assign_levels(rois, W.FRCNN_RCNN_STRIDES)
    nonempty = 0
    for l, f in enumerate(feats):
        sel = rois[levels == l]
        if len(sel):
            big = R.roi_max_pooling(np.where(np.isfinite(f), 1.0, 1.0).astype(np.float32) * 0 + 1, sel, (7, 7), 1.0 / W.FRCNN_RCNN_STRIDES[l])
            nonempty += int((big > 0).sum())
    assert abs(total - nonempty) < 0.5
