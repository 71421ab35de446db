"""GPU parity of the TMA ROIAlign kernels (csrc/roi_tma.cu): forward bit-exact, backward <= 1e-5 against the fp64-
accumulated oracle, on pyramids that mix TMA-capable levels (W % 4 == 0) with one that is not, channel counts that do not
fill the last pipeline stage, and ROIs that leave the map, are empty, or are too wide for a TMA box (direct fallback)."""
import numpy as np
import pytest
import torch

from basedet_b200 import ops
from basedet_b200 import workloads as W
from oracle import c_oracle as C
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
STRIDES = [4, 8, 16, 32]


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def special_rois(rng, B, hw, n):
    rois = W.make_rois(rng, n, B, hw[0], hw[1], 4, 500)
    h, w = hw
    rois[0, 1:] = [-30, -20, 40, 50]            # sticks out top-left: zero-padded taps
    rois[1, 1:] = [w - 40, h - 30, w + 60, h + 45]   # beyond the far corner
    rois[2, 1:] = [50, 50, 50, 50]              # empty
    rois[3, 1:] = [-400, -400, -300, -300]      # entirely outside
    rois[4, 1:] = [0, 0, w, h]                  # the whole image (wide footprint on its level)
    rois[5, 1:] = [3, 3, 3.5, 3.25]             # sub-pixel
    rois[6, 1:] = [0, 10, w, 14]                # very wide, 1 px tall
    rois[7, 1:] = [10, 0, 14, h]                # very tall
    return rois


@pytest.mark.parametrize("C_,hw,force_level", [(8, (160, 224), None), (72, (256, 320), None), (16, (256, 320), 0),
                                               (256, (128, 192), None)])
def test_tma_forward_backward_vs_oracle(C_, hw, force_level):
    rng = np.random.default_rng(20 + C_)
    B = 2
    sizes = [(-(-hw[0] // s), -(-hw[1] // s)) for s in STRIDES]
    feats = [rng.normal(0, 1, (B, C_, h, w)).astype(np.float32) for h, w in sizes]
    rois = special_rois(rng, B, hw, 60)
    levels = R.assign_levels(rois, STRIDES)
    if force_level is not None:                 # everything on the finest level: footprints up to the whole map
        levels[:] = force_level
    K = rois.shape[0]
    dout = rng.normal(0, 1, (K, C_, 7, 7)).astype(np.float32)
    scales = [1.0 / s for s in STRIDES]
    tf = [T(f) for f in feats]
    out = ops.roi_align_fwd(tf, T(rois), T(levels), scales, (7, 7)).cpu().numpy()
    grads = ops.roi_align_bwd(T(dout), [f.shape for f in feats], T(rois), T(levels), scales, (7, 7))
    base = [torch.full_like(t, 0.5) for t in tf]
    acc = ops.roi_align_bwd(T(dout), None, T(rois), T(levels), scales, (7, 7), dfeats=base, accumulate=True)
    for l, f in enumerate(feats):
        sel = np.flatnonzero(levels == l)
        if len(sel) == 0:
            assert not grads[l].any()
            continue
        ref = C.roi_align_fwd(f, rois[sel], (7, 7), scales[l])
        assert np.array_equal(out[sel], ref), (l, np.abs(out[sel] - ref).max())
        gref = C.roi_align_bwd(dout[sel], f.shape, rois[sel], (7, 7), scales[l])
        g = grads[l].cpu().numpy()
        assert np.max(np.abs(g - gref)) / max(np.abs(gref).max(), 1.0) <= 1e-5, l
        ga = acc[l].cpu().numpy()
        assert np.max(np.abs(ga - 0.5 - gref)) / max(np.abs(gref).max(), 1.0) <= 1e-5, l


def test_tma_forward_many_rois_one_level_bit_exact():
    """2 000 ROIs of every size on one 200 x 336 map (config-3 P2 shape), 32 channels: bit-exact."""
    rng = np.random.default_rng(31)
    feat = rng.normal(0, 1, (2, 32, 200, 336)).astype(np.float32)
    rois = W.make_rois(rng, 1000, 2, 800, 1344, 2, 300)
    out = ops.roi_align_fwd([T(feat)], T(rois), None, [0.25], (7, 7)).cpu().numpy()
    ref = C.roi_align_fwd(feat, rois, (7, 7), 0.25)
    assert np.array_equal(out, ref)
    dout = rng.normal(0, 1, (2000, 32, 7, 7)).astype(np.float32)
    g = ops.roi_align_bwd(T(dout), [feat.shape], T(rois), None, [0.25], (7, 7))[0].cpu().numpy()
    gref = C.roi_align_bwd(dout, feat.shape, rois, (7, 7), 0.25)
    assert np.max(np.abs(g - gref)) / max(np.abs(gref).max(), 1.0) <= 1e-5


def test_roi_processing_order_is_a_permutation_and_changes_no_result():
    """bdet_roi_order: a permutation of 0..K-1 sorted by (image, level, tile); the forward with it is bit-identical to the
    forward without it, the backward (fp32 reductions in another order) stays within 1e-5 of the oracle."""
    rng = np.random.default_rng(41)
    B, C_, hw = 3, 24, (256, 320)
    sizes = [(-(-hw[0] // s), -(-hw[1] // s)) for s in STRIDES]
    feats = [rng.normal(0, 1, (B, C_, h, w)).astype(np.float32) for h, w in sizes]
    rois = W.make_rois(rng, 500, B, hw[0], hw[1], 4, 300)
    rois = rois[rng.permutation(len(rois))]          # images interleaved, as no caller would hand them over
    rois[:8] = special_rois(rng, B, hw, 20)[:8]
    levels = R.assign_levels(rois, STRIDES)
    K = rois.shape[0]
    scales = [1.0 / s for s in STRIDES]
    tf, tr, tl = [T(f) for f in feats], T(rois), T(levels)
    shapes = [f.shape for f in feats]
    perm = ops.roi_order(shapes, tr, tl, scales, (7, 7))
    assert perm is not None
    pn = perm.cpu().numpy()
    assert np.array_equal(np.sort(pn), np.arange(K))
    key = rois[pn, 0].astype(np.int64) * len(STRIDES) + levels[pn]
    assert np.all(np.diff(key) >= 0)
    plain = ops.roi_align_fwd(tf, tr, tl, scales, (7, 7))
    ordered = ops.roi_align_fwd(tf, tr, tl, scales, (7, 7), perm=perm)
    assert torch.equal(plain, ordered)
    dout = rng.normal(0, 1, (K, C_, 7, 7)).astype(np.float32)
    grads = ops.roi_align_bwd(T(dout), shapes, tr, tl, scales, (7, 7), perm=perm)
    for l, f in enumerate(feats):
        sel = np.flatnonzero(levels == l)
        gref = C.roi_align_bwd(dout[sel], f.shape, rois[sel], (7, 7), scales[l])
        assert np.max(np.abs(grads[l].cpu().numpy() - gref)) / max(np.abs(gref).max(), 1.0) <= 1e-5, l
    assert ops.roi_order(shapes, tr[:100], tl[:100], scales, (7, 7)) is None     # below the size where it pays


def test_tma_backward_opt_in_path():
    """The cp.reduce.async.bulk.tensor backward is opt-in (BDET_ROI_BWD_TMA_CLS, read once per process; the direct
    scatter kernel measured faster): run this file's oracle comparison in a child process with it enabled for every
    footprint width."""
    import os
    import subprocess
    import sys

    if os.environ.get("BDET_ROI_BWD_TMA_CLS"):
        pytest.skip("already inside the opt-in child")
    env = dict(os.environ, BDET_ROI_BWD_TMA_CLS="6")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-m", "gpu", "-k", "vs_oracle or many_rois", "-x"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
