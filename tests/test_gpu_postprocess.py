"""GPU parity: top-k, score filter + top-k, class-aware batched NMS vs the oracle (bit-exact index gates)."""
import numpy as np
import pytest
import torch

from basedet_b200 import _lib, ops
from basedet_b200 import workloads as W
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def T(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


# ------------------------------------------------------------------ raw top-k (rpn.py:155)
@pytest.mark.parametrize("lens,k", [([1], 1), ([5, 0, 3], 4), ([819, 3150, 12600], 1000), ([201600], 2000),
                                     ([50400, 50400], 1000), ([4097], 4097), ([300], 1000)])
def test_topk_segments_bit_exact(cuda, lens, k):
    rng = np.random.default_rng(sum(lens) + k)
    total = sum(lens)
    scores = rng.normal(-3, 2, total).astype(np.float32)
    vals, idx, cnt = ops.topk_segments(T(scores, cuda), lens, k)
    vals, idx, cnt = vals.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy()
    off = 0
    for s, n in enumerate(lens):
        rv, ri = R.topk_desc(scores[off:off + n], k)
        assert cnt[s] == len(ri)
        assert np.array_equal(idx[s, :cnt[s]], ri)
        assert np.array_equal(vals[s, :cnt[s]], rv)
        off += n


def test_topk_ties_and_special_values(cuda):
    """Tie rule (oracle ASSUMED-3): equal scores keep ascending index; +-0 are one value."""
    rng = np.random.default_rng(7)
    scores = np.round(rng.normal(0, 1, 20000), 1).astype(np.float32)   # heavy quantisation -> many ties
    scores[::7] = 0.0
    scores[3::7] = -0.0
    scores[100] = np.inf
    scores[200] = -np.inf
    for k in (1, 10, 1000, 4096, 16384):
        vals, idx, cnt = ops.topk_segments(T(scores, cuda), [20000], k)
        rv, ri = R.topk_desc(scores, k)
        assert int(cnt[0]) == k
        assert np.array_equal(idx[0].cpu().numpy()[:k], ri)
    with pytest.raises(_lib.BdetError):  # k beyond the shared-memory sort network is refused, not approximated
        ops.topk_segments(T(scores, cuda), [20000], 20000)
    const = np.full(5000, 0.25, dtype=np.float32)
    vals, idx, cnt = ops.topk_segments(T(const, cuda), [5000], 1000)
    assert np.array_equal(idx[0].cpu().numpy(), np.arange(1000))


@pytest.mark.parametrize("case", ["random", "ties", "sample_high", "sample_low", "sorted_desc", "sorted_asc", "constant"])
@pytest.mark.parametrize("n,k", [(32768, 100), (100000, 1000), (201600, 2000), (300001, 2048), (150000, 5000)])
def test_topk_sampled_threshold_and_fallback(cuda, case, n, k):
    """Long segments estimate the k-th score from a strided sample (topk.cu, kSample = 16384) and fall back to the full
    select when the estimate misses; both must give the exact (score desc, index asc) top-k."""
    rng = np.random.default_rng(n + k)
    scores = rng.normal(-3, 2, n).astype(np.float32)
    sampled = np.zeros(n, bool)      # positions either sample size (4096 / 16384 keys) would read
    for m in (4096, 16384):
        sampled[(np.arange(m) * (n // m))] = True
    if case == "ties":
        scores = np.round(scores, 1)
    elif case == "sample_high":      # every sampled position is large: the estimated threshold is far too strict
        scores[sampled] += 50.0
    elif case == "sample_low":       # the sample sees none of the top scores: far too many survivors
        scores[sampled] -= 50.0
    elif case == "sorted_desc":
        scores = -np.sort(-scores)
    elif case == "sorted_asc":
        scores = np.sort(scores)
    elif case == "constant":
        scores[:] = 1.5
    for lens in ([n], [n, 777, n]):
        sc = np.concatenate([scores if m == n else scores[:m] for m in lens])
        vals, idx, cnt = ops.topk_segments(T(sc, cuda), lens, k)
        vals, idx, cnt = vals.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy()
        off = 0
        for s, m in enumerate(lens):
            rv, ri = R.topk_desc(sc[off:off + m], k)
            assert cnt[s] == len(ri)
            assert np.array_equal(idx[s, :cnt[s]], ri)
            assert np.array_equal(vals[s, :cnt[s]], rv)
            off += m


# ------------------------------------------------------------------ fused score filter + top-k
def _check_filter(cuda, logits, lens, thr, k, mode, C, ctr=None):
    dense = ops.scores(T(logits, cuda), mode, None if ctr is None else T(ctr, cuda), C).cpu().numpy()
    vals, idx, cnt = ops.score_filter_topk(T(logits, cuda), lens, thr, k, mode, None if ctr is None else T(ctr, cuda), C)
    vals, idx, cnt = vals.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy()
    off = 0
    for s, n in enumerate(lens):
        # bit-identical score tensor handed to the oracle
        keep, sc = R.filter_topk_scores(dense[off:off + n], thr, k)
        assert cnt[s] == keep.size, (s, cnt[s], keep.size)
        assert np.array_equal(idx[s, :cnt[s]], keep)
        assert np.array_equal(vals[s, :cnt[s]], sc)
        off += n
    return dense


def test_score_filter_topk_retinanet_levels(cuda):
    rng = np.random.default_rng(1)
    sizes = W.retinanet_level_sizes(256, 320)
    lens = [h * w * 9 * 80 for h, w in sizes]
    logits = np.concatenate([W.logits_level(rng, h * w * 9, 80).reshape(-1) for h, w in sizes])
    dense = _check_filter(cuda, logits, lens, 0.05, 1000, _lib.SCORE_SIGMOID, 80)
    # the CUDA sigmoid stays within fp32 ulps of the oracle's
    assert np.max(np.abs(dense - R.sigmoid_f32(logits)) / np.maximum(R.sigmoid_f32(logits), 1e-30)) < 1e-6
    # a level with no candidate is empty (retinanet.py:187-188), tiny / ragged segments
    _check_filter(cuda, np.full(1000, -20.0, np.float32), [1000], 0.05, 1000, _lib.SCORE_SIGMOID, 80)
    _check_filter(cuda, logits[:4099 + 77], [4099, 0, 77], 0.05, 50, _lib.SCORE_SIGMOID, 80)
    # many candidates (threshold far below the data): exercises the radix select proper
    _check_filter(cuda, logits[:300000], [300000], 0.0005, 1000, _lib.SCORE_SIGMOID, 80)
    _check_filter(cuda, logits[:5000], [5000], 0.0, 1000, _lib.SCORE_RAW, 1)


def test_score_filter_topk_fcos(cuda):
    rng = np.random.default_rng(2)
    n_pts, C = 22400, 80
    logits = W.logits_level(rng, n_pts, C)
    ctr = rng.normal(0, 1, (n_pts, 1)).astype(np.float32)
    lens = [16800 * C, 4200 * C, 1050 * C, 273 * C, 77 * C]
    dense = _check_filter(cuda, logits.reshape(-1), lens, 0.05, 1000, _lib.SCORE_FCOS, C, ctr.reshape(-1))
    ref = R.fcos_scores(logits, ctr).reshape(-1)
    assert np.max(np.abs(dense - ref) / np.maximum(ref, 1e-30)) < 1e-6


# ------------------------------------------------------------------ NMS
def test_batched_nms_reference_kat(cuda):
    from tests.test_oracle_kat import NMS_BOXES, NMS_LABELS, NMS_SCORES

    keep, cnt = ops.nms_batched(T(NMS_BOXES, cuda)[None], T(NMS_SCORES, cuda)[None], T(NMS_LABELS, cuda)[None], 0.4)
    assert list(keep[0, : int(cnt[0])].cpu().numpy()) == [0, 3, 4, 2]


def _dense_dets(rng, n, num_classes, img=800):
    boxes = W.make_gt(rng, n, img, img, 8, 200)[:, :4]
    # clusters of near-duplicates so that suppression actually happens
    centers = boxes[rng.integers(0, max(n // 8, 1), n)]
    boxes = (centers + rng.normal(0, 4, (n, 4))).astype(np.float32)
    boxes[:, 2:] = np.maximum(boxes[:, 2:], boxes[:, :2] + 1)
    scores = W.distinct_scores(rng, n, 0.05, 1.0)
    labels = rng.integers(0, num_classes, n).astype(np.int32)
    return boxes, scores, labels


@pytest.mark.parametrize("n,ncls,thr,max_out", [(1, 1, 0.5, None), (64, 3, 0.5, None), (65, 1, 0.5, 10), (1000, 80, 0.5, 100),
                                                 (5000, 80, 0.5, 100), (5000, 80, 0.6, None), (8819, 5, 0.7, 1000)])
def test_batched_nms_bit_exact(cuda, n, ncls, thr, max_out):
    rng = np.random.default_rng(n + ncls)
    boxes, scores, labels = _dense_dets(rng, n, ncls)
    keep, cnt = ops.nms_batched(T(boxes, cuda)[None], T(scores, cuda)[None], T(labels, cuda)[None], thr, max_out)
    ref = R.batched_nms(boxes, scores, labels, thr, max_out)
    got = keep[0, : int(cnt[0])].cpu().numpy()
    assert np.array_equal(got, ref)


def test_batched_nms_ragged_batch_float_idxs_and_ties(cuda):
    rng = np.random.default_rng(5)
    B, Nmax = 4, 700
    ns = np.array([700, 1, 0, 333], dtype=np.int32)
    boxes = np.zeros((B, Nmax, 4), np.float32)
    scores = np.zeros((B, Nmax), np.float32)
    levels = np.zeros((B, Nmax), np.float32)
    for b in range(B):
        bx, sc, lb = _dense_dets(rng, Nmax, 5)
        sc = np.round(sc, 2)  # score ties -> stable order (index asc)
        boxes[b], scores[b], levels[b] = bx, sc, lb.astype(np.float32)  # RPN passes float level ids (rpn.py:160)
    keep, cnt = ops.nms_batched(T(boxes, cuda), T(scores, cuda), T(levels, cuda), 0.7, 300, num=T(ns, cuda))
    for b in range(B):
        n = ns[b]
        ref = R.batched_nms(boxes[b, :n], scores[b, :n], levels[b, :n], 0.7, 300)
        assert int(cnt[b]) == len(ref)
        assert np.array_equal(keep[b, : len(ref)].cpu().numpy(), ref)


def test_nms_single_class_vs_reference_py_cpu_nms(cuda):
    """Second opinion: the reference's own numpy helper (post_processing.py:106-132) on tie-free data."""
    rng = np.random.default_rng(9)
    boxes, scores, _ = _dense_dets(rng, 2000, 1)
    keep, cnt = ops.nms_batched(T(boxes, cuda)[None], T(scores, cuda)[None], None, 0.5)
    got = keep[0, : int(cnt[0])].cpu().numpy()
    ref = R.py_cpu_nms(np.concatenate([boxes, scores[:, None]], axis=1), 0.5)
    assert list(got) == [int(i) for i in ref]


def test_nms_large_sort_path(cuda):
    """N > 16384 takes the tiled global bitonic sort."""
    rng = np.random.default_rng(10)
    n = 20000
    boxes, scores, labels = _dense_dets(rng, n, 1)
    keep, cnt = ops.nms_batched(T(boxes, cuda)[None], T(scores, cuda)[None], None, 0.5, None)
    ref = R.nms(boxes, scores, 0.5)
    assert np.array_equal(keep[0, : int(cnt[0])].cpu().numpy(), ref)
    keep, cnt = ops.nms_batched(T(boxes, cuda)[None], T(scores, cuda)[None], T(labels, cuda)[None] * 0 + 2, 0.5, 50)
    ref = R.batched_nms(boxes, scores, labels * 0 + 2, 0.5, 50)
    assert np.array_equal(keep[0, : int(cnt[0])].cpu().numpy(), ref)


@pytest.mark.parametrize("run_lens", [[1000, 1000, 1000, 700, 91], [2000, 2000, 2000, 2000, 819], [0, 5, 0, 0, 0], [3000], [7, 0, 0, 4096, 1]])
def test_nms_presorted_runs_equal_general_sort(cuda, run_lens):
    """bdet_nms_runs (merge by ranking of the per-level top-k runs) == bdet_nms on the same input: sorted runs with
    ties across runs, a broken promise (unsorted run -> device-side check falls back), ragged batch."""
    rng = np.random.default_rng(sum(run_lens))
    n = sum(run_lens)
    B = 3
    boxes = np.zeros((B, n, 4), np.float32)
    scores = np.zeros((B, n), np.float32)
    labels = np.zeros((B, n), np.int32)
    runs = np.zeros((B, len(run_lens)), np.int32)
    num = np.zeros((B,), np.int32)
    for b in range(B):
        bx, sc, lb = _dense_dets(rng, max(n, 1), 4)
        sc = np.round(sc, 2 if b == 1 else 6).astype(np.float32)        # image 1: many equal scores across runs
        lens = list(run_lens) if b < 2 else [m // 2 for m in run_lens]   # image 2: shorter (ragged) runs
        off = 0
        for r, m in enumerate(lens):
            seg = sc[off:off + m]
            order = np.argsort(-seg, kind="stable")
            sc[off:off + m] = seg[order]
            bx[off:off + m] = bx[off:off + m][order]
            lb[off:off + m] = lb[off:off + m][order]
            off += m
            runs[b, r] = off
        num[b] = off
        boxes[b, :n], scores[b, :n], labels[b, :n] = bx[:n], sc[:n], lb[:n]
    args = (T(boxes, cuda), T(scores, cuda), T(labels, cuda), 0.5, 300)
    k0, c0 = ops.nms_batched(*args, num=T(num, cuda))
    k1, c1 = ops.nms_batched(*args, num=T(num, cuda), runs=T(runs, cuda))
    assert np.array_equal(c0.cpu().numpy(), c1.cpu().numpy())
    for b in range(B):
        m = int(c0[b])
        assert np.array_equal(k0[b, :m].cpu().numpy(), k1[b, :m].cpu().numpy())
        ref = R.batched_nms(boxes[b, :num[b]], scores[b, :num[b]], labels[b, :num[b]], 0.5, 300)
        assert np.array_equal(k1[b, :m].cpu().numpy(), ref)
    if n > 10:  # broken promise: shuffle the scores inside the runs
        perm = rng.permutation(n)
        sh = (T(boxes[:, perm], cuda), T(scores[:, perm], cuda), T(labels[:, perm], cuda), 0.5, 300)
        k2, c2 = ops.nms_batched(*sh, num=T(num, cuda))
        k3, c3 = ops.nms_batched(*sh, num=T(num, cuda), runs=T(runs, cuda))
        assert np.array_equal(c2.cpu().numpy(), c3.cpu().numpy())
        for b in range(B):
            assert np.array_equal(k2[b, :int(c2[b])].cpu().numpy(), k3[b, :int(c3[b])].cpu().numpy())
        bad = runs.copy()
        bad[:, -1] += 1                      # inconsistent run table -> general sort as well
        k4, c4 = ops.nms_batched(*args, num=T(num, cuda), runs=T(bad, cuda))
        assert np.array_equal(c0.cpu().numpy(), c4.cpu().numpy())


def test_nms_chunked_path_edge_cases(cuda):
    """Mask/sweep path (max_output x N above the fused limit): max_output reached inside the first chunk, exactly at a
    chunk border, N not a multiple of 64, ragged batch with an empty image, all boxes identical (one survivor)."""
    rng = np.random.default_rng(21)
    n = 3000
    boxes, scores, _ = _dense_dets(rng, n, 1)
    for max_out in (1000, 1024, 1025, None):
        keep, cnt = ops.nms_batched(T(boxes, cuda)[None], T(scores, cuda)[None], None, 0.7, max_out)
        ref = R.nms(boxes, scores, 0.7, max_out)
        assert int(cnt[0]) == len(ref) and np.array_equal(keep[0, : int(cnt[0])].cpu().numpy(), ref)
    b2 = np.stack([boxes, boxes[::-1].copy(), np.tile(boxes[:1], (n, 1))])
    s2 = np.stack([scores, scores[::-1].copy(), scores])
    num = np.array([2999, 0, n], np.int32)
    keep, cnt = ops.nms_batched(T(b2, cuda), T(s2, cuda), None, 0.5, 2000, num=T(num, cuda))
    for b in range(3):
        ref = R.nms(b2[b, : num[b]], s2[b, : num[b]], 0.5, 2000)
        assert int(cnt[b]) == len(ref) and np.array_equal(keep[b, : int(cnt[b])].cpu().numpy(), ref)
    assert int(cnt[2]) == 1
