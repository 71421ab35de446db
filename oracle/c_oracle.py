"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/c/oracle.c (built by `make -C oracle`)."""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_DIR, "_build", "liboracle.so")
_lib = None
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-s", "-C", _DIR], check=True)
        _lib = ctypes.CDLL(_PATH)
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_nms_sorted.restype = ctypes.c_int
    return _lib


def num_threads():
    return load().oracle_num_threads()


def _thr(thresholds):
    return np.array([-np.inf] + [float(t) for t in thresholds] + [np.inf], dtype=np.float32)


def box_iou(b1, b2):
    b1 = np.ascontiguousarray(b1, np.float32)
    b2 = np.ascontiguousarray(b2, np.float32)
    out = np.empty((b1.shape[0], b2.shape[0]), np.float32)
    load().oracle_box_iou(b1.ctypes.data_as(ctypes.c_void_p), b1.shape[1], b1.shape[0], b2.ctypes.data_as(ctypes.c_void_p),
                          b2.shape[1], b2.shape[0], out.ctypes.data_as(ctypes.c_void_p))
    return out


def matcher(m, thresholds, labels, allow_lq):
    m = np.ascontiguousarray(m, np.float32)
    G, A = m.shape
    idx, lab = np.empty(A, np.int32), np.empty(A, np.int32)
    thr, labs = _thr(thresholds), np.asarray(labels, np.int32)
    load().oracle_matcher(m.ctypes.data_as(ctypes.c_void_p), G, A, thr.ctypes.data_as(ctypes.c_void_p),
                          labs.ctypes.data_as(ctypes.c_void_p), len(labs), int(allow_lq),
                          idx.ctypes.data_as(ctypes.c_void_p), lab.ctypes.data_as(ctypes.c_void_p))
    return idx, lab


class TargetScratch:
    def __init__(self, G, A):
        self.iou = np.empty((G, A), np.float32)
        self.idx = np.empty(A, np.int32)
        self.lab = np.empty(A, np.int32)
        self.off = np.empty((A, 4), np.float32)


def retinanet_targets_one(anchors, gt5, thresholds, labels, allow_lq, mean=(0, 0, 0, 0), std=(1, 1, 1, 1), scratch=None):
    """One image of RetinaNet.get_ground_truth (retinanet.py:211-232): materialised IoU -> Matcher -> class -> encode."""
    anchors = np.ascontiguousarray(anchors, np.float32)
    gt5 = np.ascontiguousarray(gt5, np.float32)
    A, G = anchors.shape[0], gt5.shape[0]
    s = scratch or TargetScratch(G, A)
    thr, labs = _thr(thresholds), np.asarray(labels, np.int32)
    mean, std = np.asarray(mean, np.float32), np.asarray(std, np.float32)
    vp = ctypes.c_void_p
    load().oracle_retinanet_targets(anchors.ctypes.data_as(vp), A, gt5.ctypes.data_as(vp), G, thr.ctypes.data_as(vp),
                                    labs.ctypes.data_as(vp), len(labs), int(allow_lq), mean.ctypes.data_as(vp),
                                    std.ctypes.data_as(vp), s.iou.ctypes.data_as(vp), s.idx.ctypes.data_as(vp),
                                    s.lab.ctypes.data_as(vp), s.off.ctypes.data_as(vp))
    return s.lab, s.off, s.idx


def nms(boxes, scores, iou_thresh, max_output=None):
    boxes = np.asarray(boxes, np.float32)
    scores = np.asarray(scores, np.float32)
    order = np.argsort(-scores, kind="stable").astype(np.int32)
    sb = np.ascontiguousarray(boxes[order])
    kept = np.empty(len(order), np.int32)
    n = load().oracle_nms_sorted(sb.ctypes.data_as(ctypes.c_void_p), len(order), ctypes.c_float(iou_thresh),
                                 int(max_output or 0), kept.ctypes.data_as(ctypes.c_void_p))
    return order[kept[:n]]


def roi_align_fwd(feat, rois, pool, scale, samples=(2, 2), aligned=True):
    feat = np.ascontiguousarray(feat, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    B, C, H, W = feat.shape
    out = np.empty((rois.shape[0], C, pool[0], pool[1]), np.float32)
    vp = ctypes.c_void_p
    load().oracle_roi_align_fwd(feat.ctypes.data_as(vp), B, C, H, W, rois.ctypes.data_as(vp), rois.shape[0], pool[0], pool[1],
                                samples[0], samples[1], ctypes.c_float(scale), ctypes.c_float(0.5 if aligned else 0.0),
                                out.ctypes.data_as(vp))
    return out


def roi_align_bwd(dout, feat_shape, rois, pool, scale, samples=(2, 2), aligned=True):
    dout = np.ascontiguousarray(dout, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    B, C, H, W = feat_shape
    grad = np.zeros(feat_shape, np.float64)
    vp = ctypes.c_void_p
    load().oracle_roi_align_bwd(dout.ctypes.data_as(vp), B, C, H, W, rois.ctypes.data_as(vp), rois.shape[0], pool[0], pool[1],
                                samples[0], samples[1], ctypes.c_float(scale), ctypes.c_float(0.5 if aligned else 0.0),
                                grad.ctypes.data_as(vp))
    return grad
