"""ctypes binding of libbdet.so (include/bdet.h).  No CPU fallback: a missing library is an error."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbdet.so")

BDET_OK = 0
PAIR_IOU, PAIR_IOA, PAIR_INTER, PAIR_GIOU = 0, 1, 2, 3
SCORE_RAW, SCORE_SIGMOID, SCORE_FCOS = 0, 1, 2
MAX_LEVELS = 8

vp = c_void_p
ip = POINTER(c_int)
fp = POINTER(c_float)
dp = POINTER(c_double)
lp = POINTER(c_int64)

# name -> (restype, argtypes); mirrors include/bdet.h one to one (tests/test_abi.py checks the header).
SIGNATURES = {
    "bdet_abi_version": (c_int, []),
    "bdet_last_error": (c_char_p, []),
    "bdet_device_info": (c_int, [ip, ip, ip]),
    "bdet_anchors_grid": (c_int, [vp, c_int, ip, dp, dp, ip, fp, lp, vp]),
    "bdet_points_grid": (c_int, [vp, c_int, ip, dp, dp, c_int, c_int, lp, vp]),
    "bdet_pairwise": (c_int, [vp, c_int, c_int, vp, c_int, c_int, vp, c_int64, c_int, vp]),
    "bdet_pairwise_batched": (c_int, [vp, c_int, c_int64, vp, c_int, vp, c_int, c_int64, c_int, vp, c_int64, c_int64,
                                      c_int, c_int, vp]),
    "bdet_box_center": (c_int, [vp, c_int, c_int, vp, vp]),
    "bdet_point_distance": (c_int, [vp, c_int, vp, c_int, vp, vp]),
    "bdet_match_workspace": (c_size_t, [c_int, c_int, c_int]),
    "bdet_match": (c_int, [vp, c_int64, c_int64, vp, c_int, c_int, c_int, fp, ip, c_int, c_int, vp, vp, vp, c_size_t, vp]),
    "bdet_match_rows": (c_int, [vp, c_int, c_int, vp, vp, vp]),
    "bdet_box_encode": (c_int, [vp, vp, c_int, vp, c_int, fp, fp, vp, vp]),
    "bdet_box_decode": (c_int, [vp, vp, c_int, c_int, fp, fp, vp, c_int, vp, c_int, c_int, vp]),
    "bdet_sum_encode": (c_int, [vp, vp, c_int, fp, fp, vp, vp]),
    "bdet_sum_decode": (c_int, [vp, vp, c_int, fp, fp, vp, c_int, vp]),
    "bdet_point_encode": (c_int, [vp, c_int, vp, c_int, c_int, vp, vp]),
    "bdet_point_encode_rows": (c_int, [vp, vp, c_int, c_int, vp, vp]),
    "bdet_point_decode": (c_int, [vp, vp, c_int, c_int, vp, vp, c_int, c_int, vp]),
    "bdet_assign_targets_workspace": (c_size_t, [c_int, c_int, c_int]),
    "bdet_assign_targets": (c_int, [vp, c_int, vp, c_int, vp, c_int, fp, ip, c_int, c_int, c_int, fp, fp,
                                    vp, vp, vp, vp, c_size_t, vp]),
    "bdet_assign_targets_grid": (c_int, [c_int, ip, dp, dp, ip, fp, vp, c_int, vp, c_int, fp, ip, c_int, c_int, c_int, fp, fp,
                                         vp, vp, vp, vp, vp, c_size_t, vp]),
    "bdet_topk_workspace": (c_size_t, [c_int64, c_int, c_int]),
    "bdet_topk": (c_int, [vp, lp, lp, c_int, c_int, vp, vp, vp, vp, c_size_t, vp]),
    "bdet_score_filter_topk_workspace": (c_size_t, [c_int64, c_int, c_int]),
    "bdet_score_filter_topk": (c_int, [vp, vp, c_int, lp, lp, lp, c_int, c_float, c_int, c_int, vp, vp, vp, vp,
                                       c_size_t, vp]),
    "bdet_select_decode": (c_int, [POINTER(vp), POINTER(vp), ip, c_int, c_int, c_int, c_int, c_int, c_int, vp, vp, vp,
                                   fp, fp, vp, c_int, vp, vp, vp, vp, vp, vp]),
    "bdet_finalize_detections": (c_int, [vp, vp, vp, c_int, c_int, vp, c_int, vp, vp, c_int, c_int, c_int, c_int, vp,
                                         vp]),
    "bdet_scores": (c_int, [vp, vp, c_int, c_int64, c_int, vp, vp]),
    "bdet_nms_workspace": (c_size_t, [c_int, c_int]),
    "bdet_nms": (c_int, [vp, vp, vp, c_int, vp, c_int, c_int, c_float, c_int, vp, c_int, vp, vp, c_size_t, vp]),
    "bdet_nms_runs": (c_int, [vp, vp, vp, c_int, vp, vp, c_int, c_int, c_int, c_float, c_int, vp, c_int, vp, vp,
                              c_size_t, vp]),
    "bdet_boxes_scale_clip": (c_int, [vp, c_int, c_float, c_float, c_float, c_float, vp]),
    "bdet_boxes_filter_by_size": (c_int, [vp, c_int, c_float, c_float, vp, vp]),
    "bdet_roi_assign_levels": (c_int, [vp, c_int, c_int, c_int, vp, vp]),
    "bdet_roi_align_fwd": (c_int, [POINTER(vp), c_int, ip, fp, c_int, c_int, vp, vp, c_int, c_int, c_int,
                                   c_int, c_int, c_int, vp, vp]),
    "bdet_roi_align_bwd_workspace": (c_size_t, [c_int, ip, c_int, c_int]),
    "bdet_roi_align_bwd": (c_int, [POINTER(vp), c_int, ip, fp, c_int, c_int, vp, vp, c_int, c_int, c_int,
                                   c_int, c_int, c_int, vp, c_int, vp, c_size_t, vp]),
    "bdet_roi_order": (c_int, [c_int, ip, fp, c_int, vp, vp, c_int, c_int, c_int, c_int, vp, vp]),
    "bdet_roi_align_fwd_perm": (c_int, [POINTER(vp), c_int, ip, fp, c_int, c_int, vp, vp, c_int, c_int, c_int,
                                        c_int, c_int, c_int, vp, vp, vp]),
    "bdet_roi_align_bwd_perm": (c_int, [POINTER(vp), c_int, ip, fp, c_int, c_int, vp, vp, c_int, c_int, c_int,
                                        c_int, c_int, c_int, vp, c_int, vp, vp, c_size_t, vp]),
    "bdet_roi_maxpool_fwd": (c_int, [POINTER(vp), c_int, ip, fp, c_int, c_int, vp, vp, c_int, c_int, c_int, vp, vp, vp]),
    "bdet_roi_maxpool_bwd": (c_int, [POINTER(vp), c_int, ip, c_int, c_int, vp, vp, c_int, c_int, c_int, vp, vp, c_int, vp]),
    "bdet_box_props": (c_int, [vp, c_int, c_int, c_int, vp, vp]),
    "bdet_box_convert": (c_int, [vp, c_int, c_int, c_int, vp, vp]),
    "bdet_cond_take_workspace": (c_size_t, [c_int64]),
    "bdet_cond_take": (c_int, [vp, vp, c_int64, vp, vp, vp, vp, c_size_t, vp]),
    "bdet_count_labels": (c_int, [vp, c_int, c_int, vp, vp]),
    "bdet_bw_probe": (c_int, [vp, vp, c_size_t, c_int, c_int, vp]),
    "bdet_fcos_targets": (c_int, [vp, c_int, ip, c_int, fp, fp, fp, vp, c_int, vp, c_int, vp, vp, vp, vp, vp]),
    "bdet_atss_targets_workspace": (c_size_t, [c_int, c_int]),
    "bdet_atss_targets": (c_int, [vp, c_int, ip, c_int, fp, c_int, vp, c_int, vp, c_int, vp, vp, vp, vp, vp, c_size_t, vp]),
    "bdet_sample_labels": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, vp, vp]),
    "bdet_rcnn_match": (c_int, [vp, vp, c_int, vp, vp, c_int, c_int, c_float, c_float, c_float, vp, vp, vp, vp, vp, vp, vp]),
    "bdet_rcnn_collect": (c_int, [vp, vp, vp, vp, vp, vp, c_int, vp, c_int, c_int, fp, fp, c_int, vp, vp, vp, vp, vp]),
    "bdet_score_filter_topk_nchw": (c_int, [vp, vp, c_int, c_int, lp, ip, lp, c_int, c_float, c_int, c_int, vp, vp, vp, vp,
                                            c_size_t, vp]),
    "bdet_select_decode_nchw": (c_int, [POINTER(vp), POINTER(vp), ip, ip, c_int, c_int, c_int, c_int, c_int, c_int, vp, vp, vp,
                                        fp, fp, vp, c_int, vp, vp, vp, vp, vp, vp]),
    "bdet_ota_topk_match_workspace": (c_size_t, [c_int]),
    "bdet_ota_topk_match": (c_int, [vp, c_int, vp, c_int, c_int, c_int, c_int, vp, vp, c_size_t, vp]),
    "bdet_ota_cost_workspace": (c_size_t, [c_int]),
    "bdet_ota_cost": (c_int, [vp, vp, c_int, vp, c_int, vp, c_int, vp, c_double, c_double, c_double, vp, vp, vp, c_size_t, vp]),
    "bdet_ota_collect": (c_int, [vp, vp, c_int, vp, c_int, vp, vp, vp, vp, vp]),
    "bdet_free_anchor_box_prob_workspace": (c_size_t, [c_int]),
    "bdet_free_anchor_box_prob": (c_int, [vp, c_int, vp, c_int, c_int, c_float, c_float, c_float, vp, vp, c_size_t, vp]),
    "bdet_free_anchor_bags": (c_int, [vp, c_int, c_int, vp, vp, vp, c_int, fp, fp, vp, vp, vp]),
    "bdet_dense_tail_smem": (c_size_t, [c_int, c_int, c_int]),
    "bdet_dense_tail": (c_int, [POINTER(vp), POINTER(vp), ip, ip, c_int, c_int, c_int, c_int, c_int, vp, vp, vp, fp, fp, vp,
                                c_int, c_float, c_int, vp, vp, vp]),
    "bdet_coco_format": (c_int, [vp, vp, c_int, c_int, vp, vp, c_int, vp, vp, vp, vp, vp, vp]),
    "bdet_select_decode_workspace": (c_size_t, [c_int, c_int, c_int]),
    "bdet_select_decode_ws": (c_int, [POINTER(vp), POINTER(vp), ip, ip, c_int, c_int, c_int, c_int, c_int, c_int, vp, vp, vp,
                                      fp, fp, vp, c_int, vp, vp, vp, vp, vp, vp, c_size_t, vp]),
    "bdet_profile_begin": (c_int, []),
    "bdet_profile_select": (c_int, [c_char_p]),
    "bdet_profile_collect": (c_int, [c_char_p, fp, ip]),
    "bdet_profile_report": (c_int, [c_char_p, c_size_t]),
    "bdet_profile_end": (c_int, []),
}


class BdetError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbdet error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load libbdet.so (built by ``basedet_b200.build``).  Raises if it is missing -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libbdet.so not found at %s: run `python -m basedet_b200.build` (nvcc, sm_100a). "
            "basedet_b200 has no CPU or PyTorch fallback." % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.bdet_abi_version() != 1:
        raise RuntimeError("libbdet.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != BDET_OK:
        raise BdetError(rc, load().bdet_last_error().decode("utf-8", "replace"))


def farr(values):
    return (c_float * len(values))(*[float(v) for v in values])


def iarr(values):
    return (c_int * len(values))(*[int(v) for v in values])


def darr(values):
    return (c_double * len(values))(*[float(v) for v in values])


def larr(values):
    return (c_int64 * len(values))(*[int(v) for v in values])
