"""GPU arms of bench.py: one object per BASELINE.json config, each holding device-resident synthetic inputs
for a list of GLOBAL image indices, the step (eager and as one CUDA graph), the host<->device traffic of the end-to-end
variant and the algorithmic byte counts of its kernels.

Sharding: the caller passes the image indices of its rank (distributed.shard_range); inputs are a function
of the global image index (workloads.*_image), so an N-rank run computes exactly what the 1-rank run computes.
Nothing here touches oracle/ -- the CPU arms live in bench.py / oracle/cpu_arms.py."""
import math

import numpy as np
import torch

from . import _lib, ops, pipelines
from . import workloads as W
from .layers import DefaultAnchorGenerator


def _T(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _pin(x):
    return torch.from_numpy(np.ascontiguousarray(x)).pin_memory()


def _nbytes(ts):
    return int(sum(t.numel() * t.element_size() for t in ts))


class Arm:
    name = ""
    workload = ""
    dominant = ""
    rate_unit = None          # e.g. "pair-evaluations/s" for issue-bound kernels: (name, units per launch)

    def __init__(self, images, dev):
        self.images = list(images)
        self.B = len(self.images)
        self.dev = dev
        self.graph = None
        self.h2d = []             # (pinned host tensor, device tensor) copied every end-to-end step
        self.summary = None       # small device tensor read back every end-to-end step (written by the step)
        self.kernel_bytes = {}    # kernel name -> algorithmic HBM bytes per launch
        self.kernel_units = {}    # kernel name -> (unit name, units per launch) for issue-bound kernels
        self.resident_note = ""

    # -- to be provided
    def eager(self):
        raise NotImplementedError

    # -- common
    def capture(self):
        """Warm up on a side stream, then capture ``eager`` as ONE CUDA graph (falls back to eager launches)."""
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.eager()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.eager()
            self.graph = g
        except Exception as e:  # noqa: BLE001 -- a step that cannot be captured is still measured, eagerly
            self.graph = None
            self.capture_error = repr(e)[:200]
            torch.cuda.synchronize(self.dev)
        return self.graph is not None

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.eager()

    def e2e_step(self, out_host):
        for h, d in self.h2d:
            d.copy_(h, non_blocking=True)
        self.step()
        out_host.copy_(self.summary, non_blocking=True)

    @property
    def h2d_bytes(self):
        return _nbytes([h for h, _ in self.h2d])

    @property
    def d2h_bytes(self):
        return _nbytes([self.summary])


# ------------------------------------------------------------------------------------------------------ config 1
class RetinaNetPostprocess(Arm):
    """configs[0]: RetinaNet R50-FPN post-processing of 800x800 images (A = 120 087 anchors, 80 classes):
    anchors -> sigmoid > 0.05 -> top-1000 / level -> BoxCoder.decode -> batched NMS 0.5 -> 100 dets -> scale / clip
    (models/det/retinanet.py:172-209, layers/common/post_processing.py:50-103)."""
    name = "c1_retinanet_postprocess"
    dominant = "score_filter_kernel"

    def __init__(self, images, dev, hw=(800, 800)):
        super().__init__(images, dev)
        self.sizes = W.retinanet_level_sizes(*hw)
        self.gen = DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, W.RETINANET_OFFSET)
        per = [W.retinanet_image(i, hw) for i in self.images]
        L = len(self.sizes)
        self.logits = [_T(np.stack([p["logits"][l] for p in per]), dev) for l in range(L)]
        self.offsets = [_T(np.stack([p["offsets"][l] for p in per]), dev) for l in range(L)]
        info = np.stack([p["im_info"] for p in per])
        self.im_info = _T(info, dev)
        self.h2d = [(_pin(info), self.im_info)]
        self.A = sum(h * w * 9 for h, w in self.sizes)
        self.summary = torch.zeros((self.B, 100 * 6 + 1), dtype=torch.float32, device=dev)
        self.workload = ("configs[0]: RetinaNet post-processing, %d image(s) %dx%d, A=%d anchors x 80 classes, thr 0.05, "
                         "top-1000/level, NMS 0.5, 100 dets" % (self.B, hw[0], hw[1], self.A))
        self.kernel_bytes = {"score_filter_kernel": self.B * self.A * 80 * 4}
        self.resident_note = "head outputs (%.1f MB logits + offsets) are network outputs and stay on the device; the full " \
                             "result (100 x 6 floats + count per image) is copied back" % (self.B * self.A * 84 * 4 / 1e6)

    def eager(self):
        anchors = self.gen.generate_anchors_by_features(self.sizes, self.dev)
        dets, cnt = pipelines.dense_postprocess(self.logits, self.offsets, anchors, self.im_info, 0.05, 0.5, 100, 1000)
        self.summary[:, :600].copy_(dets.reshape(self.B, 600))
        self.summary[:, 600].copy_(cnt)
        self.out = (dets, cnt)
        return self.out


# ------------------------------------------------------------------------------------------------------ config 2
class RetinaNetTargets(Arm):
    """configs[1]: RetinaNet training target assignment: anchors (regenerated per forward, retinanet.py:116) x GT IoU
    + Matcher(.4/.5, low-quality) + class labels + BoxCoder.encode + label census (models/det/retinanet.py:211-232,
    142-146), fused (bdet_assign_targets; bdet_assign_targets_grid for <= 4 images per GPU)."""
    name = "c2_retinanet_targets"
    dominant = "assign_main_kernel"

    def __init__(self, images, dev, hw=(800, 800), num_gt=100):
        super().__init__(images, dev)
        self.sizes = W.retinanet_level_sizes(*hw)
        self.gen = DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, W.RETINANET_OFFSET)
        gt = np.stack([W.target_assign_batch(1, num_gt, hw[0], hw[1], seed0=100 + i)[0][0] for i in self.images])
        ng = np.full((self.B,), num_gt, np.int32)
        self.gt_np, self.ng_np = gt, ng
        self.gt, self.ng = _T(gt, dev), _T(ng, dev)
        self.h2d = [(_pin(gt), self.gt), (_pin(ng), self.ng)]
        self.A, self.G = sum(h * w * 9 for h, w in self.sizes), num_gt
        self.plan = ops.AssignPlan(self.A, num_gt, self.B, dev)
        self.summary = torch.zeros((self.B, 3), dtype=torch.int32, device=dev)
        self.workload = ("configs[1]: RetinaNet target assignment, %d image(s), A=%d anchors (800x800), G=%d GT: anchors_grid "
                         "+ IoU + Matcher(.4/.5, low-quality) + class labels + BoxCoder.encode" % (self.B, self.A, num_gt))
        self.kernel_bytes = {"assign_main_kernel": self.A * 16 + self.B * (self.A * 24 + num_gt * 20)}
        self.kernel_units = {"assign_main_kernel": ("pair-evaluations/s", float(self.B) * num_gt * self.A)}
        self.resident_note = "labels / match indices / offsets (%.1f MB) stay on the device for the loss " \
                             "(retinanet.py:151-162); only the per-image label census (num_fg normaliser) is copied back" \
                             % (self.B * self.A * 24 / 1e6)

    def eager(self):
        # anchors are regenerated per forward in the reference (retinanet.py:116)
        m = W.RETINANET_MATCHER
        if self.B <= pipelines.GRID_ASSIGN_MAX_BATCH:
            # few images per GPU (the strong-scaling regime): the step is launch bound, so the anchors are generated in
            # registers inside the assignment kernels together with the label census -- two launches instead of four
            out = ops.assign_targets_grid(self.gen._plan(self.sizes), self.gt, self.ng, m["thresholds"], m["labels"],
                                          m["allow_low_quality"], True, plan=self.plan, counts=self.summary)
        else:
            # larger batches are issue bound: one anchor tensor shared by all images is cheaper than regenerating it per image
            anchors = self.gen.generate_all_level_anchors(self.sizes, self.dev)
            out = ops.assign_targets(anchors, self.gt, self.ng, m["thresholds"], m["labels"], m["allow_low_quality"], True,
                                     plan=self.plan)
            ops.count_labels(self.plan.labels, out=self.summary)
        self.out = out
        return out


# ------------------------------------------------------------------------------------------------------ config 3
def roi_footprint_union_bytes(rois, levels, hw_list, scales, B, C, P=7, S=2):
    """Bytes of the feature pyramid that the ROIAlign samples of ``rois`` touch at least once: per (image, level) the
    union of the ROI footprints (rows / columns floor(first sample) .. floor(last sample) + 1, clipped) x C x 4."""
    rois = np.asarray(rois, np.float64)
    levels = np.asarray(levels)
    total = 0
    f0, f1 = 0.5 / S, (S - 0.5) / S
    for l, (H, Wd) in enumerate(hw_list):
        cover = np.zeros((B, H + 1, Wd + 1), np.int32)
        for r in rois[levels == l]:
            n = int(r[0])
            if not 0 <= n < B:
                continue
            sx, sy = r[1] * scales[l] - 0.5, r[2] * scales[l] - 0.5
            bw, bh = max(r[3] * scales[l] - 0.5 - sx, 0.0) / P, max(r[4] * scales[l] - 0.5 - sy, 0.0) / P
            x0, x1 = math.floor(sx + bw * f0), math.floor(sx + bw * (P - 1 + f1)) + 1
            y0, y1 = math.floor(sy + bh * f0), math.floor(sy + bh * (P - 1 + f1)) + 1
            x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, Wd - 1), min(y1, H - 1)
            if x0 > x1 or y0 > y1:
                continue
            cover[n, y0, x0] += 1
            cover[n, y1 + 1, x1 + 1] += 1
            cover[n, y0, x1 + 1] -= 1
            cover[n, y1 + 1, x0] -= 1
        touched = (cover.cumsum(1).cumsum(2)[:, :H, :Wd] > 0).sum()
        total += int(touched) * C * 4
    return total


class FasterRCNNTrainBoxOps(Arm):
    """configs[2]: every box op of a Faster R-CNN R50-FPN training step at 800x1344 (pipelines.frcnn_train_box_ops):
    RPN anchors + proposals (top-2000 / level, decode, clip / filter, NMS 0.7 -> 1000) + RPN targets (IoU, Matcher,
    encode, sampling) + RCNN targets (IoU, argmax, sampling -> 512 rois / image) + ROIAlign 7x7 forward and backward on
    the 256-channel pyramid."""
    name = "c3_faster_rcnn_train"
    dominant = "roi_align_bwd_kernel"

    def __init__(self, images, dev, hw=W.FRCNN_HW, channels=W.FRCNN_CHANNELS):
        super().__init__(images, dev)
        self.hw, self.C = hw, channels
        self.sizes = W.frcnn_level_sizes(*hw)
        self.gen = DefaultAnchorGenerator(W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, W.FRCNN_OFFSET)
        b = W.frcnn_batch(self.images, hw)
        self.host = b
        self.scores = [_T(x, dev) for x in b["scores"]]
        self.deltas = [_T(x, dev) for x in b["deltas"]]
        self.gt, self.ng, self.im_info = _T(b["gt"], dev), _T(b["num_gt"], dev), _T(b["im_info"], dev)
        self.noise_rpn, self.noise_rcnn = _T(b["noise_rpn"], dev), _T(b["noise_rcnn"], dev)
        self.h2d = [(_pin(b["gt"]), self.gt), (_pin(b["num_gt"]), self.ng), (_pin(b["im_info"]), self.im_info)]
        self.fsizes = [(-(-hw[0] // s), -(-hw[1] // s)) for s in W.FRCNN_RCNN_STRIDES]
        # features / dout are network activations: values do not steer control flow; generated on the device per image
        self.features = [torch.empty((self.B, channels, h, w), dtype=torch.float32, device=dev) for h, w in self.fsizes]
        K = self.B * W.FRCNN_NUM_ROIS
        self.dout = torch.empty((K, channels, 7, 7), dtype=torch.float32, device=dev)
        for j, img in enumerate(self.images):
            g = torch.Generator(device=dev)
            g.manual_seed(7000 + img)
            for f in self.features:
                f[j].normal_(0.0, 1.0, generator=g)
            self.dout[j * W.FRCNN_NUM_ROIS:(j + 1) * W.FRCNN_NUM_ROIS].normal_(0.0, 1.0, generator=g)
        self.dfeats = [torch.empty_like(f) for f in self.features]
        self.A = sum(h * w * 3 for h, w in self.sizes)
        self.plan = ops.AssignPlan(self.A, b["gt"].shape[1], self.B, dev)
        self.summary = torch.zeros((self.B, 5), dtype=torch.int32, device=dev)
        self.K = K
        self.pyramid_bytes = _nbytes(self.features)
        self.out_bytes = K * channels * 49 * 4
        self.workload = ("configs[2]: Faster R-CNN R50-FPN training step box ops, %d image(s) 800x1344: anchors (A=%d) + RPN "
                         "top-2000/level + decode + NMS 0.7 -> 1000 proposals + RPN targets (IoU %dx100, Matcher .3/.7, "
                         "sampling 256) + RCNN targets (1100x100 IoU, sampling 512) + ROIAlign 7x7 fwd + bwd on 512 rois/img, "
                         "256-ch P2-P5" % (self.B, self.A, self.A))
        self.kernel_units = {"assign_main_kernel": ("pair-evaluations/s", float(self.B) * 100 * self.A)}
        self.resident_note = ("RPN head outputs (%.1f MB), the feature pyramid (%.0f MB), dout (%.0f MB) are network "
                              "activations and stay on the device, as do all results (proposals, targets, pooled %.0f MB, "
                              "dfeat %.0f MB); per-image counts (proposals, sampled rois, label census) are copied back"
                              % (_nbytes(self.scores + self.deltas) / 1e6, self.pyramid_bytes / 1e6, self.out_bytes / 1e6,
                                 self.out_bytes / 1e6, self.pyramid_bytes / 1e6))

    def eager(self):
        o = pipelines.frcnn_train_box_ops(self.gen, self.sizes, self.scores, self.deltas, self.features, self.gt, self.ng,
                                          self.im_info, self.noise_rpn, self.noise_rcnn, dout=self.dout, plan=self.plan,
                                          dfeats=self.dfeats)
        self.summary[:, 0].copy_(o["n_rois"])
        self.summary[:, 1].copy_(o["rcnn_count"])
        ops.count_labels(o["rpn_labels"], out=self._census())
        self.summary[:, 2:5].copy_(self._census_t)
        self.out = o
        return o

    def _census(self):
        if not hasattr(self, "_census_t"):
            self._census_t = torch.zeros((self.B, 3), dtype=torch.int32, device=self.dev)
        return self._census_t

    def finish_setup(self):
        """Byte counts that depend on the rois the pipeline itself produced (after one eager run)."""
        o = self.eager()
        torch.cuda.synchronize(self.dev)
        rois = o["rcnn_rois"].reshape(-1, 5).cpu().numpy()
        lv = o["levels"].cpu().numpy()
        u = roi_footprint_union_bytes(rois, lv, self.fsizes, [1.0 / s for s in W.FRCNN_RCNN_STRIDES], self.B, self.C)
        self.union_bytes = u
        self.level_hist = np.bincount(lv, minlength=len(self.fsizes)).tolist()
        self.kernel_bytes = {
            # out written once + every touched feature element read at least once
            "roi_align_fwd_kernel": self.out_bytes + self.K * 20 + u,
            "roi_align_fwd_tma_kernel": self.out_bytes + self.K * 20 + u,
            # dout read once + every touched dfeat element read-modify-written at least once
            "roi_align_bwd_kernel": self.out_bytes + self.K * 20 + 2 * u,
            "roi_align_bwd_tma_kernel": self.out_bytes + self.K * 20 + 2 * u,
        }


# ------------------------------------------------------------------------------------------------------ config 4
class FCOSPostprocess(Arm):
    """configs[3]: FCOS / ATSS dense inference at 1333x800 (22 400 points, 80 classes): sqrt(sigmoid(cls) *
    sigmoid(ctr)) > 0.05 -> top-1000 / level -> PointCoder.decode -> class-aware batched NMS 0.6 -> 100 dets
    (models/det/fcos.py:191-221, layers/common/post_processing.py:50-103)."""
    name = "c4_fcos_postprocess"
    dominant = "score_filter_kernel"

    def __init__(self, images, dev, hw=W.FCOS_HW):
        super().__init__(images, dev)
        self.sizes = W.retinanet_level_sizes(*hw)
        b = W.fcos_batch(self.images, hw)
        self.logits = [_T(x, dev) for x in b["logits"]]
        self.ctr = [_T(x, dev) for x in b["ctrness"]]
        self.offsets = [_T(x, dev) for x in b["offsets"]]
        self.im_info = _T(b["im_info"], dev)
        self.h2d = [(_pin(b["im_info"]), self.im_info)]
        self.P = sum(h * w for h, w in self.sizes)
        self.summary = torch.zeros((self.B, 601), dtype=torch.float32, device=dev)
        self.workload = ("configs[3]: FCOS dense inference, %d image(s) 800x1344, %d points x 80 classes, "
                         "sqrt(sig(cls)*sig(ctr)) > 0.05, top-1000/level, PointCoder.decode, NMS 0.6, 100 dets" % (self.B, self.P))
        self.kernel_bytes = {"score_filter_kernel": self.B * self.P * (80 + 1) * 4}
        self.resident_note = "head outputs (%.1f MB) stay on the device; the full result (100 x 6 floats + count per image) " \
                             "is copied back" % (_nbytes(self.logits + self.ctr + self.offsets) / 1e6)

    def eager(self):
        pts = ops.points_grid(self.sizes, W.RETINANET_STRIDES, [0.5 * s for s in W.RETINANET_STRIDES], 1, 0, self.dev)
        dets, cnt = pipelines.dense_postprocess(self.logits, self.offsets, pts, self.im_info, 0.05, W.FCOS_NMS_THR, 100, 1000,
                                                ctrness_list=self.ctr)
        self.summary[:, :600].copy_(dets.reshape(self.B, 600))
        self.summary[:, 600].copy_(cnt)
        self.out = (dets, cnt)
        return self.out


# ------------------------------------------------------------------------------------------------------ config 5
class CrowdedStress(Arm):
    """configs[4]: crowded-scene stress: per image a (500 x 200 000) IoU matrix + Matcher, and single-class NMS 0.5 over
    100 000 boxes with no output cap (structures/op_patch.py:33-97, layers/common/matcher.py:31-51,
    layers/common/post_processing.py:17-47)."""
    name = "c5_crowded_stress"
    dominant = "pairwise_kernel"

    def __init__(self, images, dev, n_boxes=100000, n_anchor=200000, n_gt=500):
        super().__init__(images, dev)
        per = [W.stress_image(i, n_boxes, n_anchor, n_gt) for i in self.images]
        self.host = per
        self.anchors = _T(np.stack([p["anchors"] for p in per]), dev)
        gt = np.stack([p["gt"] for p in per])
        self.gt = _T(gt, dev)
        self.boxes = _T(np.stack([p["boxes"] for p in per]), dev)
        self.scores = _T(np.stack([p["scores"] for p in per]), dev)
        self.h2d = [(_pin(gt), self.gt)]
        self.N, self.A, self.G = n_boxes, n_anchor, n_gt
        self.iou = ops._padded_rows((self.B, n_gt), n_anchor, dev)[0]
        self.nms_ws = ops._workspace(_lib.load().bdet_nms_workspace(n_boxes, self.B), dev)
        self.summary = torch.zeros((self.B, 4), dtype=torch.int32, device=dev)
        self.workload = ("configs[4]: crowded stress, %d image(s): IoU %d x %d + Matcher(.4/.5, low-quality), single-class "
                         "NMS 0.5 over %d boxes, no output cap" % (self.B, n_gt, n_anchor, n_boxes))
        self.kernel_bytes = {"pairwise_kernel": self.B * (4 * n_gt * n_anchor + 16 * (n_gt + n_anchor)),
                             "match_colmax_kernel": self.B * (4 * n_gt * n_anchor + 8 * n_anchor)}
        self.resident_note = "the IoU matrices (%.0f MB), match results and keep lists stay on the device; per-image keep " \
                             "counts and the label census are copied back" % (self.B * n_gt * n_anchor * 4 / 1e6)

    def eager(self):
        ops.pairwise_batched(self.gt, None, self.anchors, out=self.iou)
        idx, lab = ops.match(self.iou, [0.4, 0.5], [0, -1, 1], True)
        keep, cnt = ops.nms_batched(self.boxes, self.scores, None, 0.5, None, workspace=self.nms_ws)
        self.summary[:, 0].copy_(cnt)
        ops.count_labels(lab, out=self._census())
        self.summary[:, 1:4].copy_(self._census_t)
        self.out = (idx, lab, keep, cnt)
        return self.out

    def _census(self):
        if not hasattr(self, "_census_t"):
            self._census_t = torch.zeros((self.B, 3), dtype=torch.int32, device=self.dev)
        return self._census_t


CONFIG_BATCH = {"c1": 1, "c2": 16, "c3": 16, "c4": 64, "c5": 8}   # BASELINE.json configs[0..4]
ARMS = {"c1": RetinaNetPostprocess, "c2": RetinaNetTargets, "c3": FasterRCNNTrainBoxOps, "c4": FCOSPostprocess,
        "c5": CrowdedStress}
