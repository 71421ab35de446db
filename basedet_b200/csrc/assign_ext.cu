// 8(f)-3 / 8(f)-4, the rest of the widened rows:
//   FreeAnchor box ops          basedet/models/det/free_anchor.py:48-113
//     fa_rowmax_kernel / fa_scatter_kernel   box-probability scatter (:54-84) without the (G, A) matrices
//     fa_bag_kernel                          bag scores + BoxCoder targets of the per-GT top-k anchors (:95-113)
//   OTA cost construction       basedet/models/det/ota.py:91-152 (+ layers/losses/{sigmoid_focal_loss,iou_loss}.py)
//     ota_bg_kernel / ota_cost_kernel        cost (G, A) and IoU (G, A) for OTATopkMatcher (ota_match.cu)
//     ota_collect_kernel                     class / box / IoU targets of the matched anchors (:160-175)
//   COCO result records         basedet/evaluators/coco_eval.py:111-138
//     coco_format_kernel                     padded detections -> compact (image_id, xywh, score, category_id) records
#include "common.cuh"

namespace bdet {

// ---------------------------------------------------------------------------------------------------------------
// FreeAnchor box probabilities.  image_boxes_prob[a, label[g]] = clip((IoU(gt g, pred a) - t1) / (t2[g] - t1), 0, 1)
// for the non-zero entries, written in ascending (g, a) order -- the last GT wins a collision (oracle ASSUMED-11).
// t2[g] = clip(max_a IoU(g, a), t1 + eps, 1).  Pairs whose boxes do not overlap have IoU 0 and (t1 >= 0) probability 0:
// the warp-level bounding-box pruning of assign.cu applies unchanged.
constexpr int kFaThreads = 256;

struct FaArgs {
  const float* pred;    // (A, 4) decoded predictions
  const float* gt;      // (G, 5)
  int A, G, C;
  float t1, t2_lo;      // thresh1, fp32(thresh1 + clamp_eps)
  float clamp_eps;
  uint32_t* rowmax;     // (G) IoU bits, zero-initialised
  float* out;           // (A, C), zero-initialised
};

__device__ __forceinline__ void warp_bbox(float4 b, bool ok, float& x0, float& y0, float& x1, float& y1) {
  float mnx = ok ? b.x : CUDART_INF_F, mny = ok ? b.y : CUDART_INF_F, mxx = ok ? b.z : -CUDART_INF_F, mxy = ok ? b.w : -CUDART_INF_F;
  x0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(mnx)));
  y0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(mny)));
  x1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(mxx)));
  y1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(mxy)));
}

template <bool SCATTER>
__global__ void __launch_bounds__(kFaThreads) fa_kernel(const FaArgs p) {
  extern __shared__ __align__(16) unsigned char raw[];
  float4* sbox = reinterpret_cast<float4*>(raw);
  float* sarea = reinterpret_cast<float*>(sbox + p.G);
  float* sden = sarea + p.G;                                   // SCATTER: t2[g] - t1
  int* slab = reinterpret_cast<int*>(sden + p.G);              // SCATTER: class index of the GT
  uint32_t* srmax = reinterpret_cast<uint32_t*>(slab + p.G);   // !SCATTER: (warps, G)
  __shared__ int s_fill;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_fill = 1;
  __syncthreads();
  for (int g = t; g < p.G; g += kFaThreads) {
    const float* r = p.gt + g * 5;
    const float4 bx = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
    sbox[g] = bx;
    sarea[g] = box_area(bx);
    if (SCATTER) {
      const float rm = __uint_as_float(p.rowmax[g]);
      const float t2 = fminf(fmaxf(rm, p.t2_lo), 1.f);         // free_anchor.py:60-64
      sden[g] = t2 - p.t1;
      slab[g] = (int)__ldg(r + 4) - 1;                         // :52
      // the reference's empty-set test (:71): is any probability above clamp_eps?  Row maxima decide it.
      const float pm = fminf(fmaxf(__fdiv_rn(rm - p.t1, t2 - p.t1), 0.f), 1.f);
      if (pm > p.clamp_eps) s_fill = 0;
    }
  }
  if (!SCATTER)
    for (int i = t; i < (kFaThreads / 32) * p.G; i += kFaThreads) srmax[i] = 0u;
  const long long a = (long long)blockIdx.x * kFaThreads + t;
  const bool ok = a < p.A;
  const float4 an = ok ? ldg4(p.pred + a * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float aa = box_area(an);
  float bx0, by0, bx1, by1;
  warp_bbox(an, ok, bx0, by0, bx1, by1);
  __syncthreads();
  const bool fill = SCATTER && s_fill != 0;
  float* orow = p.out + a * p.C;
  if (SCATTER && fill && a == 0 && p.G > 0) {
    const int c = slab[0];                                     // :73: gt_pred_prob[0, 0] = 0.001 is the first entry scattered
    if (c >= 0 && c < p.C) orow[c] = 0.001f;
  }
  for (int g0 = 0; g0 < p.G; g0 += 32) {
    const int gl = g0 + lane;
    bool live = false;
    if (gl < p.G) {
      const float4 b = sbox[gl];
      live = !(b.z <= bx0 || b.x >= bx1 || b.w <= by0 || b.y >= by1);
    }
    uint32_t m = __ballot_sync(0xffffffffu, live);
    while (m) {
      const int g = g0 + __ffs(m) - 1;
      m &= m - 1;
      const float v = iou_pair(sbox[g], sarea[g], an, aa);     // Boxes.iou(gt, pred_box), :57
      if (!SCATTER) {
        if (__any_sync(0xffffffffu, v > 0.f)) {
          const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(v));
          if (lane == 0) srmax[warp * p.G + g] = w;
        }
      } else if (ok && !(fill && g == 0 && a == 0)) {
        const float pr = fminf(fmaxf(__fdiv_rn(v - p.t1, sden[g]), 0.f), 1.f);   // :65-66
        const int c = slab[g];
        if (pr != 0.f && c >= 0 && c < p.C) orow[c] = pr;      // ascending g: the last GT of a class wins (:83)
      }
    }
  }
  if (!SCATTER) {
    __syncthreads();
    for (int g = t; g < p.G; g += kFaThreads) {
      uint32_t u = 0u;
#pragma unroll
      for (int w = 0; w < kFaThreads / 32; ++w) u = max(u, srmax[w * p.G + g]);
      if (u) atomicMax(&p.rowmax[g], u);
    }
  } else if (fill && a == 0) {
    orow[0] = 0.f;                                             // :85-86 "remove effect of setting gt_pred_prob"
  }
}

// bag scores and targets: thread per (g, j) of matched_idx (G, K)
__global__ void __launch_bounds__(256) fa_bag_kernel(const int* __restrict__ idx, int G, int K, const float* __restrict__ anchors,
                                                     const float* __restrict__ gt, const float* __restrict__ scores, int C,
                                                     Vec4 mean, Vec4 stdv, float* __restrict__ out_score,
                                                     float* __restrict__ out_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * K) return;
  const int g = i / K, a = idx[i];
  const float* r = gt + g * 5;
  const float4 gb = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
  const int c = (int)__ldg(r + 4) - 1;
  out_score[i] = (c >= 0 && c < C) ? __ldg(scores + (long long)a * C + c) : 0.f;   // F.gather(pred_scores[idx], 2, label)
  reinterpret_cast<float4*>(out_off)[i] = encode_box<false>(ldg4(anchors + (long long)a * 4), gb, mean, stdv);
}

// ---------------------------------------------------------------------------------------------------------------
// OTA cost.  Per (GT g, point a):
//   cost = (sum_c focal(logit[a, c], onehot_g[c]) + reg_weight * -log(max(iou_ltrb, eps))) + 1e6 * !in_box_and_center
// The class sum differs from the background sum S_a = sum_c focal(x_c, 0) in one term only:
//   sum_c focal(x_c, onehot) = (S_a - focal(x_k, 0)) + focal(x_k, 1), k = class of g
// (a few ulp from the reference's own summation order, which MegDNN does not specify either: tolerance-gated).
struct OtaArgs {
  const float* pts;      // (A, 2)
  const float* radius;   // (A) stride * center_sampling_radius of the point's level
  const float* gt;       // (G, 5)
  const float* logits;   // (A, C)
  const float* deltas;   // (A, 4) predicted ltrb
  float* bg;             // (A) S_a
  float* cost;           // (G, A)
  float* ious;           // (G, A)
  int A, G, C;
  float alpha, one_minus_alpha, gamma, reg_weight, eps;
  int gamma_is_2;
};

__device__ __forceinline__ float logsigmoid_f(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

// sigmoid_focal_loss.py:30-36 for a target of exactly 0 or 1 (the products with the zero side vanish exactly)
__device__ __forceinline__ float focal_term(const OtaArgs& p, float x, bool positive) {
  const float s = sigmoid_ref(x);
  float loss = positive ? -logsigmoid_f(x) : -logsigmoid_f(-x);
  if (p.gamma != 0.f) {
    const float base = positive ? 1.f - s : s;
    loss *= p.gamma_is_2 ? base * base : powf(base, p.gamma);
  }
  if (p.alpha >= 0.f) loss *= positive ? p.alpha : p.one_minus_alpha;
  return loss;
}

__global__ void __launch_bounds__(256) ota_bg_kernel(const OtaArgs p) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.A) return;
  const float* x = p.logits + (long long)a * p.C;
  float s = 0.f;
  for (int c = 0; c < p.C; ++c) s += focal_term(p, __ldg(x + c), false);   // sequential, like the oracle's sum
  p.bg[a] = s;
}

__global__ void __launch_bounds__(256) ota_cost_kernel(const OtaArgs p) {
  extern __shared__ __align__(16) float sgt[];  // (G, 5)
  for (int i = threadIdx.x; i < p.G * 5; i += blockDim.x) sgt[i] = __ldg(p.gt + i);
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.A) return;
  const float px = __ldg(p.pts + 2 * a), py = __ldg(p.pts + 2 * a + 1), rad = __ldg(p.radius + a);
  const float4 d = ldg4(p.deltas + (long long)a * 4);
  const float bgsum = p.bg[a];
  // prediction in "ltrb as a box" form: (-l, -t, r, b), iou_loss.py:19
  const float p0 = -d.x, p1 = -d.y, p2 = d.z, p3 = d.w;
  const float parea = fmaxf(p2 - p0, 0.f) * fmaxf(p3 - p1, 0.f);
  const float* x = p.logits + (long long)a * p.C;
  for (int g = 0; g < p.G; ++g) {
    const float gx1 = sgt[g * 5], gy1 = sgt[g * 5 + 1], gx2 = sgt[g * 5 + 2], gy2 = sgt[g * 5 + 3];
    const int cls = (int)sgt[g * 5 + 4] - 1;
    // PointCoder.encode, boxcoder.py:132-133
    const float l = px - gx1, t = py - gy1, r = gx2 - px, b = gy2 - py;
    bool inside = fminf(fminf(l, t), fminf(r, b)) > 0.01f;                   // ota.py:93
    const float cx = __fdiv_rn(gx1 + gx2, 2.f), cy = __fdiv_rn(gy1 + gy2, 2.f);   // :97
    const float c0 = fmaxf(cx - rad, gx1), c1 = fmaxf(cy - rad, gy1), c2 = fminf(cx + rad, gx2), c3 = fminf(cy + rad, gy2);
    inside = inside && fminf(fminf(px - c0, py - c1), fminf(c2 - px, c3 - py)) > 0.f;   // :101-112
    // get_ltrb_boxes_iou, iou_loss.py:19-43
    const float q0 = -l, q1 = -t, q2 = r, q3 = b;
    const float garea = fmaxf(q2 - q0, 0.f) * fmaxf(q3 - q1, 0.f);
    const float w = fmaxf(fminf(p2, q2) - fmaxf(p0, q0), 0.f), h = fmaxf(fminf(p3, q3) - fmaxf(p1, q1), 0.f);
    const float inter = w * h;
    const float uni = (parea + garea) - inter;
    const float iou = __fdiv_rn(inter, fmaxf(uni, p.eps));
    const float loss_delta = -logf(fmaxf(iou, p.eps));                        // iou_loss.py:96
    float loss_cls = bgsum;
    if (cls >= 0 && cls < p.C) {
      const float xc = __ldg(x + cls);
      loss_cls = (bgsum - focal_term(p, xc, false)) + focal_term(p, xc, true);
    }
    const float cost = (loss_cls + p.reg_weight * loss_delta) + 1e6f * (inside ? 0.f : 1.f);   // ota.py:152
    p.cost[(long long)g * p.A + a] = cost;
    p.ious[(long long)g * p.A + a] = iou;
  }
}

__global__ void __launch_bounds__(256) ota_collect_kernel(const int* __restrict__ matched, const float* __restrict__ pts,
                                                          const float* __restrict__ gt, const float* __restrict__ ious, int A, int G,
                                                          float* __restrict__ cls_t, float* __restrict__ box_t,
                                                          float* __restrict__ iou_t) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  const int m = matched[a];
  float c = 0.f, io = 0.f;
  float4 bt = make_float4(0.f, 0.f, 0.f, 0.f);
  if (m >= 0 && m < G) {                                                      // ota.py:161 fg_mask
    const float* r = gt + m * 5;
    const float px = __ldg(pts + 2 * a), py = __ldg(pts + 2 * a + 1);
    c = __ldg(r + 4);                                                          // :162
    bt = make_float4(px - __ldg(r), py - __ldg(r + 1), __ldg(r + 2) - px, __ldg(r + 3) - py);   // :165-168
    io = __ldg(ious + (long long)m * A + a);                                   // :171-175
  }
  cls_t[a] = c;
  reinterpret_cast<float4*>(box_t)[a] = bt;
  iou_t[a] = io;
}

// ---------------------------------------------------------------------------------------------------------------
// COCO records.  One CTA: exclusive scan of the per-image counts, then (image, row) -> record.
__global__ void __launch_bounds__(1024) coco_format_kernel(const float* __restrict__ dets, const int* __restrict__ counts, int B, int K,
                                                           const int* __restrict__ image_ids, const int* __restrict__ cat_ids, int C,
                                                           int* __restrict__ rec_image, double* __restrict__ rec_bbox,
                                                           double* __restrict__ rec_score, int* __restrict__ rec_cat,
                                                           int* __restrict__ total) {
  extern __shared__ int soff[];  // (B + 1)
  const int t = threadIdx.x;
  if (t == 0) {
    int s = 0;
    for (int b = 0; b < B; ++b) {
      soff[b] = s;
      s += min(max(counts[b], 0), K);
    }
    soff[B] = s;
    *total = s;
  }
  __syncthreads();
  for (int i = t; i < B * K; i += blockDim.x) {
    const int b = i / K, j = i - b * K;
    if (j >= min(max(counts[b], 0), K)) continue;
    const float* d = dets + (long long)i * 6;
    const int o = soff[b] + j;
    const double x1 = (double)d[0], y1 = (double)d[1];
    rec_image[o] = image_ids[b];
    rec_bbox[4 * o] = x1;                       // coco_eval.py:125: boxes[:, 2:4] -= boxes[:, 0:2] on float64 rows
    rec_bbox[4 * o + 1] = y1;
    rec_bbox[4 * o + 2] = (double)d[2] - x1;
    rec_bbox[4 * o + 3] = (double)d[3] - y1;
    rec_score[o] = (double)d[4];
    const int lab = (int)d[5];
    rec_cat[o] = cat_ids ? ((lab >= 0 && lab < C) ? cat_ids[lab] : -1) : lab + 1;   // :131-136
  }
}

}  // namespace bdet

using namespace bdet;

extern "C" size_t bdet_free_anchor_box_prob_workspace(int G) { return align_up((size_t)(G > 0 ? G : 1) * 4, 256); }

extern "C" int bdet_free_anchor_box_prob(const float* pred_boxes, int A, const float* gt, int G, int num_classes,
                                         float box_iou_thresh, float thresh2_lower, float clamp_eps, float* box_prob,
                                         void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(A >= 0 && G >= 0 && num_classes >= 1, "bad sizes");
  BDET_REQUIRE(box_iou_thresh >= 0.f, "box_iou_thresh must be >= 0 (non-overlapping pairs are pruned as probability 0)");
  if (A == 0) return BDET_OK;
  BDET_REQUIRE(pred_boxes && box_prob && (G == 0 || gt), "null argument");
  BDET_REQUIRE(aligned16(pred_boxes), "pred_boxes must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  BDET_CUDA(cudaMemsetAsync(box_prob, 0, (size_t)A * num_classes * 4, st));
  if (G == 0) return BDET_OK;
  if (!workspace || workspace_bytes < bdet_free_anchor_box_prob_workspace(G))
    return set_error(BDET_EWORKSPACE, "bdet_free_anchor_box_prob: workspace needs %zu bytes", bdet_free_anchor_box_prob_workspace(G));
  FaArgs p;
  p.pred = pred_boxes;
  p.gt = gt;
  p.A = A;
  p.G = G;
  p.C = num_classes;
  p.t1 = box_iou_thresh;
  p.t2_lo = thresh2_lower;  // fp32(thresh1 + clamp_eps) with the sum taken in Python floats (free_anchor.py:62)
  p.clamp_eps = clamp_eps;
  p.rowmax = reinterpret_cast<uint32_t*>(workspace);
  p.out = box_prob;
  BDET_CUDA(cudaMemsetAsync(p.rowmax, 0, (size_t)G * 4, st));
  const size_t smem = (size_t)G * (16 + 4 + 4 + 4 + 4 * (kFaThreads / 32));
  if (smem > 200 * 1024) return set_error(BDET_EUNSUPPORTED, "bdet_free_anchor_box_prob: too many GT boxes for shared memory");
  if (smem > 40 * 1024) {
    BDET_CUDA(cudaFuncSetAttribute(fa_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BDET_CUDA(cudaFuncSetAttribute(fa_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int grid = ceil_div(A, kFaThreads);
  BDET_KERNEL("fa_rowmax_kernel", st, fa_kernel<false><<<grid, kFaThreads, smem, st>>>(p));
  BDET_KERNEL("fa_scatter_kernel", st, fa_kernel<true><<<grid, kFaThreads, smem, st>>>(p));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_free_anchor_bags(const int* matched_idx, int G, int K, const float* anchors, const float* gt,
                                     const float* pred_scores, int num_classes, const float* mean_host, const float* std_host,
                                     float* matched_score, float* matched_offsets, bdet_stream_t stream) {
  BDET_REQUIRE(G >= 0 && K >= 0 && num_classes >= 1, "bad sizes");
  if (G == 0 || K == 0) return BDET_OK;
  BDET_REQUIRE(matched_idx && anchors && gt && pred_scores && matched_score && matched_offsets, "null argument");
  BDET_REQUIRE(aligned16(anchors) && aligned16(matched_offsets), "anchors / matched_offsets must be 16-byte aligned");
  Vec4 mean, stdv;
  for (int i = 0; i < 4; ++i) {
    mean.v[i] = mean_host ? mean_host[i] : 0.f;
    stdv.v[i] = std_host ? std_host[i] : 1.f;
  }
  cudaStream_t st = as_stream(stream);
  BDET_KERNEL("fa_bag_kernel", st,
              fa_bag_kernel<<<ceil_div((int64_t)G * K, 256), 256, 0, st>>>(matched_idx, G, K, anchors, gt, pred_scores, num_classes,
                                                                          mean, stdv, matched_score, matched_offsets));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" size_t bdet_ota_cost_workspace(int A) { return align_up((size_t)(A > 0 ? A : 1) * 4, 256); }

extern "C" int bdet_ota_cost(const float* points, const float* radius, int A, const float* gt, int G, const float* cls_logits,
                             int num_classes, const float* pred_deltas, double alpha, double gamma, double reg_weight, float* cost,
                             float* ious, void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(A >= 0 && G >= 0 && num_classes >= 1, "bad sizes");
  if (A == 0 || G == 0) return BDET_OK;
  BDET_REQUIRE(points && radius && gt && cls_logits && pred_deltas && cost && ious, "null argument");
  BDET_REQUIRE(aligned16(pred_deltas), "pred_deltas must be 16-byte aligned");
  if (!workspace || workspace_bytes < bdet_ota_cost_workspace(A))
    return set_error(BDET_EWORKSPACE, "bdet_ota_cost: workspace needs %zu bytes", bdet_ota_cost_workspace(A));
  if ((size_t)G * 20 > 200 * 1024) return set_error(BDET_EUNSUPPORTED, "bdet_ota_cost: too many GT boxes for shared memory");
  OtaArgs p;
  p.pts = points;
  p.radius = radius;
  p.gt = gt;
  p.logits = cls_logits;
  p.deltas = pred_deltas;
  p.bg = reinterpret_cast<float*>(workspace);
  p.cost = cost;
  p.ious = ious;
  p.A = A;
  p.G = G;
  p.C = num_classes;
  p.alpha = (float)alpha;
  p.one_minus_alpha = (float)(1.0 - alpha);  // `(1 - alpha)` is a Python float in the reference
  p.gamma = (float)gamma;
  p.gamma_is_2 = gamma == 2.0;
  p.reg_weight = (float)reg_weight;
  p.eps = 1.1920928955078125e-07f;  // np.finfo(np.float32).eps, ota.py:143
  cudaStream_t st = as_stream(stream);
  const size_t smem = (size_t)G * 20;
  if (smem > 40 * 1024) BDET_CUDA(cudaFuncSetAttribute(ota_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BDET_KERNEL("ota_bg_kernel", st, ota_bg_kernel<<<ceil_div(A, 256), 256, 0, st>>>(p));
  BDET_KERNEL("ota_cost_kernel", st, ota_cost_kernel<<<ceil_div(A, 256), 256, smem, st>>>(p));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_ota_collect(const int* matched_gt, const float* points, int A, const float* gt, int G, const float* ious,
                                float* gt_classes, float* box_targets, float* iou_targets, bdet_stream_t stream) {
  BDET_REQUIRE(A >= 0 && G >= 0, "bad sizes");
  if (A == 0) return BDET_OK;
  BDET_REQUIRE(matched_gt && points && gt_classes && box_targets && iou_targets && (G == 0 || (gt && ious)), "null argument");
  BDET_REQUIRE(aligned16(box_targets), "box_targets must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  BDET_KERNEL("ota_collect_kernel", st,
              ota_collect_kernel<<<ceil_div(A, 256), 256, 0, st>>>(matched_gt, points, gt, ious, A, G, gt_classes, box_targets, iou_targets));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_coco_format(const float* dets, const int* counts, int B, int K, const int* image_ids,
                                const int* category_ids, int num_classes, int* rec_image_id, double* rec_bbox_xywh,
                                double* rec_score, int* rec_category_id, int* total, bdet_stream_t stream) {
  BDET_REQUIRE(B >= 0 && K >= 0, "bad sizes");
  BDET_REQUIRE(total, "null total");
  cudaStream_t st = as_stream(stream);
  if (B == 0 || K == 0) {
    BDET_CUDA(cudaMemsetAsync(total, 0, 4, st));
    return BDET_OK;
  }
  BDET_REQUIRE(dets && counts && image_ids && rec_image_id && rec_bbox_xywh && rec_score && rec_category_id, "null argument");
  const size_t smem = (size_t)(B + 1) * 4;
  if (smem > 200 * 1024) return set_error(BDET_EUNSUPPORTED, "bdet_coco_format: batch too large for one call");
  if (smem > 40 * 1024) BDET_CUDA(cudaFuncSetAttribute(coco_format_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BDET_KERNEL("coco_format_kernel", st,
              coco_format_kernel<<<1, 1024, smem, st>>>(dets, counts, B, K, image_ids, category_ids, num_classes, rec_image_id,
                                                        rec_bbox_xywh, rec_score, rec_category_id, total));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
