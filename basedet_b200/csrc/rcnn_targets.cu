// the label glue of the second stage, RCNN.get_ground_truth (training branch).
// Reference: basedet/layers/head/rcnn.py:95-147.  Per image the reference concatenates proposals and GT boxes,
// computes the (R, G) IoU matrix, max / argmax over G, fg / bg masks, subsamples both with sample_labels
// (layers/common/sampling.py:7-30) and gathers rois / labels / BoxCoder targets of the survivors.
//   rcnn_match_kernel   : one thread per row of all_rois: IoU against the image's GT held in shared memory, running
//                         (max, first argmax), class of the matched GT, fg / bg flags (rcnn.py:108-123); no (R, G) matrix.
//   bdet_sample_labels  : (topk.cu) the two subsampling steps on the flag arrays, budgets on the device (:125-128).
//   rcnn_collect_kernel : ordered compaction of the kept rows (every 256-row CTA counts the kept rows before it),
//                         labels[bg] = 0, BoxCoder.encode against the matched GT (:130-137).
#include "common.cuh"

namespace bdet {

struct RcnnArgs {
  const float* rois;     // (B, Rmax, 5) rows [batch, x1, y1, x2, y2], image b's proposals first n_rois[b] rows
  const int* n_rois;     // (B)
  const float* gt;       // (B, Gmax, 5)
  const int* num_gt;     // (B)
  int B, Rmax, Gmax, N;  // N = Rmax + Gmax rows per image in the work arrays
  float fg_thr, bg_lo, bg_hi;
  float* all_rois;       // (B, N, 5)
  int* n_all;            // (B)
  int* assign;           // (B, N)
  float* cls;            // (B, N) class of the matched GT (fp32 as in the reference)
  int* fg;               // (B, N) 0 / 1
  int* bg;               // (B, N) 0 / 1
  // collect
  int num_out;
  Vec4 mean, stdv;
  float* out_rois;       // (B, num_out, 5)
  int* out_labels;       // (B, num_out)
  float* out_targets;    // (B, num_out, 4)
  int* out_count;        // (B)
};

constexpr int kRcnnThreads = 128;

__global__ void __launch_bounds__(kRcnnThreads) rcnn_match_kernel(const RcnnArgs p) {
  extern __shared__ __align__(16) unsigned char raw[];
  float4* sbox = reinterpret_cast<float4*>(raw);            // Gmax
  float* sarea = reinterpret_cast<float*>(sbox + p.Gmax);   // Gmax
  float* scls = sarea + p.Gmax;                              // Gmax
  const int b = blockIdx.y, t = threadIdx.x;
  const int G = max(0, min(p.num_gt[b], p.Gmax));
  const int nr = max(0, min(p.n_rois[b], p.Rmax));
  const float* gtb = p.gt + (long long)b * p.Gmax * 5;
  for (int g = t; g < G; g += kRcnnThreads) {
    const float* r = gtb + g * 5;
    const float4 bx = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
    sbox[g] = bx;
    sarea[g] = box_area(bx);
    scls[g] = __ldg(r + 4);
  }
  __syncthreads();
  if (blockIdx.x == 0 && t == 0) p.n_all[b] = nr + G;
  const int i = blockIdx.x * kRcnnThreads + t;
  if (i >= p.N) return;
  const long long o = (long long)b * p.N + i;
  float row[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  int fg = 0, bg = 0, bi = 0;
  float c = 0.f;
  if (i < nr + G) {
    if (i < nr) {
      const float* r = p.rois + ((long long)b * p.Rmax + i) * 5;
#pragma unroll
      for (int j = 0; j < 5; ++j) row[j] = __ldg(r + j);
    } else {  // gt_rois = [bid, gt box], rcnn.py:108-109
      const float4 q = sbox[i - nr];
      row[0] = (float)b;
      row[1] = q.x;
      row[2] = q.y;
      row[3] = q.z;
      row[4] = q.w;
    }
    if (G > 0) {
      const float4 bx = make_float4(row[1], row[2], row[3], row[4]);
      const float ba = box_area(bx);
      float best = -1.f;  // every IoU is >= 0: the first GT always takes the lead, ties keep the first index (:115-116)
      for (int g = 0; g < G; ++g) {
        const float v = iou_pair(bx, ba, sbox[g], sarea[g]);
        if (v > best) {
          best = v;
          bi = g;
        }
      }
      c = scls[bi];                                              // :117
      fg = (best >= p.fg_thr) && (c >= 0.f);                     // :119
      bg = (best >= p.bg_lo) && (best < p.bg_hi);                // :120-123
    }
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) p.all_rois[o * 5 + j] = row[j];
  p.assign[o] = bi;
  p.cls[o] = c;
  p.fg[o] = fg;
  p.bg[o] = bg;
}

constexpr int kCollectThreads = 256;

// grid (ceil(N / 256), B): a CTA owns 256 consecutive rows of all_rois; the output slot of its first row is the number
// of kept rows before it, which every CTA counts for itself from the flag arrays (<= a few thousand flags) -- no
// per-image serial compaction.
__global__ void __launch_bounds__(kCollectThreads) rcnn_collect_kernel(const RcnnArgs p) {
  __shared__ int warp_cnt[kCollectThreads / 32];
  __shared__ int sred[kCollectThreads / 32];
  const int b = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n = min(p.n_all[b], p.N);
  const int i0 = blockIdx.x * kCollectThreads;
  const long long rowb = (long long)b * p.N;
  // kept rows before this CTA's first row (and, in the last CTA of the image, the image's total)
  int before = 0;
  for (int i = t; i < min(i0, n); i += kCollectThreads) before += (p.fg[rowb + i] | p.bg[rowb + i]) != 0;
  before = __reduce_add_sync(0xffffffffu, before);
  if (lane == 0) sred[warp] = before;
  const int i = i0 + t;
  const long long o = rowb + i;
  int fg = 0, bg = 0;
  if (i < n) {
    fg = p.fg[o];
    bg = p.bg[o];
  }
  const bool keep = (fg | bg) != 0;                             // rcnn.py:132
  const uint32_t bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  int base = 0, inblock = 0, mine = 0;
  for (int w = 0; w < kCollectThreads / 32; ++w) {
    base += sred[w];
    if (w < warp) mine += warp_cnt[w];
    inblock += warp_cnt[w];
  }
  if (t == 0 && i0 < n && i0 + kCollectThreads >= n) p.out_count[b] = min(base + inblock, p.num_out);
  if (t == 0 && n == 0 && blockIdx.x == 0) p.out_count[b] = 0;
  const int pos = base + mine + __popc(bal & ((1u << lane) - 1u));
  if (!keep || pos >= p.num_out) return;
  const float* gtb = p.gt + (long long)b * p.Gmax * 5;
  const long long d = (long long)b * p.num_out + pos;
  const float* r = p.all_rois + o * 5;
#pragma unroll
  for (int j = 0; j < 5; ++j) p.out_rois[d * 5 + j] = r[j];      // :134
  p.out_labels[d] = bg ? 0 : (int)p.cls[o];                      // :130, :133
  const float* gr = gtb + (long long)p.assign[o] * 5;            // :135
  const float4 tg = encode_box<false>(make_float4(r[1], r[2], r[3], r[4]),
                                      make_float4(__ldg(gr), __ldg(gr + 1), __ldg(gr + 2), __ldg(gr + 3)), p.mean, p.stdv);
  reinterpret_cast<float4*>(p.out_targets)[d] = tg;              // :136-137
}

}  // namespace bdet

using namespace bdet;

extern "C" int bdet_rcnn_match(const float* rois, const int* n_rois_dev, int Rmax, const float* gt, const int* num_gt_dev,
                               int Gmax, int B, float fg_thresh, float bg_thresh_low, float bg_thresh_high, float* all_rois,
                               int* n_all, int* assign, float* matched_class, int* fg_mask, int* bg_mask,
                               bdet_stream_t stream) {
  BDET_REQUIRE(Rmax >= 0 && Gmax >= 0 && B >= 0, "negative size");
  if (B == 0) return BDET_OK;
  BDET_REQUIRE(n_rois_dev && num_gt_dev && n_all, "null argument");
  const int N = Rmax + Gmax;
  cudaStream_t st = as_stream(stream);
  if (N == 0) {
    BDET_CUDA(cudaMemsetAsync(n_all, 0, (size_t)B * 4, st));
    return BDET_OK;
  }
  BDET_REQUIRE((Rmax == 0 || rois) && (Gmax == 0 || gt) && all_rois && assign && matched_class && fg_mask && bg_mask, "null argument");
  BDET_REQUIRE(B <= 65535, "B > 65535");
  RcnnArgs a = {};
  a.rois = rois;
  a.n_rois = n_rois_dev;
  a.gt = gt;
  a.num_gt = num_gt_dev;
  a.B = B;
  a.Rmax = Rmax;
  a.Gmax = Gmax;
  a.N = N;
  a.fg_thr = fg_thresh;
  a.bg_lo = bg_thresh_low;
  a.bg_hi = bg_thresh_high;
  a.all_rois = all_rois;
  a.n_all = n_all;
  a.assign = assign;
  a.cls = matched_class;
  a.fg = fg_mask;
  a.bg = bg_mask;
  const size_t smem = (size_t)Gmax * 24;
  if (smem > 40 * 1024) BDET_CUDA(cudaFuncSetAttribute(rcnn_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BDET_KERNEL("rcnn_match_kernel", st, rcnn_match_kernel<<<dim3(ceil_div(N, kRcnnThreads), B), kRcnnThreads, smem, st>>>(a));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_rcnn_collect(const float* all_rois, const int* n_all, const int* assign, const float* matched_class,
                                 const int* fg_mask, const int* bg_mask, int N, const float* gt, int Gmax, int B,
                                 const float* mean_host, const float* std_host, int num_out, float* out_rois,
                                 int* out_labels, float* out_targets, int* out_count, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0 && Gmax >= 0 && B >= 0 && num_out >= 0, "negative size");
  if (B == 0) return BDET_OK;
  BDET_REQUIRE(out_count && n_all, "null argument");
  cudaStream_t st = as_stream(stream);
  if (N == 0 || num_out == 0) {
    BDET_CUDA(cudaMemsetAsync(out_count, 0, (size_t)B * 4, st));
    return BDET_OK;
  }
  BDET_REQUIRE(all_rois && assign && matched_class && fg_mask && bg_mask && gt && out_rois && out_labels && out_targets,
               "null argument");
  BDET_REQUIRE(aligned16(out_targets), "out_targets must be 16-byte aligned");
  RcnnArgs a = {};
  a.all_rois = const_cast<float*>(all_rois);
  a.n_all = const_cast<int*>(n_all);
  a.assign = const_cast<int*>(assign);
  a.cls = const_cast<float*>(matched_class);
  a.fg = const_cast<int*>(fg_mask);
  a.bg = const_cast<int*>(bg_mask);
  a.gt = gt;
  a.B = B;
  a.Gmax = Gmax;
  a.N = N;
  a.num_out = num_out;
  for (int i = 0; i < 4; ++i) {
    a.mean.v[i] = mean_host ? mean_host[i] : 0.f;
    a.stdv.v[i] = std_host ? std_host[i] : 1.f;
  }
  a.out_rois = out_rois;
  a.out_labels = out_labels;
  a.out_targets = out_targets;
  a.out_count = out_count;
  BDET_CUDA(cudaMemsetAsync(out_rois, 0, (size_t)B * num_out * 5 * 4, st));
  BDET_CUDA(cudaMemsetAsync(out_labels, 0, (size_t)B * num_out * 4, st));
  BDET_CUDA(cudaMemsetAsync(out_targets, 0, (size_t)B * num_out * 16, st));
  BDET_REQUIRE(B <= 65535, "B > 65535");
  BDET_KERNEL("rcnn_collect_kernel", st, rcnn_collect_kernel<<<dim3(ceil_div(N, kCollectThreads), B), kCollectThreads, 0, st>>>(a));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
