// Fused target assignment: IoU (G,A) -> Matcher -> labels[fg]=class -> BoxCoder.encode for B images,
// without materialising the (G,A) matrix.
// Reference: basedet/models/det/retinanet.py:211-232 (get_ground_truth), which composes
//   structures/op_patch.py:33-97 (box_iou), layers/common/matcher.py:31-51 (Matcher),
//   structures/boxcoder.py:61-73 (BoxCoder.encode); RPN uses the same sequence (models/det/rpn.py:215-226).
//
// Data flow per CTA (kAPT*256 consecutive anchors of one image):
//   anchors: one 128-bit load each, kept in registers; GT boxes (+area, class) staged in shared memory;
//   the CTA reduces the bounding box of its anchors and keeps only the GTs that can intersect it
//   (ascending order, so "first argmax" survives); every other (g, anchor) pair has IoU exactly +0 and
//   can never win a strict '>' against the running maximum that starts at (0, index 0).
//   Row maxima (needed by allow_low_quality_matches) are folded per warp with redux.sync and kept in
//   shared memory; the CTA publishes only its non-zero row maxima.
//   Kernel 2 re-evaluates a (g, CTA) segment only where the CTA's maximum equals the global row maximum.
// HBM traffic: 16 B/anchor read + 24 B/anchor written (labels, indices, offsets) per image.
#include "common.cuh"

namespace bdet {

constexpr int kAT = 256;

struct AssignArgs {
  const float* anchors;
  const float* gt;  // (B, Gmax, 5)
  const int* num_gt;
  int* labels;
  int* idx;
  float* offsets;
  uint32_t* rowmax;  // (B, Gmax) fp32 bits (IoU >= 0, so uint order == float order); zero-initialised
  int* blk_count;    // (B, tiles)
  uint2* blk_list;   // (B, tiles, Gmax): (g, local max bits)
  int A, Gmax, tiles, allow_lq, apply_class;
  MatchCfg cfg;
  Vec4 mean, stdv;
};

struct AssignSmem {
  float4* box;
  float* area;
  float* cls;
  uint32_t* rmax;
  int* list;
};
__device__ __forceinline__ AssignSmem carve(unsigned char* raw, int Gmax) {
  AssignSmem s;
  s.box = reinterpret_cast<float4*>(raw);
  s.area = reinterpret_cast<float*>(s.box + Gmax);
  s.cls = s.area + Gmax;
  s.rmax = reinterpret_cast<uint32_t*>(s.cls + Gmax);
  s.list = reinterpret_cast<int*>(s.rmax + Gmax);
  return s;
}

template <int APT>
__global__ void __launch_bounds__(kAT) assign_main_kernel(const AssignArgs p) {
  extern __shared__ __align__(16) unsigned char raw[];
  AssignSmem s = carve(raw, p.Gmax);
  __shared__ uint32_t sred[4];
  __shared__ float sbb[4];
  __shared__ int scount, spub;
  const int b = blockIdx.y, tile = blockIdx.x, t = threadIdx.x, lane = t & 31;
  const int G = min(p.num_gt[b], p.Gmax);

  if (t < 4) sred[t] = (t < 2) ? 0xffffffffu : 0u;
  if (t == 0) spub = 0;
  const float* gt = p.gt + (long long)b * p.Gmax * 5;
  for (int g = t; g < G; g += kAT) {
    const float* r = gt + g * 5;
    float4 bx = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
    s.box[g] = bx;
    s.area[g] = box_area(bx);
    s.cls[g] = __ldg(r + 4);
    s.rmax[g] = 0u;
  }
  float4 an[APT];
  float aa[APT];
  bool ok[APT];
  float mnx = CUDART_INF_F, mny = CUDART_INF_F, mxx = -CUDART_INF_F, mxy = -CUDART_INF_F;
  const long long col0 = (long long)tile * (kAT * APT) + t;
#pragma unroll
  for (int j = 0; j < APT; ++j) {
    long long c = col0 + j * kAT;
    ok[j] = c < p.A;
    an[j] = ok[j] ? ldg4(p.anchors + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    aa[j] = box_area(an[j]);
    if (ok[j]) {
      mnx = fminf(mnx, an[j].x);
      mny = fminf(mny, an[j].y);
      mxx = fmaxf(mxx, an[j].z);
      mxy = fmaxf(mxy, an[j].w);
    }
  }
  __syncthreads();
  {
    uint32_t r0 = __reduce_min_sync(0xffffffffu, f2ord(mnx));
    uint32_t r1 = __reduce_min_sync(0xffffffffu, f2ord(mny));
    uint32_t r2 = __reduce_max_sync(0xffffffffu, f2ord(mxx));
    uint32_t r3 = __reduce_max_sync(0xffffffffu, f2ord(mxy));
    if (lane == 0) {
      atomicMin(&sred[0], r0);
      atomicMin(&sred[1], r1);
      atomicMax(&sred[2], r2);
      atomicMax(&sred[3], r3);
    }
  }
  __syncthreads();
  if (t < 4) sbb[t] = ord2f(sred[t]);
  __syncthreads();
  // Order-preserving compaction of the GTs that can overlap this CTA's anchors (warp 0).
  if (t < 32) {
    const float bb0 = sbb[0], bb1 = sbb[1], bb2 = sbb[2], bb3 = sbb[3];
    int n = 0;
    for (int g0 = 0; g0 < G; g0 += 32) {
      int g = g0 + lane;
      bool live = false;
      if (g < G) {
        float4 a = s.box[g];
        live = !(a.z <= bb0 || a.x >= bb2 || a.w <= bb1 || a.y >= bb3);  // NaN GT -> live
      }
      uint32_t mask = __ballot_sync(0xffffffffu, live);
      if (live) s.list[n + __popc(mask & ((1u << lane) - 1u))] = g;
      n += __popc(mask);
    }
    if (lane == 0) scount = n;
  }
  __syncthreads();
  const int n_live = scount;

  // Running (max, first argmax) over G.  All IoUs are >= +0 and never NaN, so the state after the
  // (possibly skipped) row 0 is at least (0, 0); skipped rows are exact zeros and cannot win '>'.
  float best[APT];
  int bidx[APT];
#pragma unroll
  for (int j = 0; j < APT; ++j) {
    best[j] = 0.f;
    bidx[j] = 0;
  }
  for (int i = 0; i < n_live; ++i) {
    const int g = s.list[i];
    const float4 a = s.box[g];
    const float ga = s.area[g];
    float rv = 0.f;
#pragma unroll
    for (int j = 0; j < APT; ++j) {
      float v = ok[j] ? iou_pair(a, ga, an[j], aa[j]) : 0.f;
      if (v > best[j]) {
        best[j] = v;
        bidx[j] = g;
      }
      rv = fmaxf(rv, v);
    }
    if (p.allow_lq && __any_sync(0xffffffffu, rv > 0.f)) {
      uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(rv));
      if (lane == 0) atomicMax(&s.rmax[g], w);
    }
  }

#pragma unroll
  for (int j = 0; j < APT; ++j) {
    if (!ok[j]) continue;
    long long o = (long long)b * p.A + col0 + j * kAT;
    int label = threshold_label(p.cfg, best[j]);
    float4 off = make_float4(0.f, 0.f, 0.f, 0.f);
    if (G > 0) {
      if (p.apply_class && label == 1) label = (int)s.cls[bidx[j]];  // retinanet.py:222-223 astype("int32")
      off = encode_box(an[j], s.box[bidx[j]], p.mean, p.stdv);
    }
    p.labels[o] = label;
    p.idx[o] = bidx[j];
    reinterpret_cast<float4*>(p.offsets)[o] = off;
  }

  if (p.allow_lq) {
    __syncthreads();
    uint2* lst = p.blk_list + ((long long)b * p.tiles + tile) * p.Gmax;
    for (int g = t; g < G; g += kAT) {
      uint32_t u = s.rmax[g];
      if (u) {
        atomicMax(&p.rowmax[(long long)b * p.Gmax + g], u);
        lst[atomicAdd(&spub, 1)] = make_uint2((uint32_t)g, u);
      }
    }
    __syncthreads();
    if (t == 0) p.blk_count[(long long)b * p.tiles + tile] = spub;
  }
}

// allow_low_quality_matches, matcher.py:47-49: every anchor whose IoU with g equals the row maximum of g
// gets label 1 (then the class of ITS OWN matched GT, retinanet.py:222-223) -- including rows whose
// maximum is 0, where every zero-IoU anchor qualifies.
template <int APT>
__global__ void __launch_bounds__(kAT) assign_lq_kernel(const AssignArgs p) {
  extern __shared__ __align__(16) unsigned char raw[];
  int* hit = reinterpret_cast<int*>(raw);  // Gmax
  __shared__ int nhit;
  const int b = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const int G = min(p.num_gt[b], p.Gmax);
  if (t == 0) nhit = 0;
  __syncthreads();
  const uint32_t* rm = p.rowmax + (long long)b * p.Gmax;
  const uint2* lst = p.blk_list + ((long long)b * p.tiles + tile) * p.Gmax;
  const int n = p.blk_count[(long long)b * p.tiles + tile];
  for (int i = t; i < n; i += kAT) {
    uint2 e = lst[i];
    if (e.y == rm[e.x]) hit[atomicAdd(&nhit, 1)] = (int)e.x;
  }
  for (int g = t; g < G; g += kAT)
    if (rm[g] == 0u) hit[atomicAdd(&nhit, 1)] = g;  // zero-maximum rows never appear in a CTA list
  __syncthreads();
  const int nh = nhit;
  if (nh == 0) return;
  const float* gt = p.gt + (long long)b * p.Gmax * 5;
  const long long col0 = (long long)tile * (kAT * APT) + t;
#pragma unroll
  for (int j = 0; j < APT; ++j) {
    long long c = col0 + j * kAT;
    if (c >= p.A) continue;
    float4 an = ldg4(p.anchors + c * 4);
    float aa = box_area(an);
    bool lq = false;
    for (int i = 0; i < nh; ++i) {
      const float* r = gt + hit[i] * 5;
      float4 a = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
      float v = iou_pair(a, box_area(a), an, aa);
      lq |= (__float_as_uint(v) == rm[hit[i]]);
    }
    if (lq) {
      long long o = (long long)b * p.A + c;
      p.labels[o] = p.apply_class ? (int)__ldg(gt + p.idx[o] * 5 + 4) : 1;
    }
  }
}

static int assign_apt(int A, int B) {
  return ((long long)ceil_div(A, kAT * 2) * B >= (long long)sm_count() * 4) ? 2 : 1;
}

}  // namespace bdet

using namespace bdet;

extern "C" size_t bdet_assign_targets_workspace(int Gmax, int A, int B) {
  if (Gmax <= 0 || A <= 0 || B <= 0) return 16;
  int tiles = ceil_div(A, kAT * assign_apt(A, B));
  return align_up((size_t)B * Gmax * 4, 256) + align_up((size_t)B * tiles * 4, 256) + (size_t)B * tiles * Gmax * 8 + 256;
}

extern "C" int bdet_assign_targets(const float* anchors, int A, const float* gt, int Gmax, const int* num_gt_dev, int B,
                                   const float* thresholds_host, const int* labels_host, int n_labels,
                                   int allow_low_quality, int apply_class, const float* mean_host,
                                   const float* std_host, int* labels, int* match_idx, float* offsets,
                                   void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(A >= 0 && Gmax >= 0 && B >= 0, "negative size");
  AssignArgs a;
  int rc = make_match_cfg(&a.cfg, thresholds_host, labels_host, n_labels);
  if (rc) return rc;
  if (A == 0 || B == 0) return BDET_OK;
  BDET_REQUIRE(anchors && labels && match_idx && offsets && num_gt_dev, "null argument");
  BDET_REQUIRE(Gmax == 0 || gt, "null gt");
  BDET_REQUIRE(aligned16(anchors) && aligned16(offsets), "anchors/offsets must be 16-byte aligned");
  BDET_REQUIRE(B <= 65535, "B > 65535");
  const size_t smem = (size_t)max(Gmax, 1) * 32;
  if (smem > 200 * 1024) return set_error(BDET_EUNSUPPORTED, "bdet_assign_targets: Gmax > 6400 does not fit shared memory");
  const int apt = assign_apt(A, B);
  a.anchors = anchors;
  a.gt = gt;
  a.num_gt = num_gt_dev;
  a.labels = labels;
  a.idx = match_idx;
  a.offsets = offsets;
  a.A = A;
  a.Gmax = Gmax;
  a.tiles = ceil_div(A, kAT * apt);
  a.allow_lq = (allow_low_quality != 0 && Gmax > 0) ? 1 : 0;
  a.apply_class = apply_class != 0;
  for (int i = 0; i < 4; ++i) {
    a.mean.v[i] = mean_host ? mean_host[i] : 0.f;
    a.stdv.v[i] = std_host ? std_host[i] : 1.f;
  }
  a.rowmax = nullptr;
  a.blk_count = nullptr;
  a.blk_list = nullptr;
  cudaStream_t st = as_stream(stream);
  if (a.allow_lq) {
    const size_t need = bdet_assign_targets_workspace(Gmax, A, B);
    if (!workspace || workspace_bytes < need)
      return set_error(BDET_EWORKSPACE, "bdet_assign_targets: workspace needs %zu bytes", need);
    BDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "workspace must be 8-byte aligned");
    char* w = reinterpret_cast<char*>(workspace);
    a.rowmax = reinterpret_cast<uint32_t*>(w);
    w += align_up((size_t)B * Gmax * 4, 256);
    a.blk_count = reinterpret_cast<int*>(w);
    w += align_up((size_t)B * a.tiles * 4, 256);
    a.blk_list = reinterpret_cast<uint2*>(w);
    BDET_CUDA(cudaMemsetAsync(a.rowmax, 0, (size_t)B * Gmax * 4, st));
  }
  dim3 grid(a.tiles, B);
  if (apt == 2) {
    if (smem > 40 * 1024) {
      BDET_CUDA(cudaFuncSetAttribute(assign_main_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    BDET_KERNEL("assign_main_kernel", st, assign_main_kernel<2><<<grid, kAT, smem, st>>>(a));
    if (a.allow_lq) BDET_KERNEL("assign_lq_kernel", st, assign_lq_kernel<2><<<grid, kAT, (size_t)max(Gmax, 1) * 4, st>>>(a));
  } else {
    if (smem > 40 * 1024) {
      BDET_CUDA(cudaFuncSetAttribute(assign_main_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    BDET_KERNEL("assign_main_kernel", st, assign_main_kernel<1><<<grid, kAT, smem, st>>>(a));
    if (a.allow_lq) BDET_KERNEL("assign_lq_kernel", st, assign_lq_kernel<1><<<grid, kAT, (size_t)max(Gmax, 1) * 4, st>>>(a));
  }
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
