// Fused target assignment: IoU (G,A) -> Matcher -> labels[fg]=class -> BoxCoder.encode for B images,
// without materialising the (G,A) matrix.
// Reference: basedet/models/det/retinanet.py:211-232 (get_ground_truth), which composes
//   structures/op_patch.py:33-97 (box_iou), layers/common/matcher.py:31-51 (Matcher),
//   structures/boxcoder.py:61-73 (BoxCoder.encode); RPN uses the same sequence (models/det/rpn.py:215-226).
//
// assign_main_kernel -- one CTA = 8 warps x 64 consecutive anchors of one image:
//   * anchors: one 128-bit load each, kept in registers; GT boxes (+area, class) staged in shared memory;
//   * every WARP reduces the bounding box of its 64 anchors (4 redux.sync) and walks the GT list 32 at a time:
//     lane l tests GT g0+l against the warp box, one __ballot_sync gives the survivors, and the warp visits only
//     the set bits (ascending g, so "first argmax" survives).  Every skipped (g, anchor) pair has IoU exactly +0
//     and can never win the strict '>' against the running maximum, which starts at (0, index 0);
//   * row maxima (needed by allow_low_quality_matches) are folded per warp with redux.sync into a shared-memory
//     table; the CTA publishes its table (transposed, (B, G, tiles)) and atomicMax-es the global row maxima.
// assign_lq_kernel -- one CTA per (image, GT): finds the tiles whose maximum equals the row maximum and
//   re-evaluates only those 512-anchor segments (matcher.py:47-49), including rows whose maximum is 0.
// HBM traffic: 16 B/anchor read + 24 B/anchor written (labels, indices, offsets) per image.
#include "anchor_levels.cuh"

namespace bdet {

constexpr int kAT = 256;                   // threads per CTA
constexpr int kAPW = 64;                   // anchors per warp (2 per lane)
constexpr int kATile = (kAT / 32) * kAPW;  // 512 anchors per CTA
constexpr int kLqThreads = 128;

struct AssignArgs {
  const float* anchors;
  const float* gt;  // (B, Gmax, 5)
  const int* num_gt;
  int* labels;
  int* idx;
  float* offsets;
  uint32_t* rowmax;  // (B, Gmax) fp32 bits (IoU >= 0, so uint order == float order); zero-initialised
  uint32_t* blkmax;  // (B, Gmax, tiles)
  int* counts;       // optional (B, 3): labels < 0, == 0, > 0 (the census retinanet.py:142-146 / rpn.py:231 needs)
  int A, Gmax, tiles, allow_lq, apply_class, unit_coder;
  MatchCfg cfg;
  Vec4 mean, stdv;
};

// anchors == nullptr: the anchors are generated in registers from the grid description (the same expression as
// anchors_grid_kernel, bit for bit) -- no anchor tensor, no separate launch.
__device__ __forceinline__ float4 assign_anchor(const AssignArgs& p, const AnchorLevels& lv, long long c) {
  return p.anchors ? ldg4(p.anchors + c * 4) : anchor_at(lv, c, nullptr);
}

__global__ void __launch_bounds__(kAT) assign_main_kernel(const AssignArgs p, const __grid_constant__ AnchorLevels lv) {
  extern __shared__ __align__(16) unsigned char raw[];
  float4* sbox = reinterpret_cast<float4*>(raw);
  float* sarea = reinterpret_cast<float*>(sbox + p.Gmax);
  float* scls = sarea + p.Gmax;
  uint32_t* srmax = reinterpret_cast<uint32_t*>(scls + p.Gmax);  // (8 warps, Gmax): each warp visits a GT at most once
  const int b = blockIdx.y, tile = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int G = min(p.num_gt[b], p.Gmax);

  const float* gt = p.gt + (long long)b * p.Gmax * 5;
  for (int g = t; g < G; g += kAT) {
    const float* r = gt + g * 5;
    float4 bx = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
    sbox[g] = bx;
    sarea[g] = box_area(bx);
    scls[g] = __ldg(r + 4);
  }
  if (p.allow_lq)
    for (int i = t; i < (kAT / 32) * p.Gmax; i += kAT) srmax[i] = 0u;
  // this lane's two anchors: consecutive runs of 32 inside the warp's 64
  const long long c0 = (long long)tile * kATile + warp * kAPW + lane;
  const long long c1 = c0 + 32;
  const bool ok0 = c0 < p.A, ok1 = c1 < p.A;
  const float4 an0 = ok0 ? assign_anchor(p, lv, c0) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 an1 = ok1 ? assign_anchor(p, lv, c1) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float aa0 = box_area(an0), aa1 = box_area(an1);
  // warp bounding box (NaN coordinates are ignored by fmin/fmax; such anchors give IoU 0 against everything)
  float mnx = CUDART_INF_F, mny = CUDART_INF_F, mxx = -CUDART_INF_F, mxy = -CUDART_INF_F;
  if (ok0) {
    mnx = an0.x;
    mny = an0.y;
    mxx = an0.z;
    mxy = an0.w;
  }
  if (ok1) {
    mnx = fminf(mnx, an1.x);
    mny = fminf(mny, an1.y);
    mxx = fmaxf(mxx, an1.z);
    mxy = fmaxf(mxy, an1.w);
  }
  const float bb0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(mnx)));
  const float bb1 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(mny)));
  const float bb2 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(mxx)));
  const float bb3 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(mxy)));
  __syncthreads();

  // Running (max, first argmax) over G.  All IoUs are >= +0 and never NaN, so the state after the (possibly
  // skipped) row 0 is at least (0, 0); skipped rows are exact zeros and cannot win '>'.
  float best0 = 0.f, best1 = 0.f;
  int bi0 = 0, bi1 = 0;
  for (int g0 = 0; g0 < G; g0 += 32) {
    const int gl = g0 + lane;
    bool live = false;
    if (gl < G) {
      const float4 a = sbox[gl];
      live = !(a.z <= bb0 || a.x >= bb2 || a.w <= bb1 || a.y >= bb3);  // NaN GT -> live (evaluated, gives 0)
    }
    uint32_t m = __ballot_sync(0xffffffffu, live);
    while (m) {
      const int g = g0 + __ffs(m) - 1;
      m &= m - 1;
      const float4 a = sbox[g];
      const float ga = sarea[g];
      const float v0 = iou_pair(a, ga, an0, aa0);  // out-of-range lanes hold a zero-size anchor: IoU 0 with anything
      const float v1 = iou_pair(a, ga, an1, aa1);
      if (v0 > best0) {
        best0 = v0;
        bi0 = g;
      }
      if (v1 > best1) {
        best1 = v1;
        bi1 = g;
      }
      if (p.allow_lq) {
        const float rv = fmaxf(v0, v1);
        if (__any_sync(0xffffffffu, rv > 0.f)) {
          const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(rv));
          if (lane == 0) srmax[warp * p.Gmax + g] = w;
        }
      }
    }
  }

  int n_neg = 0, n_zero = 0, n_pos = 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const bool ok = j ? ok1 : ok0;
    if (!ok) continue;
    const long long o = (long long)b * p.A + (j ? c1 : c0);
    const float best = j ? best1 : best0;
    const int bi = j ? bi1 : bi0;
    int label = threshold_label(p.cfg, best);
    float4 off = make_float4(0.f, 0.f, 0.f, 0.f);
    if (G > 0) {
      if (p.apply_class && label == 1) label = (int)scls[bi];  // retinanet.py:222-223 astype("int32")
      off = p.unit_coder ? encode_box<true>(j ? an1 : an0, sbox[bi], p.mean, p.stdv)
                         : encode_box<false>(j ? an1 : an0, sbox[bi], p.mean, p.stdv);
    }
    p.labels[o] = label;
    p.idx[o] = bi;
    reinterpret_cast<float4*>(p.offsets)[o] = off;
    n_neg += label < 0;
    n_zero += label == 0;
    n_pos += label > 0;
  }
  if (p.counts) {  // per-warp totals, three atomics per warp
    n_neg = __reduce_add_sync(0xffffffffu, n_neg);
    n_zero = __reduce_add_sync(0xffffffffu, n_zero);
    n_pos = __reduce_add_sync(0xffffffffu, n_pos);
    if (lane == 0) {
      if (n_neg) atomicAdd(p.counts + b * 3, n_neg);
      if (n_zero) atomicAdd(p.counts + b * 3 + 1, n_zero);
      if (n_pos) atomicAdd(p.counts + b * 3 + 2, n_pos);
    }
  }

  if (p.allow_lq) {
    __syncthreads();
    uint32_t* bm = p.blkmax + (long long)b * p.Gmax * p.tiles + tile;
    for (int g = t; g < G; g += kAT) {
      uint32_t u = 0u;
#pragma unroll
      for (int w = 0; w < kAT / 32; ++w) u = max(u, srmax[w * p.Gmax + g]);
      bm[(long long)g * p.tiles] = u;
      if (u) atomicMax(&p.rowmax[(long long)b * p.Gmax + g], u);
    }
  }
}

// allow_low_quality_matches, matcher.py:47-49: every anchor whose IoU with g equals the row maximum of g gets
// label 1 (then the class of ITS OWN matched GT, retinanet.py:222-223).
__global__ void __launch_bounds__(kLqThreads) assign_lq_kernel(const AssignArgs p, const __grid_constant__ AnchorLevels lv) {
  extern __shared__ int stiles[];  // tiles
  __shared__ int nhit;
  const int g = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const int G = min(p.num_gt[b], p.Gmax);
  if (g >= G) return;
  if (t == 0) nhit = 0;
  __syncthreads();
  const uint32_t rm = p.rowmax[(long long)b * p.Gmax + g];
  const uint32_t* bm = p.blkmax + ((long long)b * p.Gmax + g) * p.tiles;
  for (int i = t; i < p.tiles; i += kLqThreads)
    if (bm[i] == rm) stiles[atomicAdd(&nhit, 1)] = i;  // rm == 0: every tile (all their zero-IoU anchors qualify)
  __syncthreads();
  const int nh = nhit;
  if (nh == 0) return;
  const float* r = p.gt + ((long long)b * p.Gmax + g) * 5;
  const float4 a = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
  const float ga = box_area(a);
  const float* gtb = p.gt + (long long)b * p.Gmax * 5;
  for (int h = 0; h < nh; ++h) {
    const long long base = (long long)stiles[h] * kATile;
    for (int i = t; i < kATile; i += kLqThreads) {
      const long long c = base + i;
      if (c >= p.A) break;
      const float4 an = assign_anchor(p, lv, c);
      const float v = iou_pair(a, ga, an, box_area(an));
      if (__float_as_uint(v) == rm) {
        const long long o = (long long)b * p.A + c;
        const int nl = p.apply_class ? (int)__ldg(gtb + p.idx[o] * 5 + 4) : 1;
        if (p.counts) {  // several GT rows may promote the same anchor: the exchange tells who changed the census
          const int old = atomicExch(p.labels + o, nl);
          const int oc = old < 0 ? 0 : (old == 0 ? 1 : 2), nc = nl < 0 ? 0 : (nl == 0 ? 1 : 2);
          if (oc != nc) {
            atomicSub(p.counts + b * 3 + oc, 1);
            atomicAdd(p.counts + b * 3 + nc, 1);
          }
        } else {
          p.labels[o] = nl;
        }
      }
    }
  }
}

}  // namespace bdet

using namespace bdet;

extern "C" size_t bdet_assign_targets_workspace(int Gmax, int A, int B) {
  if (Gmax <= 0 || A <= 0 || B <= 0) return 16;
  const int tiles = ceil_div(A, kATile);
  return align_up((size_t)B * Gmax * 4, 256) + (size_t)B * Gmax * tiles * 4 + 256;
}

static int assign_targets_impl(const float* anchors, const AnchorLevels* grid, int A, const float* gt, int Gmax,
                               const int* num_gt_dev, int B, const float* thresholds_host, const int* labels_host, int n_labels,
                               int allow_low_quality, int apply_class, const float* mean_host, const float* std_host,
                               int* labels, int* match_idx, float* offsets, int* counts, void* workspace,
                               size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(A >= 0 && Gmax >= 0 && B >= 0, "negative size");
  AssignArgs a;
  int rc = make_match_cfg(&a.cfg, thresholds_host, labels_host, n_labels);
  if (rc) return rc;
  if (A == 0 || B == 0) return BDET_OK;
  BDET_REQUIRE((anchors || grid) && labels && match_idx && offsets && num_gt_dev, "null argument");
  BDET_REQUIRE(Gmax == 0 || gt, "null gt");
  BDET_REQUIRE((!anchors || aligned16(anchors)) && aligned16(offsets), "anchors/offsets must be 16-byte aligned");
  BDET_REQUIRE(B <= 65535 && Gmax <= 65535, "B / Gmax > 65535");
  const size_t smem = (size_t)max(Gmax, 1) * (24 + 4 * (kAT / 32));
  if (smem > 200 * 1024) return set_error(BDET_EUNSUPPORTED, "bdet_assign_targets: Gmax > 3600 does not fit shared memory");
  a.anchors = anchors;
  a.gt = gt;
  a.num_gt = num_gt_dev;
  a.labels = labels;
  a.idx = match_idx;
  a.offsets = offsets;
  a.A = A;
  a.Gmax = Gmax;
  a.tiles = ceil_div(A, kATile);
  a.allow_lq = (allow_low_quality != 0 && Gmax > 0) ? 1 : 0;
  a.apply_class = apply_class != 0;
  a.unit_coder = 1;
  for (int i = 0; i < 4; ++i) {
    a.mean.v[i] = mean_host ? mean_host[i] : 0.f;
    a.stdv.v[i] = std_host ? std_host[i] : 1.f;
    if (a.mean.v[i] != 0.f || a.stdv.v[i] != 1.f) a.unit_coder = 0;  // (t - 0) / 1 == t bit for bit
  }
  a.rowmax = nullptr;
  a.blkmax = nullptr;
  a.counts = counts;
  cudaStream_t st = as_stream(stream);
  if (counts) BDET_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * 3 * 4, st));
  static const AnchorLevels kNoGrid = {};
  const AnchorLevels& lv = grid ? *grid : kNoGrid;
  if (a.allow_lq) {
    const size_t need = bdet_assign_targets_workspace(Gmax, A, B);
    if (!workspace || workspace_bytes < need)
      return set_error(BDET_EWORKSPACE, "bdet_assign_targets: workspace needs %zu bytes", need);
    BDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 3u) == 0, "workspace must be 4-byte aligned");
    char* w = reinterpret_cast<char*>(workspace);
    a.rowmax = reinterpret_cast<uint32_t*>(w);
    a.blkmax = reinterpret_cast<uint32_t*>(w + align_up((size_t)B * Gmax * 4, 256));
    BDET_CUDA(cudaMemsetAsync(a.rowmax, 0, (size_t)B * Gmax * 4, st));
  }
  if (smem > 40 * 1024)
    BDET_CUDA(cudaFuncSetAttribute(assign_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BDET_KERNEL("assign_main_kernel", st, assign_main_kernel<<<dim3(a.tiles, B), kAT, smem, st>>>(a, lv));
  if (a.allow_lq) {
    const size_t lq_smem = (size_t)a.tiles * 4;
    if (lq_smem > 40 * 1024)
      BDET_CUDA(cudaFuncSetAttribute(assign_lq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lq_smem));
    BDET_KERNEL("assign_lq_kernel", st, assign_lq_kernel<<<dim3(Gmax, B), kLqThreads, lq_smem, st>>>(a, lv));
  }
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_assign_targets(const float* anchors, int A, const float* gt, int Gmax, const int* num_gt_dev, int B,
                                   const float* thresholds_host, const int* labels_host, int n_labels,
                                   int allow_low_quality, int apply_class, const float* mean_host,
                                   const float* std_host, int* labels, int* match_idx, float* offsets,
                                   void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  return assign_targets_impl(anchors, nullptr, A, gt, Gmax, num_gt_dev, B, thresholds_host, labels_host, n_labels,
                             allow_low_quality, apply_class, mean_host, std_host, labels, match_idx, offsets, nullptr, workspace,
                             workspace_bytes, stream);
}

extern "C" int bdet_assign_targets_grid(int n_levels, const int* hw_host, const double* stride_host, const double* shift_host,
                                        const int* n_base_host, const float* base_host, const float* gt, int Gmax,
                                        const int* num_gt_dev, int B, const float* thresholds_host, const int* labels_host,
                                        int n_labels, int allow_low_quality, int apply_class, const float* mean_host,
                                        const float* std_host, int* labels, int* match_idx, float* offsets, int* counts,
                                        void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(hw_host && stride_host && n_base_host && base_host, "null argument");
  static thread_local AnchorLevels lv;
  static thread_local int64_t off[BDET_MAX_LEVELS];
  if (n_levels < 1 || n_levels > BDET_MAX_LEVELS) return set_error(BDET_EINVAL, "bdet_assign_targets_grid: n_levels must be in [1, %d]", BDET_MAX_LEVELS);
  int64_t o = 0;
  for (int l = 0; l < n_levels; ++l) {
    off[l] = o;
    o += (int64_t)hw_host[2 * l] * hw_host[2 * l + 1] * n_base_host[l];
  }
  int rc = fill_levels(&lv, n_levels, hw_host, stride_host, shift_host, n_base_host, 0, base_host, off);
  if (rc) return rc;
  if (o > 0x7fffffffLL) return set_error(BDET_EUNSUPPORTED, "bdet_assign_targets_grid: more than 2^31 anchors");
  return assign_targets_impl(nullptr, &lv, (int)o, gt, Gmax, num_gt_dev, B, thresholds_host, labels_host, n_labels,
                             allow_low_quality, apply_class, mean_host, std_host, labels, match_idx, offsets, counts, workspace,
                             workspace_bytes, stream);
}
