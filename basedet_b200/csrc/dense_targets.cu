// target assignment of the anchor-free dense heads, fused (no (G, A) tensor is materialised).
// Reference: basedet/models/det/fcos.py:222-293 (FCOS.get_ground_truth: level range of max(l,t,r,b), centre sampling,
//            smallest-area GT, PointCoder.encode, centerness)
//            basedet/models/det/atss.py:17-86   (ATSS.get_ground_truth: per-level top-k nearest points, mean+std IoU
//            threshold, in-box test, argmax IoU, PointCoder.encode, centerness).
// The reference builds (G, A, 4) offsets, (G, A) masks / areas / IoUs per image and reduces over G; here one thread owns
// one point (FCOS) or one CTA owns a group of four GTs (ATSS candidate search) and the reductions run in registers /
// shared memory.
// Arithmetic follows the reference op order (one fp32 rounding per op, -fmad=false), elementwise max/min are MegDNN's
// x>y?x:y / x<y?x:y (oracle ASSUMED-1), argmin / argmax keep the FIRST index (ASSUMED-2/9), the ATSS threshold sums its
// candidates sequentially in index order (ASSUMED-8).
#include <limits>

#include "common.cuh"

namespace bdet {

__device__ __forceinline__ float emaxf(float x, float y) { return x > y ? x : y; }  // MegDNN Elemwise MAX
__device__ __forceinline__ float eminf(float x, float y) { return x < y ? x : y; }  // MegDNN Elemwise MIN

struct DenseArgs {
  const float* points;  // (A, 2)
  const float* gt;      // (B, Gmax, 5)
  const int* num_gt;    // (B)
  int A, Gmax, B, L;
  int lvl_start[BDET_MAX_LEVELS + 1];
  float lo[BDET_MAX_LEVELS], hi[BDET_MAX_LEVELS];  // FCOS object sizes of interest
  float radius[BDET_MAX_LEVELS];                    // FCOS stride * center_sampling_radius; ATSS stride * scale / 2
  int use_center;                                   // FCOS: centre sampling on
  int topk;                                         // ATSS
  int* labels;                                      // (B, A)
  float* offsets;                                   // (B, A, 4)
  float* ctrness;                                   // (B, A)
  int* match_idx;                                   // (B, A) optional
  unsigned long long* best;                         // ATSS (B, A): (IoU bits << 32) | ~g of the best proposal, 0 = none
};

__device__ __forceinline__ int level_of(const DenseArgs& p, int a) {
  int l = 0;
  while (l + 1 < p.L && a >= p.lvl_start[l + 1]) ++l;
  return l;
}

// labels / PointCoder.encode / centerness of one point against its matched GT (fcos.py:276-287, atss.py:67-77)
__device__ __forceinline__ void write_point(const DenseArgs& p, int b, int a, float px, float py, const float* gtb, int G,
                                            int bi, bool fg) {
  const long long o = (long long)b * p.A + a;
  int label = 0;
  float4 off = make_float4(0.f, 0.f, 0.f, 0.f);
  float ctr = 0.f;
  if (G > 0) {
    const float* r = gtb + (long long)bi * 5;
    const float x1 = __ldg(r), y1 = __ldg(r + 1), x2 = __ldg(r + 2), y2 = __ldg(r + 3);
    if (fg) label = (int)__ldg(r + 4);  // gt_boxes_matched[:, 4].astype("int32"); background rows are set to 0
    off = make_float4(px - x1, py - y1, x2 - px, y2 - py);  // boxcoder.py:132-133
    const float qa = __fdiv_rn(fminf(off.x, off.z), fmaxf(off.x, off.z));
    const float qb = __fdiv_rn(fminf(off.y, off.w), fmaxf(off.y, off.w));
    ctr = sqrtf(emaxf(qa, 0.f) * emaxf(qb, 0.f));
  }
  p.labels[o] = label;
  reinterpret_cast<float4*>(p.offsets)[o] = off;
  p.ctrness[o] = ctr;
  if (p.match_idx) p.match_idx[o] = bi;
}

// ------------------------------------------------------------------------------------------------ FCOS
constexpr int kDenseThreads = 256;

__global__ void __launch_bounds__(kDenseThreads) fcos_targets_kernel(const DenseArgs p) {
  extern __shared__ __align__(16) unsigned char raw[];
  float4* sbox = reinterpret_cast<float4*>(raw);              // Gmax
  float2* sctr = reinterpret_cast<float2*>(sbox + p.Gmax);    // Gmax: box centres, op_patch.py:100-113
  float* sarea = reinterpret_cast<float*>(sctr + p.Gmax);     // Gmax: Boxes.area, boxes.py:36-42
  const int b = blockIdx.y, t = threadIdx.x;
  const int G = min(p.num_gt[b], p.Gmax);
  const float* gtb = p.gt + (long long)b * p.Gmax * 5;
  for (int g = t; g < G; g += kDenseThreads) {
    const float* r = gtb + g * 5;
    const float4 bx = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
    sbox[g] = bx;
    sctr[g] = make_float2(__fdiv_rn(bx.x + bx.z, 2.f), __fdiv_rn(bx.y + bx.w, 2.f));
    sarea[g] = (bx.z - bx.x) * (bx.w - bx.y);
  }
  __syncthreads();
  const int a = blockIdx.x * kDenseThreads + t;
  const bool valid = a < p.A;
  const int lane = t & 31;
  const float2 pt = valid ? __ldg(reinterpret_cast<const float2*>(p.points) + a) : make_float2(0.f, 0.f);
  const int lv = level_of(p, valid ? a : p.A - 1);
  const float lo = p.lo[lv], hi = p.hi[lv], rad = p.radius[lv];
  const float inf = CUDART_INF_F;
  // Warp pruning: a GT can only own a point that lies strictly inside its (centre) box, so a GT whose box misses the
  // bounding box of the warp's 32 points is skipped for the whole warp.  Visiting the survivors in ascending g keeps
  // the first-index argmin.  The warp's largest radius makes the test conservative across a level boundary.
  const float bx0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(valid ? pt.x : inf)));
  const float by0 = ord2f(__reduce_min_sync(0xffffffffu, f2ord(valid ? pt.y : inf)));
  const float bx1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(valid ? pt.x : -inf)));
  const float by1 = ord2f(__reduce_max_sync(0xffffffffu, f2ord(valid ? pt.y : -inf)));
  const float wrad = ord2f(__reduce_max_sync(0xffffffffu, f2ord(rad)));
  float best = inf;
  int bi = 0;
  for (int g0 = 0; g0 < G; g0 += 32) {
    const int gl = g0 + lane;
    bool live = false;
    if (gl < G) {
      float4 q = sbox[gl];
      if (p.use_center) {
        const float2 c = sctr[gl];
        q = make_float4(emaxf(c.x - wrad, q.x), emaxf(c.y - wrad, q.y), eminf(c.x + wrad, q.z), eminf(c.y + wrad, q.w));
      }
      live = !(q.z <= bx0 || q.x >= bx1 || q.w <= by0 || q.y >= by1);  // NaN -> live (evaluated exactly below)
    }
    uint32_t m = __ballot_sync(0xffffffffu, live);
    while (m) {
      const int g = g0 + __ffs(m) - 1;
      m &= m - 1;
      const float4 bx = sbox[g];
      const float l = pt.x - bx.x, tt = pt.y - bx.y, r = bx.z - pt.x, bb = bx.w - pt.y;  // fcos.py:231
      const float mx = fmaxf(fmaxf(l, tt), fmaxf(r, bb));                                // :245
      bool ok = (mx >= lo) && (mx <= hi);                                                // :246-249
      if (p.use_center) {                                                                // :251-264
        const float2 c = sctr[g];
        const float cx1 = emaxf(c.x - rad, bx.x), cy1 = emaxf(c.y - rad, bx.y);
        const float cx2 = eminf(c.x + rad, bx.z), cy2 = eminf(c.y + rad, bx.w);
        const float mn = fminf(fminf(pt.x - cx1, pt.y - cy1), fminf(cx2 - pt.x, cy2 - pt.y));
        ok = ok && (mn > 0.f);
      } else {
        ok = ok && (fminf(fminf(l, tt), fminf(r, bb)) > 0.f);                            // :266
      }
      const float ar = ok ? sarea[g] : inf;                                              // :268-270
      if (ar < best) {                                                                   // :272 argmin, first index
        best = ar;
        bi = g;
      }
    }
  }
  if (!valid) return;
  write_point(p, b, a, pt.x, pt.y, gtb, G, bi, best != inf);                           // :274-287
}

// ------------------------------------------------------------------------------------------------ ATSS
constexpr int kAtssWarps = 8;
constexpr int kAtssThreads = kAtssWarps * 32;
constexpr int kMaxTopk = 16;
constexpr int kMaxCand = BDET_MAX_LEVELS * kMaxTopk;
constexpr int kNear = 128;  // points collected inside the distance bound (typically 10-30)

// Sorted insertion into a register-resident ascending list (the compiler keeps kk[] in registers: every index is static).
__device__ __forceinline__ void topk_insert(unsigned long long (&kk)[kMaxTopk], unsigned long long key) {
#pragma unroll
  for (int j = kMaxTopk - 1; j >= 0; --j) {
    const unsigned long long prev = j > 0 ? kk[j - 1] : 0ull;
    if (key < kk[j]) kk[j] = (j > 0 && key < prev) ? prev : key;
  }
}

// One CTA per (image, group of kGPC GTs): every loaded point is tested against the group's centres, which divides
// the L2 -> SM point traffic (the bound of a one-GT-per-CTA version) by kGPC.  Level by level the CTA finds, per GT,
// the `topk` points nearest to its centre (atss.py:39-44): a first sweep bounds the k-th distance, a second collects
// the handful of points inside the bound, rank counting orders them.  Then warp q evaluates GT q's candidates: IoUs,
// the mean + std threshold (:49-51), and proposes (IoU, g) to every candidate point that passes the threshold and lies
// inside the GT (:52-61) with one 64-bit atomicMax.
constexpr int kGPC = 4;
constexpr int kGroups = kAtssThreads / 4;  // four-lane groups

__global__ void __launch_bounds__(kAtssThreads) atss_candidates_kernel(const DenseArgs p) {
  __shared__ unsigned long long slist[kAtssWarps][32][kMaxTopk];  // overflow path only
  __shared__ unsigned long long swbest[kAtssWarps][kMaxTopk];
  __shared__ unsigned long long skeys[kGPC][kNear];
  __shared__ float sgmin[kGPC][kGroups];
  __shared__ float sU[kGPC];
  __shared__ int sn[kGPC];
  __shared__ int scand[kGPC][kMaxCand];
  __shared__ float siou[kGPC][kMaxCand];
  __shared__ int sslot[BDET_MAX_LEVELS + 1];
  const int b = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int G = min(p.num_gt[b], p.Gmax);
  const int g0 = blockIdx.x * kGPC;
  if (g0 >= G) return;
  const int ng = min(kGPC, G - g0);
  float cx[kGPC], cy[kGPC];
#pragma unroll
  for (int q = 0; q < kGPC; ++q) {  // slots past the last GT repeat it; their results are never used
    const float* r = p.gt + ((long long)b * p.Gmax + g0 + min(q, ng - 1)) * 5;
    cx[q] = __fdiv_rn(__ldg(r) + __ldg(r + 2), 2.f);      // gt_boxes.centers, op_patch.py:100-113
    cy[q] = __fdiv_rn(__ldg(r + 1) + __ldg(r + 3), 2.f);
  }
  if (t == 0) {
    int s = 0;
    for (int l = 0; l < p.L; ++l) {
      sslot[l] = s;
      s += min(p.topk, p.lvl_start[l + 1] - p.lvl_start[l]);
    }
    sslot[p.L] = s;
  }
  __syncthreads();
  const float2* pts = reinterpret_cast<const float2*>(p.points);
  for (int l = 0; l < p.L; ++l) {
    const int s = p.lvl_start[l], e = p.lvl_start[l + 1];
    const int k = min(p.topk, e - s);
    if (k == 0) continue;  // CTA-uniform
    // pass 1: squared distances (atss.py:40-42 without the monotone sqrt); every thread the minimum over its strided
    // slice; the k-th smallest of the 64 four-lane group minima bounds the k-th nearest squared distance
    float mn[kGPC];
#pragma unroll
    for (int q = 0; q < kGPC; ++q) mn[q] = CUDART_INF_F;
    for (int i = s + t; i < e; i += kAtssThreads) {
      const float2 pt = __ldg(pts + i);
#pragma unroll
      for (int q = 0; q < kGPC; ++q) {
        const float dx = cx[q] - pt.x, dy = cy[q] - pt.y;
        mn[q] = fminf(mn[q], dx * dx + dy * dy);
      }
    }
#pragma unroll
    for (int q = 0; q < kGPC; ++q) {
      float v = mn[q];
      v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1));
      v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 2));
      if ((lane & 3) == 0) sgmin[q][t >> 2] = v;
    }
    if (t < kGPC) sn[t] = 0;
    __syncthreads();
    {
      const int q = t / kGroups, j0 = t % kGroups;  // kGPC * kGroups == kAtssThreads
      const float v = sgmin[q][j0];
      int rank = 0;
      for (int j = 0; j < kGroups; ++j) {
        const float o = sgmin[q][j];
        rank += (o < v) || (o == v && j < j0);
      }
      if (rank == k - 1) sU[q] = v;  // +inf when fewer than k groups saw a point: everything is collected
    }
    __syncthreads();
    // sqrt is monotone, so U = sqrt(bound) bounds the k-th distance; a point can round to d <= U only if
    // d2 <= bound * (1 + 2^-21), the cheap pre-filter in front of the exact test
    float U2[kGPC], U[kGPC];
#pragma unroll
    for (int q = 0; q < kGPC; ++q) {
      U2[q] = sU[q] * 1.000001f;
      U[q] = sqrtf(sU[q]);
    }
    // pass 2: the few points within U, then their exact order (distance asc, index asc) by rank counting
    for (int i = s + t; i < e; i += kAtssThreads) {
      const float2 pt = __ldg(pts + i);
#pragma unroll
      for (int q = 0; q < kGPC; ++q) {
        const float dx = cx[q] - pt.x, dy = cy[q] - pt.y;
        const float d2 = dx * dx + dy * dy;
        if (d2 > U2[q]) continue;
        const float d = sqrtf(d2);
        if (!(d > U[q])) {
          const int slot = atomicAdd(&sn[q], 1);
          if (slot < kNear) skeys[q][slot] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(i - s);
        }
      }
    }
    __syncthreads();
    for (int q = 0; q < ng; ++q) {
      const int m = sn[q];
      if (m <= kNear) {
        if (t < m) {
          const unsigned long long key = skeys[q][t];
          int rank = 0;
          for (int j = 0; j < m; ++j) rank += skeys[q][j] < key;
          if (rank < k) scand[q][sslot[l] + rank] = s + (int)(unsigned)key;  // :43-44 base + topk_idxs (ASSUMED-9)
        }
        continue;
      }
      // more than kNear points inside U (masses of equidistant / duplicate points): per-lane sorted lists, 32 lane
      // lists -> the warp's k best, 8 warp lists -> the level's k candidates
      unsigned long long kk[kMaxTopk];
#pragma unroll
      for (int j = 0; j < kMaxTopk; ++j) kk[j] = ~0ull;
      float ccx = cx[0], ccy = cy[0];
#pragma unroll
      for (int qq = 1; qq < kGPC; ++qq)
        if (qq == q) {
          ccx = cx[qq];
          ccy = cy[qq];
        }
      for (int i = s + t; i < e; i += kAtssThreads) {
        const float2 pt = __ldg(pts + i);
        const float dx = ccx - pt.x, dy = ccy - pt.y;
        const unsigned long long key = ((unsigned long long)__float_as_uint(sqrtf(dx * dx + dy * dy)) << 32) | (unsigned)(i - s);
        if (key < kk[kMaxTopk - 1]) topk_insert(kk, key);
      }
#pragma unroll
      for (int j = 0; j < kMaxTopk; ++j) slist[warp][lane][j] = kk[j];
      __syncwarp();
      int head = 0;
      for (int round = 0; round < k; ++round) {
        const unsigned long long mine = head < kMaxTopk ? slist[warp][lane][head] : ~0ull;
        const unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)(mine >> 32));
        const unsigned lo = __reduce_min_sync(0xffffffffu, (unsigned)(mine >> 32) == hi ? (unsigned)mine : 0xffffffffu);
        if ((unsigned)(mine >> 32) == hi && (unsigned)mine == lo) {
          ++head;
          swbest[warp][round] = mine;
        }
      }
      __syncthreads();
      if (warp == 0) {
        int wh = 0;
        for (int round = 0; round < k; ++round) {
          const unsigned long long mine = (lane < kAtssWarps && wh < k) ? swbest[lane][wh] : ~0ull;
          const unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)(mine >> 32));
          const unsigned lo = __reduce_min_sync(0xffffffffu, (unsigned)(mine >> 32) == hi ? (unsigned)mine : 0xffffffffu);
          if ((unsigned)(mine >> 32) == hi && (unsigned)mine == lo) {
            ++wh;
            scand[q][sslot[l] + round] = s + (int)lo;
          }
        }
      }
      __syncthreads();
    }
    __syncthreads();
  }
  if (warp >= ng) return;
  // ---- warp q: candidates of GT g0 + q
  const int q = warp, g = g0 + q;
  const float* r = p.gt + ((long long)b * p.Gmax + g) * 5;
  const float4 bx = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
  const int nc = sslot[p.L];
  const float garea = box_area(bx);
  for (int j = lane; j < nc; j += 32) {
    const int a = scand[q][j];
    const int l = level_of(p, a);
    const float2 pt = __ldg(pts + a);
    const float h = p.radius[l];
    const float4 ab = make_float4(pt.x - h, pt.y - h, pt.x + h, pt.y + h);  // :31-37
    siou[q][j] = iou_pair(bx, garea, ab, box_area(ab));
  }
  __syncwarp();
  float thr = 0.f;
  if (lane == 0) {  // :50-51, sequential fp32 accumulation (ASSUMED-8)
    float sum = 0.f;
    for (int j = 0; j < nc; ++j) sum += siou[q][j];
    const float mean = __fdiv_rn(sum, (float)nc);
    float sq = 0.f;
    for (int j = 0; j < nc; ++j) {
      const float d = siou[q][j] - mean;
      sq += d * d;
    }
    thr = mean + sqrtf(__fdiv_rn(sq, (float)nc));
  }
  thr = __shfl_sync(0xffffffffu, thr, 0);
  for (int j = lane; j < nc; j += 32) {
    const float v = siou[q][j];
    if (!(v >= thr)) continue;  // :52-54
    const int a = scand[q][j];
    const float2 pt = __ldg(pts + a);
    const float mn = fminf(fminf(pt.x - bx.x, pt.y - bx.y), fminf(bx.z - pt.x, bx.w - pt.y));
    if (!(mn > 0.f)) continue;  // :56-58
    const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | (0xffffffffu - (unsigned)g);
    atomicMax(p.best + (long long)b * p.A + a, key);  // :63 argmax over G: highest IoU, lowest g among equals
  }
}

__global__ void __launch_bounds__(kDenseThreads) atss_finish_kernel(const DenseArgs p) {
  const int b = blockIdx.y;
  const int a = blockIdx.x * kDenseThreads + threadIdx.x;
  if (a >= p.A) return;
  const int G = min(p.num_gt[b], p.Gmax);
  const unsigned long long key = p.best[(long long)b * p.A + a];
  const bool fg = key != 0ull;  // anchor_max_iou == -1 -> background, match index 0 (atss.py:63-68)
  const int bi = fg ? (int)(0xffffffffu - (unsigned)key) : 0;
  const float2 pt = __ldg(reinterpret_cast<const float2*>(p.points) + a);
  write_point(p, b, a, pt.x, pt.y, p.gt + (long long)b * p.Gmax * 5, G, bi, fg);
}

static int fill_common(DenseArgs& a, const char* who, const float* points, int A, const int* level_start_host, int L,
                       const float* gt, int Gmax, const int* num_gt_dev, int B, int* labels, float* offsets, float* ctrness,
                       int* match_idx) {
  if (A < 0 || Gmax < 0 || B < 0 || L < 1 || L > BDET_MAX_LEVELS) return set_error(BDET_EINVAL, "%s: bad sizes", who);
  if (!level_start_host || level_start_host[0] != 0 || level_start_host[L] != A)
    return set_error(BDET_EINVAL, "%s: level_start must run from 0 to A", who);
  for (int l = 0; l < L; ++l)
    if (level_start_host[l + 1] < level_start_host[l]) return set_error(BDET_EINVAL, "%s: level_start must be non-decreasing", who);
  if (A == 0 || B == 0) return BDET_OK;
  if (!points || !num_gt_dev || !labels || !offsets || !ctrness || (Gmax > 0 && !gt))
    return set_error(BDET_EINVAL, "%s: null argument", who);
  if (!aligned16(offsets) || (reinterpret_cast<uintptr_t>(points) & 7u))
    return set_error(BDET_EINVAL, "%s: offsets must be 16-byte and points 8-byte aligned", who);
  if (B > 65535 || Gmax > 65535) return set_error(BDET_EINVAL, "%s: B / Gmax > 65535", who);
  a.points = points;
  a.gt = gt;
  a.num_gt = num_gt_dev;
  a.A = A;
  a.Gmax = Gmax;
  a.B = B;
  a.L = L;
  for (int l = 0; l <= L; ++l) a.lvl_start[l] = level_start_host[l];
  a.labels = labels;
  a.offsets = offsets;
  a.ctrness = ctrness;
  a.match_idx = match_idx;
  a.best = nullptr;
  a.use_center = 0;
  a.topk = 0;
  return BDET_OK;
}

}  // namespace bdet

using namespace bdet;

extern "C" int bdet_fcos_targets(const float* points, int A, const int* level_start_host, int L, const float* radius_host,
                                 const float* size_lo_host, const float* size_hi_host, const float* gt, int Gmax,
                                 const int* num_gt_dev, int B, int* labels, float* offsets, float* ctrness, int* match_idx,
                                 bdet_stream_t stream) {
  DenseArgs a;
  int rc = fill_common(a, "bdet_fcos_targets", points, A, level_start_host, L, gt, Gmax, num_gt_dev, B, labels, offsets, ctrness,
                       match_idx);
  if (rc || A == 0 || B == 0) return rc;
  BDET_REQUIRE(size_lo_host && size_hi_host, "null size ranges");
  for (int l = 0; l < L; ++l) {
    a.lo[l] = size_lo_host[l];
    a.hi[l] = size_hi_host[l];
    a.radius[l] = radius_host ? radius_host[l] : 0.f;
    if (radius_host && radius_host[l] > 0.f) a.use_center = 1;
  }
  cudaStream_t st = as_stream(stream);
  const size_t smem = (size_t)Gmax * (16 + 8 + 4);
  if (smem > 40 * 1024)
    BDET_CUDA(cudaFuncSetAttribute(fcos_targets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BDET_KERNEL("fcos_targets_kernel", st, fcos_targets_kernel<<<dim3(ceil_div(A, kDenseThreads), B), kDenseThreads, smem, st>>>(a));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" size_t bdet_atss_targets_workspace(int A, int B) {
  if (A <= 0 || B <= 0) return 16;
  return (size_t)A * B * 8 + 256;
}

extern "C" int bdet_atss_targets(const float* points, int A, const int* level_start_host, int L, const float* half_size_host,
                                 int topk, const float* gt, int Gmax, const int* num_gt_dev, int B, int* labels,
                                 float* offsets, float* ctrness, int* match_idx, void* workspace, size_t workspace_bytes,
                                 bdet_stream_t stream) {
  DenseArgs a;
  int rc = fill_common(a, "bdet_atss_targets", points, A, level_start_host, L, gt, Gmax, num_gt_dev, B, labels, offsets, ctrness,
                       match_idx);
  if (rc || A == 0 || B == 0) return rc;
  BDET_REQUIRE(half_size_host, "null half sizes");
  BDET_REQUIRE(topk >= 1 && topk <= kMaxTopk, "topk must be in [1, 16]");
  const size_t need = bdet_atss_targets_workspace(A, B);
  if (!workspace || workspace_bytes < need) return set_error(BDET_EWORKSPACE, "bdet_atss_targets: workspace needs %zu bytes", need);
  BDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "workspace must be 8-byte aligned");
  for (int l = 0; l < L; ++l) a.radius[l] = half_size_host[l];
  a.topk = topk;
  a.best = reinterpret_cast<unsigned long long*>(workspace);
  cudaStream_t st = as_stream(stream);
  BDET_CUDA(cudaMemsetAsync(a.best, 0, (size_t)A * B * 8, st));
  if (Gmax > 0) BDET_KERNEL("atss_candidates_kernel", st, atss_candidates_kernel<<<dim3(ceil_div(Gmax, kGPC), B), kAtssThreads, 0, st>>>(a));
  BDET_KERNEL("atss_finish_kernel", st, atss_finish_kernel<<<dim3(ceil_div(A, kDenseThreads), B), kDenseThreads, 0, st>>>(a));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
