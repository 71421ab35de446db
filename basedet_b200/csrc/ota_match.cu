// SURVEY 8(f) rank 3 (matching part): OTATopkMatcher, the dynamic-k matcher of OTA / YOLOX.
// Reference: basedet/layers/common/matcher.py:134-161.  Inputs are the (G, A) cost and IoU matrices the caller's losses
// produce (models/det/ota.py:76-180); per GT the k = clip(int(sum of its candidate_k largest IoUs), 1) anchors of
// smallest cost are matched, an anchor matched by several GTs goes to the GT of smallest cost (over ALL GTs, :156),
// unmatched anchors get the background index G.
//   ota_rows_kernel   : one CTA per GT row: both per-row selections are "k <= 16 smallest keys of a long row" -- a first
//                       sweep bounds the k-th key from four-lane group minima, a second collects the few keys inside
//                       the bound, rank counting orders them (the pattern of dense_targets.cu) -- then one atomicAdd per
//                       matched anchor on its match counter;
//   ota_resolve_kernel: one thread per anchor: 0 matches -> G, 1 -> that GT, several -> argmin of the cost column.
// Order contract (oracle ASSUMED-2/3/9): top-k descending = (value desc, index asc), ascending = (value asc, index
// asc), argmin / argmax = first index; the top-k IoU sum is accumulated sequentially in descending order (ASSUMED-8).
#include "common.cuh"

namespace bdet {

constexpr int kOtaThreads = 256;
constexpr int kOtaMaxK = 16;
constexpr int kOtaNear = 256;
constexpr int kOtaGroups = kOtaThreads / 4;

struct OtaArgs {
  const float* cost;  // (G, A) row stride ldc
  const float* ious;  // (G, A) row stride ldi
  int G, A, ldc, ldi, k;
  int* cnt;           // (A) number of GTs that matched the anchor
  int* who;           // (A) one of them (the only one when cnt == 1)
  int* out;           // (A)
};

struct OtaSmem {
  unsigned long long gmin[kOtaGroups];
  unsigned long long near[kOtaNear];
  unsigned long long bound;
  int n;
};

// The k smallest keys of key(i), i < n, ascending, into res[0..k) (k <= kOtaMaxK <= n); every thread of the CTA calls it.
template <class KeyFn>
__device__ void row_select_small(const KeyFn& key, int n, int k, OtaSmem& sm, unsigned long long* res) {
  const int t = threadIdx.x, lane = t & 31;
  unsigned long long mn = ~0ull;
  for (int i = t; i < n; i += kOtaThreads) mn = min(mn, key(i));
  mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, 1));
  mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, 2));
  if ((lane & 3) == 0) sm.gmin[t >> 2] = mn;
  if (t == 0) sm.n = 0;
  __syncthreads();
  if (t < kOtaGroups) {  // k-th smallest group minimum: k distinct keys lie at or below it
    const unsigned long long v = sm.gmin[t];
    int rank = 0;
    for (int j = 0; j < kOtaGroups; ++j) {
      const unsigned long long o = sm.gmin[j];
      rank += (o < v) || (o == v && j < t);
    }
    if (rank == k - 1) sm.bound = v;
  }
  __syncthreads();
  const unsigned long long U = sm.bound;
  for (int i = t; i < n; i += kOtaThreads) {
    const unsigned long long q = key(i);
    if (q <= U) {
      const int slot = atomicAdd(&sm.n, 1);
      if (slot < kOtaNear) sm.near[slot] = q;
    }
  }
  __syncthreads();
  const int m = sm.n;
  if (m <= kOtaNear) {
    if (t < m) {
      const unsigned long long q = sm.near[t];
      int rank = 0;
      for (int j = 0; j < m; ++j) rank += sm.near[j] < q;
      if (rank < k) res[rank] = q;
    }
  } else {
    // more keys inside the bound than the list holds (one thread's slice full of small values): k rounds of
    // "smallest key above the previous one"
    unsigned long long last = 0ull;
    for (int r = 0; r < k; ++r) {
      unsigned long long best = ~0ull;
      for (int i = t; i < n; i += kOtaThreads) {
        const unsigned long long q = key(i);
        if ((r == 0 || q > last) && q < best) best = q;
      }
      best = min(best, __shfl_xor_sync(0xffffffffu, best, 16));
      best = min(best, __shfl_xor_sync(0xffffffffu, best, 8));
      best = min(best, __shfl_xor_sync(0xffffffffu, best, 4));
      best = min(best, __shfl_xor_sync(0xffffffffu, best, 2));
      best = min(best, __shfl_xor_sync(0xffffffffu, best, 1));
      __syncthreads();
      if (lane == 0) sm.gmin[t >> 5] = best;
      __syncthreads();
      best = sm.gmin[0];
      for (int w = 1; w < kOtaThreads / 32; ++w) best = min(best, sm.gmin[w]);
      if (t == 0) res[r] = best;
      last = best;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kOtaThreads) ota_rows_kernel(const OtaArgs p) {
  __shared__ OtaSmem sm;
  __shared__ unsigned long long res[kOtaMaxK];
  __shared__ int sdyn;
  const int g = blockIdx.x, t = threadIdx.x;
  const float* irow = p.ious + (long long)g * p.ldi;
  const float* crow = p.cost + (long long)g * p.ldc;
  const int k = min(p.k, p.A);
  // candidate_k largest IoUs: key order = (value desc, index asc)  (matcher.py:145)
  auto ikey = [&](int i) { return ((unsigned long long)(~f2ord(__ldg(irow + i) + 0.f)) << 32) | (unsigned)i; };
  row_select_small(ikey, p.A, k, sm, res);
  if (t == 0) {  // :146 sum (descending order, sequential), int32 truncation, at least 1
    float s = 0.f;
    for (int j = 0; j < k; ++j) s += ord2f(~(unsigned)(res[j] >> 32));
    sdyn = max((int)s, 1);
  }
  __syncthreads();
  const int dk = min(sdyn, p.A);
  // dk smallest costs: key order = (value asc, index asc)  (:148); dynamic k <= candidate_k <= 16
  auto ckey = [&](int i) { return ((unsigned long long)f2ord(__ldg(crow + i) + 0.f) << 32) | (unsigned)i; };
  row_select_small(ckey, p.A, min(dk, kOtaMaxK), sm, res);
  if (t < min(dk, kOtaMaxK)) {  // :149 matching_matrix[gt_idx, anchor_idx] = 1
    const int a = (int)(unsigned)res[t];
    atomicAdd(p.cnt + a, 1);
    p.who[a] = g;
  }
}

__global__ void __launch_bounds__(256) ota_resolve_kernel(const OtaArgs p) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= p.A) return;
  const int c = p.cnt[a];
  int m = p.G;  // :160-161 the appended row of ones wins when nothing matched
  if (c == 1) {
    m = p.who[a];
  } else if (c > 1) {  // :154-158 argmin over ALL GTs of the cost column (first index)
    float best = __ldg(p.cost + a);
    m = 0;
    for (int g = 1; g < p.G; ++g) {
      const float v = __ldg(p.cost + (long long)g * p.ldc + a);
      if (v < best) {
        best = v;
        m = g;
      }
    }
  }
  p.out[a] = m;
}

}  // namespace bdet

using namespace bdet;

extern "C" size_t bdet_ota_topk_match_workspace(int A) { return A <= 0 ? 16 : (size_t)A * 8 + 256; }

extern "C" int bdet_ota_topk_match(const float* cost, int ldc, const float* ious, int ldi, int G, int A, int candidate_k,
                                   int* matched_gt, void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(G >= 0 && A >= 0 && ldc >= A && ldi >= A, "bad shape");
  BDET_REQUIRE(candidate_k >= 1 && candidate_k <= kOtaMaxK, "candidate_k must be in [1, 16]");
  if (A == 0) return BDET_OK;
  BDET_REQUIRE(matched_gt, "null output");
  const size_t need = bdet_ota_topk_match_workspace(A);
  if (!workspace || workspace_bytes < need) return set_error(BDET_EWORKSPACE, "bdet_ota_topk_match: workspace needs %zu bytes", need);
  BDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 3u) == 0, "workspace must be 4-byte aligned");
  BDET_REQUIRE(G == 0 || (cost && ious), "null argument");
  if (G > 65535) return set_error(BDET_EUNSUPPORTED, "bdet_ota_topk_match: G > 65535");
  OtaArgs a;
  a.cost = cost;
  a.ious = ious;
  a.G = G;
  a.A = A;
  a.ldc = ldc;
  a.ldi = ldi;
  a.k = candidate_k;
  a.cnt = reinterpret_cast<int*>(workspace);
  a.who = a.cnt + A;
  a.out = matched_gt;
  cudaStream_t st = as_stream(stream);
  BDET_CUDA(cudaMemsetAsync(a.cnt, 0, (size_t)A * 4, st));
  if (G > 0) BDET_KERNEL("ota_rows_kernel", st, ota_rows_kernel<<<G, kOtaThreads, 0, st>>>(a));
  BDET_KERNEL("ota_resolve_kernel", st, ota_resolve_kernel<<<ceil_div(A, 256), 256, 0, st>>>(a));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
