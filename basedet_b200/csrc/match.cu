// a5: max-IoU Matcher on a materialised (G, A) matrix.
// Reference: basedet/layers/common/matcher.py:31-51.
//
// The reference streams the matrix >= 4 times (col max, argmax, row max, == mask, sum).  Here:
//   kernel 1 (one pass over the matrix, HBM-read bound, 4*G*A bytes):
//       lanes run along A (coalesced 128-byte row segments), every thread keeps a running
//       (max, first argmax) for its columns while walking the G rows; per row the warp folds its
//       values with one redux.sync and keeps a CTA-local row maximum in shared memory;
//       at the end the CTA publishes its row maxima (global atomicMax + a per-CTA copy).
//   kernel 2 (low-quality fix-up): one CTA per (image, row) re-reads a column segment only if that segment's
//       maximum equals the row maximum -- ~G segments in total instead of a second full pass.
// (R, G) layout of RCNN (layers/head/rcnn.py:113-116): one warp per row, warp-shuffle argmax.
#include "common.cuh"

namespace bdet {

constexpr int kMatchThreads = 256;

struct MatchPlan {
  int cpt;    // columns per thread
  int tiles;  // column tiles per image
};

static MatchPlan match_plan(int A, int B) {
  MatchPlan p;
  const long long want = (long long)sm_count() * 6;
  p.cpt = 4;
  while (p.cpt > 1 && (long long)ceil_div(A, kMatchThreads * p.cpt) * B < want) p.cpt >>= 1;
  p.tiles = ceil_div(A, kMatchThreads * p.cpt);
  return p;
}

struct MatchArgs {
  const float* m;
  long long bs, ld;
  const int* g_dev;
  int Gmax, A, tiles, allow_lq;
  int* idx;
  int* labels;
  uint32_t* rowmax;  // (B, Gmax)   order-encoded, zero-initialised
  uint32_t* blkmax;  // (B, Gmax, tiles)
  MatchCfg cfg;
};

template <int CPT>
__global__ void __launch_bounds__(kMatchThreads) match_colmax_kernel(const MatchArgs p) {
  extern __shared__ uint32_t srow[];
  const int b = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
  const int G = p.g_dev ? min(p.g_dev[b], p.Gmax) : p.Gmax;
  const long long col0 = (long long)tile * (kMatchThreads * CPT) + t;
  for (int g = t; g < G; g += kMatchThreads) srow[g] = 0u;
  __syncthreads();

  float best[CPT];
  int bidx[CPT];
  bool ok[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    best[j] = -CUDART_INF_F;
    bidx[j] = 0;
    ok[j] = col0 + j * kMatchThreads < p.A;
  }
  const float* base = p.m + b * p.bs + col0;
  constexpr int U = 4;
  for (int g0 = 0; g0 < G; g0 += U) {
    float v[U][CPT];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < CPT; ++j)
        v[u][j] = (g0 + u < G && ok[j]) ? __ldcs(base + (long long)(g0 + u) * p.ld + j * kMatchThreads) : -CUDART_INF_F;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (g0 + u < G) {  // uniform across the CTA
        float rv = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          if (v[u][j] > best[j]) {  // strict: first index wins ties
            best[j] = v[u][j];
            bidx[j] = g0 + u;
          }
          rv = fmaxf(rv, v[u][j]);
        }
        uint32_t w = __reduce_max_sync(0xffffffffu, f2ord(rv));
        if ((t & 31) == 0) atomicMax(&srow[g0 + u], w);
      }
    }
  }
  if (G == 0) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) best[j] = 0.f;  // documented extension: no GT -> IoU 0
  }
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    if (ok[j]) {
      long long o = (long long)b * p.A + col0 + j * kMatchThreads;
      p.idx[o] = bidx[j];
      p.labels[o] = threshold_label(p.cfg, best[j]);
    }
  }
  if (p.allow_lq) {
    __syncthreads();
    uint32_t* bm = p.blkmax + (long long)b * p.Gmax * p.tiles + tile;  // (B, Gmax, tiles): row-contiguous for kernel 2
    for (int g = t; g < G; g += kMatchThreads) {
      uint32_t u = srow[g];
      bm[(long long)g * p.tiles] = u;
      atomicMax(&p.rowmax[(long long)b * p.Gmax + g], u);
    }
  }
}

// Low-quality fix-up, matcher.py:47-49: labels[(matrix == rowmax).sum(0) > 0] = 1.
// One CTA per (image, row): finds the column tiles whose maximum equals the row maximum (float compare, so
// -0 == +0 as in the reference) and re-reads only those segments of the row.
constexpr int kLqThreads = 128;
template <int CPT>
__global__ void __launch_bounds__(kLqThreads) match_lq_kernel(const MatchArgs p) {
  extern __shared__ int stiles[];
  __shared__ int nhit;
  const int g = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const int G = p.g_dev ? min(p.g_dev[b], p.Gmax) : p.Gmax;
  if (g >= G) return;
  if (t == 0) nhit = 0;
  __syncthreads();
  const float target = ord2f(p.rowmax[(long long)b * p.Gmax + g]);
  const uint32_t* bm = p.blkmax + ((long long)b * p.Gmax + g) * p.tiles;
  for (int i = t; i < p.tiles; i += kLqThreads)
    if (ord2f(bm[i]) == target) stiles[atomicAdd(&nhit, 1)] = i;
  __syncthreads();
  const int nh = nhit;
  const float* row = p.m + b * p.bs + (long long)g * p.ld;
  constexpr int kCols = kMatchThreads * CPT;
  for (int h = 0; h < nh; ++h) {
    const long long base = (long long)stiles[h] * kCols;
    for (int i = t; i < kCols; i += kLqThreads) {
      const long long c = base + i;
      if (c < p.A && __ldg(row + c) == target) p.labels[(long long)b * p.A + c] = 1;
    }
  }
}

// One warp per row of an (R, G) matrix: lanes stride over G, then a 5-step shuffle argmax
// (larger value wins, lower index wins ties == first argmax).
__global__ void __launch_bounds__(256) match_rows_kernel(const float* __restrict__ m, int R, int G, float* __restrict__ mx, int* __restrict__ am) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* p = m + (long long)row * G;
  float best = -CUDART_INF_F;
  int bi = 0x7fffffff;
  for (int g = lane; g < G; g += 32) {
    float v = __ldg(p + g);
    if (v > best || bi == 0x7fffffff) {  // first element seeds (handles all -inf rows)
      best = v;
      bi = g;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, s);
    int oi = __shfl_xor_sync(0xffffffffu, bi, s);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if (lane == 0) {
    if (mx) mx[row] = best;
    am[row] = (bi == 0x7fffffff) ? 0 : bi;
  }
}

template <int CPT>
static void launch_match(const MatchArgs& a, int B, cudaStream_t st) {
  dim3 grid(a.tiles, B);
  size_t smem = (size_t)max(a.Gmax, 1) * 4;
  BDET_KERNEL("match_colmax_kernel", st, match_colmax_kernel<CPT><<<grid, kMatchThreads, smem, st>>>(a));
  if (a.allow_lq)
    BDET_KERNEL("match_lq_kernel", st, match_lq_kernel<CPT><<<dim3(a.Gmax, B), kLqThreads, (size_t)a.tiles * 4, st>>>(a));
}

}  // namespace bdet

using namespace bdet;

extern "C" size_t bdet_match_workspace(int Gmax, int A, int B) {
  if (Gmax <= 0 || A <= 0 || B <= 0) return 16;
  MatchPlan pl = match_plan(A, B);
  return align_up((size_t)B * Gmax * 4, 256) + (size_t)B * pl.tiles * Gmax * 4 + 256;
}

extern "C" int bdet_match(const float* matrix, int64_t ld, int64_t batch_stride, const int* g_dev, int Gmax, int A, int B,
                          const float* thresholds_host, const int* labels_host, int n_labels, int allow_low_quality,
                          int* match_idx, int* labels, void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(Gmax >= 0 && A >= 0 && B >= 0, "negative size");
  MatchArgs a;
  int rc = make_match_cfg(&a.cfg, thresholds_host, labels_host, n_labels);
  if (rc) return rc;
  if (A == 0 || B == 0) return BDET_OK;
  BDET_REQUIRE(match_idx && labels, "null output");
  BDET_REQUIRE(Gmax == 0 || matrix, "null matrix");
  BDET_REQUIRE(ld >= A, "row stride smaller than A");
  BDET_REQUIRE(Gmax <= 12000, "Gmax too large for the shared-memory row table");
  BDET_REQUIRE(A <= 10000 * kMatchThreads, "A too large for the tile table");
  MatchPlan pl = match_plan(A, B);
  a.m = matrix;
  a.bs = batch_stride;
  a.ld = ld;
  a.g_dev = g_dev;
  a.Gmax = Gmax;
  a.A = A;
  a.tiles = pl.tiles;
  a.allow_lq = (allow_low_quality != 0 && Gmax > 0) ? 1 : 0;
  a.idx = match_idx;
  a.labels = labels;
  a.rowmax = nullptr;
  a.blkmax = nullptr;
  cudaStream_t st = as_stream(stream);
  if (a.allow_lq) {
    if (!workspace || workspace_bytes < bdet_match_workspace(Gmax, A, B))
      return set_error(BDET_EWORKSPACE, "bdet_match: workspace needs %zu bytes", bdet_match_workspace(Gmax, A, B));
    BDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 3u) == 0, "workspace must be 4-byte aligned");
    a.rowmax = reinterpret_cast<uint32_t*>(workspace);
    a.blkmax = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(workspace) + align_up((size_t)B * Gmax * 4, 256));
    BDET_CUDA(cudaMemsetAsync(a.rowmax, 0, (size_t)B * Gmax * 4, st));
  }
  if (B > 65535) return set_error(BDET_EUNSUPPORTED, "bdet_match: B > 65535");
  switch (pl.cpt) {
    case 4: launch_match<4>(a, B, st); break;
    case 2: launch_match<2>(a, B, st); break;
    default: launch_match<1>(a, B, st); break;
  }
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_match_rows(const float* matrix, int R, int G, float* max_out, int* argmax_out, bdet_stream_t stream) {
  BDET_REQUIRE(R >= 0 && G >= 1, "need G >= 1");
  if (R == 0) return BDET_OK;
  BDET_REQUIRE(matrix && argmax_out, "null argument");
  BDET_KERNEL("match_rows_kernel", as_stream(stream), match_rows_kernel<<<ceil_div(R, 8), 256, 0, as_stream(stream)>>>(matrix, R, G, max_out, argmax_out));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
