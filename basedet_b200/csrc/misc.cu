// Small Boxes / glue ops of the drop-in layer: width/height/area, BoxConverter, cond_take, label census.
// Reference: basedet/structures/boxes.py:36-52, basedet/structures/box_convert.py:51-82,
//            basedet/layers/common/function.py:19-23 (non_zeros -> F.cond_take).
#include "common.cuh"

namespace bdet {

__global__ void __launch_bounds__(256) box_props_kernel(const float* __restrict__ b, int ld, int N, int mode, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* p = b + (long long)i * ld;
  float w = __ldg(p + 2) - __ldg(p), h = __ldg(p + 3) - __ldg(p + 1);
  out[i] = mode == 0 ? w : (mode == 1 ? h : w * h);
}

// box_convert.py:58-82: everything goes through XYWH.
__global__ void __launch_bounds__(256) box_convert_kernel(const float4* __restrict__ in, int N, int from, int to, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float4 b = __ldg(in + i);
  if (from == 0) {  // XYXY -> XYWH
    b.z = b.z - b.x;
    b.w = b.w - b.y;
  } else if (from == 2) {  // XcYcWH -> XYWH
    b.x = b.x - __fdiv_rn(b.z, 2.f);
    b.y = b.y - __fdiv_rn(b.w, 2.f);
  }
  if (to == 0) {  // XYWH -> XYXY
    b.z = b.x + b.z;
    b.w = b.y + b.w;
  } else if (to == 2) {  // XYWH -> XcYcWH
    b.x = b.x + __fdiv_rn(b.z, 2.f);
    b.y = b.y + __fdiv_rn(b.w, 2.f);
  }
  out[i] = b;
}

// ---- ordered stream compaction: per-CTA counts -> single-CTA exclusive scan -> scatter ------------------
constexpr int kCtThreads = 256;
constexpr int kCtItems = 8;
constexpr int kCtTile = kCtThreads * kCtItems;

__device__ __forceinline__ bool ct_pred(const float* x, const uint8_t* mask, long long i) {
  return mask ? (mask[i] != 0) : (__ldg(x + i) != 0.f);
}

__global__ void __launch_bounds__(kCtThreads) cond_take_count_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, long long n,
                                                                     int* __restrict__ tile_count) {
  __shared__ int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  long long base = (long long)blockIdx.x * kCtTile;
  int c = 0;
#pragma unroll
  for (int j = 0; j < kCtItems; ++j) {
    long long i = base + j * kCtThreads + threadIdx.x;
    c += (i < n && ct_pred(x, mask, i)) ? 1 : 0;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) tile_count[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) cond_take_scan_kernel(int* tile_count, int tiles, int* count_dev) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < tiles; base += 1024) {
    int i = base + t;
    int v = i < tiles ? tile_count[i] : 0;
    int incl = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= s) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int add = carry;
    for (int w = 0; w < warp; ++w) add += warp_tot[w];
    if (i < tiles) tile_count[i] = add + incl - v;  // exclusive prefix
    __syncthreads();
    if (t == 1023) carry = add + incl;
    __syncthreads();
  }
  if (t == 0) *count_dev = carry;
}

__global__ void __launch_bounds__(kCtThreads) cond_take_scatter_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, long long n,
                                                                       const int* __restrict__ tile_off, float* __restrict__ out_vals,
                                                                       int* __restrict__ out_idx) {
  __shared__ int warp_cnt[kCtItems][kCtThreads / 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long base = (long long)blockIdx.x * kCtTile;
  bool pr[kCtItems];
  uint32_t bal[kCtItems];
#pragma unroll
  for (int j = 0; j < kCtItems; ++j) {
    long long i = base + j * kCtThreads + t;
    pr[j] = i < n && ct_pred(x, mask, i);
    bal[j] = __ballot_sync(0xffffffffu, pr[j]);
    if (lane == 0) warp_cnt[j][warp] = __popc(bal[j]);
  }
  __syncthreads();
  int run = tile_off[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kCtItems; ++j) {
    int before = 0;
    for (int w = 0; w < kCtThreads / 32; ++w) {
      int c = warp_cnt[j][w];
      if (w < warp) before += c;
    }
    int row_total = 0;
    for (int w = 0; w < kCtThreads / 32; ++w) row_total += warp_cnt[j][w];
    if (pr[j]) {
      long long i = base + j * kCtThreads + t;
      int pos = run + before + __popc(bal[j] & ((1u << lane) - 1u));
      out_idx[pos] = (int)i;
      if (out_vals) out_vals[pos] = __ldg(x + i);
    }
    run += row_total;
  }
}

// Few CTAs per image and ONE atomic triple per CTA: thousands of same-address atomics serialise in L2 and used to
// cost more than reading the labels.
__global__ void __launch_bounds__(256) count_labels_kernel(const int* __restrict__ labels, int A, int* __restrict__ counts) {
  __shared__ int sred[3][8];
  const int b = blockIdx.y, t = threadIdx.x;
  const int* row = labels + (long long)b * A;
  int neg = 0, zero = 0, pos = 0;
  auto tally = [&](int l) {
    neg += l < 0;
    zero += l == 0;
    pos += l > 0;
  };
  const int head = min(A, (int)((4 - (((uintptr_t)row >> 2) & 3)) & 3));  // elements before the first 16-byte boundary
  const int nvec = (A - head) >> 2;
  const int4* v = reinterpret_cast<const int4*>(row + head);
  for (int i = blockIdx.x * 256 + t; i < nvec; i += gridDim.x * 256) {
    const int4 q = __ldg(v + i);
    tally(q.x);
    tally(q.y);
    tally(q.z);
    tally(q.w);
  }
  if (blockIdx.x == 0) {
    if (t < head) tally(__ldg(row + t));
    const int tail0 = head + nvec * 4;
    if (tail0 + t < A) tally(__ldg(row + tail0 + t));  // < 4 elements
  }
  neg = __reduce_add_sync(0xffffffffu, neg);
  zero = __reduce_add_sync(0xffffffffu, zero);
  pos = __reduce_add_sync(0xffffffffu, pos);
  if ((t & 31) == 0) {
    sred[0][t >> 5] = neg;
    sred[1][t >> 5] = zero;
    sred[2][t >> 5] = pos;
  }
  __syncthreads();
  if (t < 3) {
    int sum = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += sred[t][w];
    if (sum) atomicAdd(counts + b * 3 + t, sum);
  }
}

}  // namespace bdet

using namespace bdet;

extern "C" int bdet_box_props(const float* boxes, int ld, int N, int mode, float* out, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0 && ld >= 4 && mode >= 0 && mode <= 2, "bad arguments");
  if (N == 0) return BDET_OK;
  BDET_REQUIRE(boxes && out, "null argument");
  BDET_KERNEL("box_props_kernel", as_stream(stream), box_props_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(boxes, ld, N, mode, out));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_box_convert(const float* boxes, int N, int from_mode, int to_mode, float* out, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0 && from_mode >= 0 && from_mode <= 2 && to_mode >= 0 && to_mode <= 2, "bad arguments");
  if (N == 0) return BDET_OK;
  BDET_REQUIRE(boxes && out && aligned16(boxes) && aligned16(out), "boxes/out must be 16-byte aligned device pointers");
  BDET_KERNEL("box_convert_kernel", as_stream(stream), box_convert_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(boxes), N, from_mode, to_mode,
                                                                      reinterpret_cast<float4*>(out)));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" size_t bdet_cond_take_workspace(int64_t n) { return (size_t)(ceil_div(n < 1 ? 1 : n, kCtTile) + 1) * 4 + 256; }

extern "C" int bdet_cond_take(const float* x, const uint8_t* mask, int64_t n, float* out_vals, int* out_idx, int* count_dev,
                              void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(n >= 0 && n <= 0x7fffffffLL, "n out of range");
  BDET_REQUIRE(count_dev, "null count");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    BDET_CUDA(cudaMemsetAsync(count_dev, 0, 4, st));
    return BDET_OK;
  }
  BDET_REQUIRE((x || mask) && out_idx, "null argument");
  BDET_REQUIRE(!out_vals || x, "values requested without x");
  if (!workspace || workspace_bytes < bdet_cond_take_workspace(n))
    return set_error(BDET_EWORKSPACE, "bdet_cond_take: workspace needs %zu bytes", bdet_cond_take_workspace(n));
  int* tile_count = reinterpret_cast<int*>(workspace);
  const int tiles = ceil_div(n, kCtTile);
  BDET_KERNEL("cond_take_count_kernel", st, cond_take_count_kernel<<<tiles, kCtThreads, 0, st>>>(x, mask, n, tile_count));
  BDET_KERNEL("cond_take_scan_kernel", st, cond_take_scan_kernel<<<1, 1024, 0, st>>>(tile_count, tiles, count_dev));
  BDET_KERNEL("cond_take_scatter_kernel", st, cond_take_scatter_kernel<<<tiles, kCtThreads, 0, st>>>(x, mask, n, tile_count, out_vals, out_idx));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_count_labels(const int* labels, int A, int B, int* counts, bdet_stream_t stream) {
  BDET_REQUIRE(A >= 0 && B >= 0 && B <= 65535, "bad shape");
  if (B == 0) return BDET_OK;
  BDET_REQUIRE(counts, "null counts");
  cudaStream_t st = as_stream(stream);
  BDET_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * 3 * 4, st));
  if (A == 0) return BDET_OK;
  BDET_REQUIRE(labels, "null labels");
  BDET_KERNEL("count_labels_kernel", st, count_labels_kernel<<<dim3(min(ceil_div(A, 8192), 16), B), 256, 0, st>>>(labels, A, counts));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
