// Shared-memory bitonic network and 64-bit (score, index) keys shared by top-k and NMS.
#pragma once
#include "common.cuh"

namespace bdet {

// Unique, totally ordered key: ascending key order == (score descending, index ascending).
__device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) {
  float f = score + 0.f;  // -0 -> +0 so that equal scores give equal high words
  return ((uint64_t)(~f2ord(f)) << 32) | idx;
}
__device__ __forceinline__ float key_score(uint64_t key) { return ord2f(~(uint32_t)(key >> 32)); }

// One compare-exchange step of the network for pair index i (stride, direction bit `size`, global offset `base`).
__device__ __forceinline__ void bitonic_cx(uint64_t* a, int i, int stride, long long base, long long size) {
  const int lo = 2 * i - (i & (stride - 1));
  const int hi = lo + stride;
  const bool up = (((base + lo) & size) == 0);
  const uint64_t x = a[lo], y = a[hi];
  if ((x > y) == up) {
    a[lo] = y;
    a[hi] = x;
  }
}

// All strides first_stride, first_stride/2, ..., 1 of merge level `size` on `n` elements held in shared memory.
// ITERS = pairs per thread (compile-time, so the independent shared-memory accesses of a stage overlap instead of
// running as a latency chain); 0 = generic loop.
template <int ITERS>
__device__ __forceinline__ void bitonic_strides(uint64_t* a, int n, long long base, long long size, int first_stride) {
  const int t = threadIdx.x, nt = blockDim.x;
  for (int stride = first_stride; stride > 0; stride >>= 1) {
    __syncthreads();
    if constexpr (ITERS > 0) {
      // all loads of the thread's (disjoint) pairs first, then the stores: the compiler cannot prove the pairs
      // disjoint, so left to itself it serialises load -> store -> load
      uint64_t x[ITERS], y[ITERS];
      int lo[ITERS];
#pragma unroll
      for (int u = 0; u < ITERS; ++u) {
        const int i = t + u * nt;
        lo[u] = 2 * i - (i & (stride - 1));
        x[u] = a[lo[u]];
        y[u] = a[lo[u] + stride];
      }
#pragma unroll
      for (int u = 0; u < ITERS; ++u) {
        const bool up = (((base + lo[u]) & size) == 0);
        if ((x[u] > y[u]) == up) {
          a[lo[u]] = y[u];
          a[lo[u] + stride] = x[u];
        }
      }
    } else {
      for (int i = t; i < (n >> 1); i += nt) bitonic_cx(a, i, stride, base, size);
    }
  }
}

__device__ __forceinline__ void bitonic_strides_any(uint64_t* a, int n, long long base, long long size, int first_stride) {
  const int pairs = n >> 1, nt = blockDim.x;
  if (pairs == nt) bitonic_strides<1>(a, n, base, size, first_stride);
  else if (pairs == 2 * nt) bitonic_strides<2>(a, n, base, size, first_stride);
  else if (pairs == 4 * nt) bitonic_strides<4>(a, n, base, size, first_stride);
  else if (pairs == 8 * nt) bitonic_strides<8>(a, n, base, size, first_stride);
  else bitonic_strides<0>(a, n, base, size, first_stride);
}

// Sort a[0..P) ascending; P is a power of two; all threads of the CTA participate.
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* a, int P) {
  for (int size = 2; size <= P; size <<= 1) bitonic_strides_any(a, P, 0, size, size >> 1);
  __syncthreads();
}

// The strides < tile part of one bitonic merge level `size` on a tile held in shared memory.
// `base` = global index of a[0] (decides the sort direction of each pair).
__device__ __forceinline__ void bitonic_merge_tail_smem(uint64_t* a, int tile, long long base, long long size, int first_stride) {
  bitonic_strides_any(a, tile, base, size, first_stride);
  __syncthreads();
}

static inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace bdet
