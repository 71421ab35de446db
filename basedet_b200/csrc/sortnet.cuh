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

// Sort a[0..P) ascending; P is a power of two; all threads of the CTA participate.
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* a, int P) {
  const int t = threadIdx.x, nt = blockDim.x;
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = t; i < (P >> 1); i += nt) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        uint64_t x = a[lo], y = a[hi];
        if ((x > y) == up) {
          a[lo] = y;
          a[hi] = x;
        }
      }
    }
  }
  __syncthreads();
}

// The strides < tile part of one bitonic merge level `size` on a tile held in shared memory.
// `base` = global index of a[0] (decides the sort direction of each pair).
__device__ __forceinline__ void bitonic_merge_tail_smem(uint64_t* a, int tile, long long base, long long size, int first_stride) {
  const int t = threadIdx.x, nt = blockDim.x;
  for (int stride = first_stride; stride > 0; stride >>= 1) {
    __syncthreads();
    for (int i = t; i < (tile >> 1); i += nt) {
      int lo = 2 * i - (i & (stride - 1));
      int hi = lo + stride;
      bool up = (((base + lo) & size) == 0);
      uint64_t x = a[lo], y = a[hi];
      if ((x > y) == up) {
        a[lo] = y;
        a[hi] = x;
      }
    }
  }
  __syncthreads();
}

static inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace bdet
