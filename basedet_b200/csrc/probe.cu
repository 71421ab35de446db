// Measurement aid (bench / profiles only): plain streaming kernels that establish the write-only, read-only and
// copy HBM ceilings of the device the box-op kernels are compared against.  Not part of the product path.
#include "common.cuh"

namespace bdet {

__global__ void __launch_bounds__(256) probe_write_kernel(float4* __restrict__ dst, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (; i + 3 * stride < n4; i += 4 * stride) {
    dst[i] = v;
    dst[i + stride] = v;
    dst[i + 2 * stride] = v;
    dst[i + 3 * stride] = v;
  }
  for (; i < n4; i += stride) dst[i] = v;
}

__global__ void __launch_bounds__(256) probe_read_kernel(const float4* __restrict__ src, size_t n4, float* __restrict__ sink) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    acc += (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w) + (c.x + c.y + c.z + c.w) + (d.x + d.y + d.z + d.w);
  }
  for (; i < n4; i += stride) {
    float4 a = __ldcs(src + i);
    acc += a.x + a.y + a.z + a.w;
  }
  if (acc == 123456.789f) *sink = acc;  // keeps the loads alive
}

__global__ void __launch_bounds__(256) probe_copy_kernel(float4* __restrict__ dst, const float4* __restrict__ src, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    dst[i] = a;
    dst[i + stride] = b;
    dst[i + 2 * stride] = c;
    dst[i + 3 * stride] = d;
  }
  for (; i < n4; i += stride) dst[i] = __ldcs(src + i);
}

}  // namespace bdet

using namespace bdet;

extern "C" int bdet_bw_probe(void* dst, const void* src, size_t bytes, int mode, int ctas_per_sm, bdet_stream_t stream) {
  BDET_REQUIRE(bytes % 16 == 0 && mode >= 0 && mode <= 3, "bytes must be a multiple of 16; mode in [0, 3]");
  BDET_REQUIRE(aligned16(dst) && (mode == 0 || mode == 3 || aligned16(src)), "pointers must be 16-byte aligned");
  const size_t n4 = bytes / 16;
  const int grid = sm_count() * (ctas_per_sm > 0 ? ctas_per_sm : 8);
  cudaStream_t st = as_stream(stream);
  if (mode == 0) {
    BDET_KERNEL("probe_write_kernel", st, probe_write_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<float4*>(dst), n4));
  } else if (mode == 1) {
    BDET_KERNEL("probe_read_kernel", st, probe_read_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(src), n4, reinterpret_cast<float*>(dst)));
  } else if (mode == 2) {
    BDET_KERNEL("probe_copy_kernel", st, probe_copy_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src), n4));
  } else {
    BDET_CUDA(cudaMemsetAsync(dst, 0, bytes, st));
  }
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
