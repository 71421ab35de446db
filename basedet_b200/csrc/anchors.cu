// a1/a2: anchor boxes and anchor points for all FPN levels in ONE launch.
// Reference: basedet/layers/common/anchor_generator.py:23-30 (create_anchor_grid),
// :111-122 (DefaultAnchorGenerator), :152-165 (AnchorPointGenerator), :175-182 (FastPointGenerator).
// Pure write kernel: one thread per output box (128-bit store) / point (64-bit store).
#include "anchor_levels.cuh"

namespace bdet {

__global__ void __launch_bounds__(256) anchors_grid_kernel(float4* __restrict__ out, const __grid_constant__ AnchorLevels p) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= p.start[p.n_levels]) return;
  long long slot;
  const float4 v = anchor_at(p, i, &slot);
  out[slot] = v;
}

// mode 0: AnchorPointGenerator -- (x, y) for (h, w) row-major, each repeated num_anchors times.
// mode 1: FastPointGenerator   -- row j*H + i holds (i*stride, j*stride), i < H, j < W (reference quirk).
__global__ void __launch_bounds__(256) points_grid_kernel(float2* __restrict__ out, const __grid_constant__ AnchorLevels p, int mode) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= p.start[p.n_levels]) return;
  int l = 0;
#pragma unroll
  for (int k = 1; k < BDET_MAX_LEVELS; ++k)
    if (k < p.n_levels && i >= p.start[k]) l = k;
  long long r = i - p.start[l];
  float2 v;
  if (mode == 0) {
    long long pos = r / p.n_base[l];
    int w = (int)(pos % p.W[l]);
    int h = (int)(pos / p.W[l]);
    v.x = (float)(p.shift[l] + (double)w * p.stride[l]);
    v.y = (float)(p.shift[l] + (double)h * p.stride[l]);
  } else {
    int ii = (int)(r % p.H[l]);
    int jj = (int)(r / p.H[l]);
    float s = (float)p.stride[l];
    v.x = (float)ii * s;
    v.y = (float)jj * s;
  }
  out[p.out_off[l] + r] = v;
}

}  // namespace bdet

using namespace bdet;

extern "C" int bdet_anchors_grid(float* out, int n_levels, const int* hw_host, const double* stride_host,
                                 const double* shift_host, const int* n_base_host, const float* base_host,
                                 const int64_t* out_offset_host, bdet_stream_t stream) {
  BDET_REQUIRE(out && hw_host && stride_host && n_base_host && base_host && out_offset_host, "null argument");
  BDET_REQUIRE(aligned16(out), "out must be 16-byte aligned");
  AnchorLevels p;
  int rc = fill_levels(&p, n_levels, hw_host, stride_host, shift_host, n_base_host, 0, base_host, out_offset_host);
  if (rc) return rc;
  long long total = p.start[n_levels];
  if (total == 0) return BDET_OK;
  BDET_KERNEL("anchors_grid_kernel", as_stream(stream), anchors_grid_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(out), p));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_points_grid(float* out, int n_levels, const int* hw_host, const double* stride_host,
                                const double* shift_host, int num_anchors, int mode,
                                const int64_t* out_offset_host, bdet_stream_t stream) {
  BDET_REQUIRE(out && hw_host && stride_host && out_offset_host, "null argument");
  BDET_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (AnchorPoint) or 1 (FastPoint)");
  BDET_REQUIRE((reinterpret_cast<uintptr_t>(out) & 7u) == 0, "out must be 8-byte aligned");
  AnchorLevels p;
  int rc = fill_levels(&p, n_levels, hw_host, stride_host, shift_host, nullptr, mode == 0 ? num_anchors : 1,
                       nullptr, out_offset_host);
  if (rc) return rc;
  long long total = p.start[n_levels];
  if (total == 0) return BDET_OK;
  BDET_KERNEL("points_grid_kernel", as_stream(stream), points_grid_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<float2*>(out), p, mode));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
