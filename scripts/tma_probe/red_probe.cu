// Probe: cost of fp32 global reductions (REDG) per lane vs per byte on B200: the same floats added as scalar red.add.f32
// (one float per lane), red.add.v2.f32 and red.add.v4.f32 (16 bytes per lane).  Pattern of the ROIAlign backward: a
// warp adds a footprint row of `cols` consecutive floats of one channel plane, rows W floats apart, planes H*W apart.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_probe red_probe.cu ; run: ./red_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void red_kernel(float* base, int W, int H, int planes, int rows, int cols, int iters) {
  // one warp per plane slice; lanes cover `cols` floats of a row in units of V floats
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int lanes_per_row = cols / V;            // cols is a multiple of 4
  const int rows_per_pass = 32 / lanes_per_row;  // several rows per warp instruction when the row is short
  const int lr = lane / lanes_per_row, lx = lane - lr * lanes_per_row;
  if (lr >= rows_per_pass) return;
  for (int it = 0; it < iters; ++it) {
    const int plane = (warp * 7 + it * 131) % planes;
    float* p = base + (size_t)plane * H * W + (size_t)((it * 37) % (H - rows)) * W + ((it * 12) % (W - cols) & ~3);
    for (int r = lr; r < rows; r += rows_per_pass) {
      float* q = p + (size_t)r * W + lx * V;
      const float v = 1.0f;
      if (V == 1) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q), "f"(v) : "memory");
      if (V == 2) asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(q), "f"(v) : "memory");
      if (V == 4) asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(q), "f"(v) : "memory");
    }
  }
}

int main() {
  const int W = 336, H = 200, planes = 4096;  // 1.1 GB of planes: the adds spread over HBM-resident lines
  float* d;
  cudaMalloc(&d, (size_t)planes * H * W * 4);
  cudaMemset(d, 0, (size_t)planes * H * W * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int ctas = 148 * 8, threads = 256, iters = 64, rows = 12;
  for (int cols : {8, 16, 32}) {
    for (int v : {1, 2, 4}) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (v == 1) red_kernel<1><<<ctas, threads>>>(d, W, H, planes, rows, cols, iters);
        if (v == 2) red_kernel<2><<<ctas, threads>>>(d, W, H, planes, rows, cols, iters);
        if (v == 4) red_kernel<4><<<ctas, threads>>>(d, W, H, planes, rows, cols, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      const double floats = (double)ctas * (threads / 32) * iters * rows * cols;
      printf("cols=%2d v%d: %.3f ms  %.1f Gfloat/s  %.2f ns per 1000 floats  err=%s\n", cols, v, best, floats / best / 1e6,
             best * 1e6 / (floats / 1000), cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
