// a3/a4: pairwise box ops -- IoU / IoA / intersection / GIoU matrices, box centers, point distance.
// Reference: basedet/structures/op_patch.py:33-97 (IOU), :169-227 (IOA), :100-166 (center, distance),
//            basedet/structures/boxes.py:74-95 (giou), :114-130 (intersection).
//
// Kernel shape (HBM-write bound: 4*N*M bytes out vs 16*(N+M) in):
//   * a CTA owns kCols consecutive columns (boxes2) x a chunk of rows (boxes1);
//   * boxes2 are read once with 128-bit loads and kept in registers (box + area);
//   * the row chunk of boxes1 (+ area) is staged in shared memory and broadcast to all lanes;
//   * the CTA reduces the bounding box of its columns; rows that cannot intersect it are written
//     as zeros without evaluating any pair (IoU/IoA/INTER are exactly +0 there);
//   * each warp store covers 32 consecutive floats of one output row.
#include "common.cuh"

namespace bdet {

constexpr int kPairThreads = 256;

template <int MODE>
__device__ __forceinline__ float pair_value(float4 a, float aa, float4 b, float ab) {
  if (MODE == BDET_PAIR_IOU) return iou_pair(a, aa, b, ab);
  float iw = fminf(a.z, b.z) - fmaxf(a.x, b.x);
  float ih = fminf(a.w, b.w) - fmaxf(a.y, b.y);
  float inter = fmaxf(iw, 0.f) * fmaxf(ih, 0.f);
  if (MODE == BDET_PAIR_INTER) return inter;
  if (MODE == BDET_PAIR_IOA) {
    // op_patch.py:199-204: max(inter / area2, 0); inter == 0 gives +0 (or NaN -> 0)
    return inter > 0.f ? fmaxf(__fdiv_rn(inter, ab), 0.f) : 0.f;
  }
  // GIoU, boxes.py:83-95 (no clamp of iou; hull via F.clip(rb - lt, lower=0))
  float uni = (aa + ab) - inter;
  float iou = __fdiv_rn(inter, uni);
  float hw = fmaxf(fmaxf(a.z, b.z) - fminf(a.x, b.x), 0.f);
  float hh = fmaxf(fmaxf(a.w, b.w) - fminf(a.y, b.y), 0.f);
  float hull = hw * hh;
  return iou - __fdiv_rn(hull - uni, hull);
}

struct PairArgs {
  const float* b1;
  const float* b2;
  float* out;
  const int* n1;  // per-batch valid rows (device) or nullptr
  long long bs1, bs2, bs_out;
  int ld1, ld2, N, M, rows_per_cta;
  long long ldo;  // output row stride in floats (>= M)
};

// VECO: the thread owns 4 CONSECUTIVE columns and writes one 128-bit store per row (needs ldo % 4 == 0 and a
// 16-byte aligned output; scalar 4-byte stores top out near 3.5 TB/s on B200, 128-bit stores reach the copy peak).
template <int MODE, int CPT, bool VEC2, bool VECO>
__global__ void __launch_bounds__(kPairThreads) pairwise_kernel(const PairArgs p) {
  extern __shared__ float4 smem4[];
  const int b = blockIdx.z;
  const int n_rows_total = p.n1 ? min(p.n1[b], p.N) : p.N;
  const int row0 = blockIdx.y * p.rows_per_cta;
  if (row0 >= n_rows_total) return;
  const int rows = min(p.rows_per_cta, n_rows_total - row0);
  float4* sbox = smem4;                                         // rows_per_cta
  float* sarea = reinterpret_cast<float*>(smem4 + p.rows_per_cta);  // rows_per_cta
  __shared__ float sbb[4];
  __shared__ uint32_t sred[4];

  const float* b1 = p.b1 + b * p.bs1;
  const float* b2 = p.b2 + b * p.bs2;
  float* out = p.out + b * p.bs_out;
  const int t = threadIdx.x;
  const long long col0 = (long long)blockIdx.x * (kPairThreads * CPT);

  if (t < 4) sred[t] = (t < 2) ? 0xffffffffu : 0u;
  for (int r = t; r < rows; r += kPairThreads) {
    float4 a = load_box<false>(b1, row0 + r, p.ld1);
    sbox[r] = a;
    sarea[r] = box_area(a);
  }

  float4 bx[CPT];
  float ar[CPT];
  bool ok[CPT];
  float mnx = CUDART_INF_F, mny = CUDART_INF_F, mxx = -CUDART_INF_F, mxy = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    long long c = VECO ? (col0 + 4 * t + j) : (col0 + j * kPairThreads + t);
    ok[j] = c < p.M;
    bx[j] = ok[j] ? load_box<VEC2>(b2, c, p.ld2) : make_float4(0.f, 0.f, 0.f, 0.f);
    ar[j] = box_area(bx[j]);
    if (ok[j]) {
      mnx = fminf(mnx, bx[j].x);
      mny = fminf(mny, bx[j].y);
      mxx = fmaxf(mxx, bx[j].z);
      mxy = fmaxf(mxy, bx[j].w);
    }
  }
  __syncthreads();
  if (MODE != BDET_PAIR_GIOU) {
    // CTA bounding box of the columns (NaN coordinates are ignored by fmin/fmax; such columns are 0 anyway)
    uint32_t r0 = __reduce_min_sync(0xffffffffu, f2ord(mnx));
    uint32_t r1 = __reduce_min_sync(0xffffffffu, f2ord(mny));
    uint32_t r2 = __reduce_max_sync(0xffffffffu, f2ord(mxx));
    uint32_t r3 = __reduce_max_sync(0xffffffffu, f2ord(mxy));
    if ((t & 31) == 0) {
      atomicMin(&sred[0], r0);
      atomicMin(&sred[1], r1);
      atomicMax(&sred[2], r2);
      atomicMax(&sred[3], r3);
    }
    __syncthreads();
    if (t < 4) sbb[t] = ord2f(sred[t]);
    __syncthreads();
  }
  const float bb0 = sbb[0], bb1 = sbb[1], bb2 = sbb[2], bb3 = sbb[3];

  float* orow = out + (long long)row0 * p.ldo + col0 + (VECO ? 4 * t : t);
#pragma unroll 2
  for (int r = 0; r < rows; ++r, orow += p.ldo) {
    const float4 a = sbox[r];
    float v[CPT];
    bool live = true;
    if (MODE != BDET_PAIR_GIOU)
      live = !(a.z <= bb0 || a.x >= bb2 || a.w <= bb1 || a.y >= bb3);  // NaN row -> live (computed)
    if (live) {
      const float aa = sarea[r];
#pragma unroll
      for (int j = 0; j < CPT; ++j) v[j] = pair_value<MODE>(a, aa, bx[j], ar[j]);
    } else {
#pragma unroll
      for (int j = 0; j < CPT; ++j) v[j] = 0.f;
    }
    if (VECO) {
      // columns >= M of the last quad land in the row padding (ldo is a multiple of 4)
      if (ok[0]) *reinterpret_cast<float4*>(orow) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < CPT; ++j)
        if (ok[j]) orow[j * kPairThreads] = v[j];
    }
  }
}

template <int MODE>
static int launch_pairwise(const PairArgs& a0, int B, cudaStream_t st) {
  PairArgs a = a0;
  constexpr int CPT = 4;
  const int col_tiles = ceil_div(a.M, kPairThreads * CPT);
  // enough CTAs to fill the machine a few times over, but long row loops to amortise the column loads
  const int want = sm_count() * 8;
  int row_tiles = max(1, min(ceil_div(a.N, 8), ceil_div(want, max(1, col_tiles * B))));
  a.rows_per_cta = ceil_div(a.N, row_tiles);
  if (a.rows_per_cta > 1024) a.rows_per_cta = 1024;
  row_tiles = ceil_div(a.N, a.rows_per_cta);
  if (row_tiles > 65535 || B > 65535) return set_error(BDET_EUNSUPPORTED, "pairwise: too many row tiles / batches");
  dim3 grid(col_tiles, row_tiles, B);
  size_t smem = (size_t)a.rows_per_cta * 20;
  const bool vec2 = (a.ld2 == 4) && aligned16(a.b2) && (a.bs2 % 4 == 0);
  const bool veco = (a.ldo % 4 == 0) && aligned16(a.out) && (a.bs_out % 4 == 0) && a.ldo >= (long long)(a.M + 3) / 4 * 4;
  static_assert(CPT == 4, "the 128-bit store path writes CPT == 4 columns");
  if (vec2 && veco)
    BDET_KERNEL("pairwise_kernel", st, pairwise_kernel<MODE, CPT, true, true><<<grid, kPairThreads, smem, st>>>(a));
  else if (vec2)
    BDET_KERNEL("pairwise_kernel", st, pairwise_kernel<MODE, CPT, true, false><<<grid, kPairThreads, smem, st>>>(a));
  else if (veco)
    BDET_KERNEL("pairwise_kernel", st, pairwise_kernel<MODE, CPT, false, true><<<grid, kPairThreads, smem, st>>>(a));
  else
    BDET_KERNEL("pairwise_kernel", st, pairwise_kernel<MODE, CPT, false, false><<<grid, kPairThreads, smem, st>>>(a));
  return BDET_OK;
}

__global__ void __launch_bounds__(256) box_center_kernel(const float* __restrict__ b, int ld, int N, float2* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* p = b + (long long)i * ld;
  // op_patch.py:106-109: (tl + br) / 2
  out[i] = make_float2(__fdiv_rn(p[0] + p[2], 2.f), __fdiv_rn(p[1] + p[3], 2.f));
}

__global__ void __launch_bounds__(256) point_distance_kernel(const float2* __restrict__ p1, int N, const float2* __restrict__ p2, int M,
                                                             float* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)N * M) return;
  int r = (int)(i / M), c = (int)(i % M);
  float2 a = __ldg(p1 + r), b = __ldg(p2 + c);
  float dx = a.x - b.x, dy = a.y - b.y;
  // op_patch.py:139-146: pow(sum(pow(diff, 2)), 0.5)
  out[i] = sqrtf(dx * dx + dy * dy);
}

}  // namespace bdet

using namespace bdet;

extern "C" int bdet_pairwise_batched(const float* boxes1, int ld1, int64_t bs1, const int* n1_dev, int N,
                                     const float* boxes2, int ld2, int64_t bs2, int M, float* out, int64_t ldo,
                                     int64_t bs_out, int B, int mode, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0 && M >= 0 && B >= 0, "negative size");
  BDET_REQUIRE(ld1 >= 4 && ld2 >= 4, "boxes must have >= 4 columns");
  BDET_REQUIRE(mode >= BDET_PAIR_IOU && mode <= BDET_PAIR_GIOU, "unknown mode");
  if (N == 0 || M == 0 || B == 0) return BDET_OK;
  BDET_REQUIRE(boxes1 && boxes2 && out, "null argument");
  BDET_REQUIRE(ldo >= M, "output row stride smaller than M");
  PairArgs a{boxes1, boxes2, out, n1_dev, bs1, bs2, bs_out, ld1, ld2, N, M, 0, ldo};
  int rc;
  switch (mode) {
    case BDET_PAIR_IOU: rc = launch_pairwise<BDET_PAIR_IOU>(a, B, as_stream(stream)); break;
    case BDET_PAIR_IOA: rc = launch_pairwise<BDET_PAIR_IOA>(a, B, as_stream(stream)); break;
    case BDET_PAIR_INTER: rc = launch_pairwise<BDET_PAIR_INTER>(a, B, as_stream(stream)); break;
    default: rc = launch_pairwise<BDET_PAIR_GIOU>(a, B, as_stream(stream)); break;
  }
  if (rc) return rc;
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_pairwise(const float* boxes1, int ld1, int N, const float* boxes2, int ld2, int M, float* out,
                             int64_t ldo, int mode, bdet_stream_t stream) {
  return bdet_pairwise_batched(boxes1, ld1, 0, nullptr, N, boxes2, ld2, 0, M, out, ldo, 0, 1, mode, stream);
}

extern "C" int bdet_box_center(const float* boxes, int ld, int N, float* out, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0 && ld >= 4, "bad shape");
  if (N == 0) return BDET_OK;
  BDET_REQUIRE(boxes && out, "null argument");
  BDET_KERNEL("box_center_kernel", as_stream(stream), box_center_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(boxes, ld, N, reinterpret_cast<float2*>(out)));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_point_distance(const float* p1, int N, const float* p2, int M, float* out, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0 && M >= 0, "bad shape");
  if (N == 0 || M == 0) return BDET_OK;
  BDET_REQUIRE(p1 && p2 && out, "null argument");
  BDET_KERNEL("point_distance_kernel", as_stream(stream), point_distance_kernel<<<ceil_div((int64_t)N * M, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float2*>(p1), N, reinterpret_cast<const float2*>(p2), M, out));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
