// Shared helpers for libbdet.so (sm_100a only).  Compiled with -fmad=false: every fp32
// op below rounds once, in the reference's operation order.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "bdet.h"

namespace bdet {

int set_error(int code, const char* fmt, ...);
int sm_count();

#define BDET_REQUIRE(cond, msg)                                        \
  do {                                                                 \
    if (!(cond)) return ::bdet::set_error(BDET_EINVAL, "%s: %s", __func__, msg); \
  } while (0)

#define BDET_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess)                                                              \
      return ::bdet::set_error(BDET_ECUDA, "%s: %s", __func__, cudaGetErrorString(e__)); \
  } while (0)

#define BDET_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return ::bdet::set_error(BDET_ECUDA, "%s: %s", __func__, cudaGetErrorString(e__)); \
  } while (0)

// Measurement hooks (bdet_profile_*): bracket a named kernel launch with an event pair when profiling is on.
void prof_pre(cudaStream_t st, const char* name);
void prof_post(cudaStream_t st);
#define BDET_KERNEL(name, st, ...) \
  do {                             \
    ::bdet::prof_pre(st, name);    \
    __VA_ARGS__;                   \
    ::bdet::prof_post(st);         \
  } while (0)

static inline cudaStream_t as_stream(bdet_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// ---- device helpers -------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Load a box row with arbitrary stride / alignment.
template <bool kVec>
__device__ __forceinline__ float4 load_box(const float* base, int64_t row, int ld) {
  if (kVec) return ldg4(base + row * 4);
  const float* p = base + row * ld;
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

__device__ __forceinline__ float box_area(float4 b) { return (b.z - b.x) * (b.w - b.y); }

// Pairwise IoU in the op order of structures/op_patch.py:50-74.  FMNMX replaces the reference's
// x>y?x:y selects: identical for all non-NaN inputs after the clamps; any NaN coordinate makes
// the union NaN and the final max(.,0) returns 0 in both formulations.
// inter == 0 (or NaN) always yields exactly +0, so the IEEE division is skipped for it.
__device__ __forceinline__ float iou_pair(float4 a, float area_a, float4 b, float area_b) {
  float iw = fminf(a.z, b.z) - fmaxf(a.x, b.x);
  float ih = fminf(a.w, b.w) - fmaxf(a.y, b.y);
  iw = fmaxf(iw, 0.f);
  ih = fmaxf(ih, 0.f);
  float inter = iw * ih;
  float r = 0.f;
  if (inter > 0.f) {
    float uni = (area_a + area_b) - inter;
    r = fmaxf(__fdiv_rn(inter, uni), 0.f);
  }
  return r;
}

// NMS overlap predicate (oracle ASSUMED-5): IoU = inter / (Sa + Sb - inter); suppress iff IoU > thr.  inter == 0 can only
// exceed a negative threshold, so the division is skipped for it when thr >= 0.
__device__ __forceinline__ bool nms_overlap(float4 a, float sa, float4 b, float sb, float thr) {
  float w = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.f);
  float h = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.f);
  float inter = w * h;
  if (!(inter > 0.f) && thr >= 0.f) return false;
  return __fdiv_rn(inter, (sa + sb) - inter) > thr;
}

// Order-preserving float <-> uint32 map (ascending).
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}

// Matcher threshold labelling, layers/common/matcher.py:43-45 (intervals [thr[k], thr[k+1])).
struct MatchCfg {
  float thr[BDET_MAX_MATCH_LABELS + 1];  // thr[0] = -inf, thr[n] = +inf
  int lab[BDET_MAX_MATCH_LABELS];
  int n;
};
__device__ __forceinline__ int threshold_label(const MatchCfg& c, float m) {
  int out = -1;
#pragma unroll
  for (int k = 0; k < BDET_MAX_MATCH_LABELS; ++k)
    if (k < c.n && m >= c.thr[k] && m < c.thr[k + 1]) out = c.lab[k];
  return out;
}
int make_match_cfg(MatchCfg* cfg, const float* thresholds_host, const int* labels_host, int n_labels);

struct Vec4 {
  float v[4];
};

// BoxCoder._box_ltrb_to_cs_opr + encode, structures/boxcoder.py:44-73.
// kUnit: mean == 0 and std == 1, where (t - 0) / 1 == t bit for bit (only the sign of a zero could differ:
// -0 - 0 = -0, -0 / 1 = -0), so the four subtractions and IEEE divisions are skipped.
template <bool kUnit = false>
__device__ __forceinline__ float4 encode_box(float4 b, float4 g, const Vec4& mean, const Vec4& stdv) {
  float bw = b.z - b.x, bh = b.w - b.y;
  float bcx = b.x + 0.5f * bw, bcy = b.y + 0.5f * bh;
  float gw = g.z - g.x, gh = g.w - g.y;
  float gcx = g.x + 0.5f * gw, gcy = g.y + 0.5f * gh;
  float dx = __fdiv_rn(gcx - bcx, bw);
  float dy = __fdiv_rn(gcy - bcy, bh);
  float dw = logf(__fdiv_rn(gw, bw));
  float dh = logf(__fdiv_rn(gh, bh));
  if (kUnit) return make_float4(dx, dy, dw, dh);
  float4 t;
  t.x = __fdiv_rn(dx - mean.v[0], stdv.v[0]);
  t.y = __fdiv_rn(dy - mean.v[1], stdv.v[1]);
  t.z = __fdiv_rn(dw - mean.v[2], stdv.v[2]);
  t.w = __fdiv_rn(dh - mean.v[3], stdv.v[3]);
  return t;
}

}  // namespace bdet
