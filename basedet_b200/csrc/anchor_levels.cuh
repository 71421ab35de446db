// The anchor grid of all FPN levels as a value (kernel parameter): shared by anchors.cu (materialises the anchors) and
// assign.cu (generates them in registers).  Reference: basedet/layers/common/anchor_generator.py:23-30, 111-122.
#pragma once
#include "common.cuh"

namespace bdet {

constexpr int kMaxBase = 64;

struct AnchorLevels {
  int n_levels;
  int H[BDET_MAX_LEVELS], W[BDET_MAX_LEVELS], n_base[BDET_MAX_LEVELS], base_off[BDET_MAX_LEVELS];
  double stride[BDET_MAX_LEVELS], shift[BDET_MAX_LEVELS];
  long long out_off[BDET_MAX_LEVELS];    // in boxes / points
  long long start[BDET_MAX_LEVELS + 1];  // prefix of per-level work items
  float base[kMaxBase * 4];
};

// anchor i of the level-concatenated (h, w, base) order; *slot = where bdet_anchors_grid stores it
__device__ __forceinline__ float4 anchor_at(const AnchorLevels& p, long long i, long long* slot) {
  int l = 0;
#pragma unroll
  for (int k = 1; k < BDET_MAX_LEVELS; ++k)
    if (k < p.n_levels && i >= p.start[k]) l = k;
  const long long r = i - p.start[l];
  const int nb = p.n_base[l];
  const int a = (int)(r % nb);
  const long long pos = r / nb;
  const int w = (int)(pos % p.W[l]);
  const int h = (int)(pos / p.W[l]);
  // F.arange(shift, n*stride + shift, stride): fp32(start + i*step) evaluated in fp64 (oracle ASSUMED-7)
  const float x = (float)(p.shift[l] + (double)w * p.stride[l]);
  const float y = (float)(p.shift[l] + (double)h * p.stride[l]);
  const float* b = p.base + (p.base_off[l] + a) * 4;
  if (slot) *slot = p.out_off[l] + r;
  return make_float4(x + b[0], y + b[1], x + b[2], y + b[3]);
}

static inline int fill_levels(AnchorLevels* p, int n_levels, const int* hw, const double* stride, const double* shift,
                       const int* n_base, int num_anchors, const float* base, const int64_t* out_off) {
  if (n_levels < 1 || n_levels > BDET_MAX_LEVELS)
    return set_error(BDET_EINVAL, "anchors: n_levels must be in [1, %d]", BDET_MAX_LEVELS);
  p->n_levels = n_levels;
  p->start[0] = 0;
  int boff = 0;
  for (int l = 0; l < n_levels; ++l) {
    p->H[l] = hw[2 * l];
    p->W[l] = hw[2 * l + 1];
    if (p->H[l] < 0 || p->W[l] < 0) return set_error(BDET_EINVAL, "anchors: negative feature size");
    p->n_base[l] = n_base ? n_base[l] : num_anchors;
    if (p->n_base[l] < 1) return set_error(BDET_EINVAL, "anchors: need >= 1 anchor per cell");
    p->base_off[l] = boff;
    boff += p->n_base[l];
    p->stride[l] = stride[l];
    p->shift[l] = shift ? shift[l] : 0.0;
    p->out_off[l] = out_off[l];
    p->start[l + 1] = p->start[l] + (long long)p->H[l] * p->W[l] * p->n_base[l];
  }
  if (base) {
    if (boff > kMaxBase) return set_error(BDET_EUNSUPPORTED, "anchors: more than %d base anchors", kMaxBase);
    for (int i = 0; i < boff * 4; ++i) p->base[i] = base[i];
  }
  return BDET_OK;
}

}  // namespace bdet
