// a7-a9, a12: BoxCoder / SumBoxCoder / PointCoder encode+decode, Boxes.scale/clip/filter_by_size.
// Reference: basedet/structures/boxcoder.py:44-141, basedet/structures/boxes.py:132-212.
// Coalesced elementwise kernels: one box (128-bit load/store) per thread; ~25 tiny MegEngine oprs
// per call in the reference collapse into one pass (16 B/box in per operand, 16 B/box out).
#include "common.cuh"

namespace bdet {

static Vec4 to_vec4(const float* h, float dflt) {
  Vec4 v;
  for (int i = 0; i < 4; ++i) v.v[i] = h ? h[i] : dflt;
  return v;
}

__global__ void __launch_bounds__(256) box_encode_kernel(const float4* __restrict__ bbox, const float* __restrict__ gt, int gt_ld,
                                                         const int* __restrict__ gather, int N, Vec4 mean, Vec4 stdv,
                                                         float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  long long r = gather ? gather[i] : i;
  float4 g = (gt_ld == 4 && !gather) ? __ldg(reinterpret_cast<const float4*>(gt) + i) : load_box<false>(gt, r, gt_ld);
  out[i] = encode_box(__ldg(bbox + i), g, mean, stdv);
}

// decode for one (anchor, delta quadruple); boxcoder.py:75-98 (no clamp on dw/dh)
__device__ __forceinline__ float4 decode_box(float4 a, float4& d, const Vec4& mean, const Vec4& stdv) {
  d.x = d.x * stdv.v[0] + mean.v[0];
  d.y = d.y * stdv.v[1] + mean.v[1];
  d.z = d.z * stdv.v[2] + mean.v[2];
  d.w = d.w * stdv.v[3] + mean.v[3];
  float aw = a.z - a.x, ah = a.w - a.y;
  float acx = a.x + 0.5f * aw, acy = a.y + 0.5f * ah;
  float pcx = acx + d.x * aw;
  float pcy = acy + d.y * ah;
  float pw = aw * expf(d.z);
  float ph = ah * expf(d.w);
  float hw = 0.5f * pw, hh = 0.5f * ph;
  return make_float4(pcx - hw, pcy - hh, pcx + hw, pcy + hh);
}

// deltas (N, 4k): thread per (row, quadruple).  sel != nullptr: thread i decodes row sel[i] / sel_div.
__global__ void __launch_bounds__(256) box_decode_kernel(const float4* __restrict__ anchors, float4* deltas, long long total, int k,
                                                         Vec4 mean, Vec4 stdv, float4* __restrict__ out, int writeback,
                                                         const int* __restrict__ sel, int sel_div) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (sel) {
    long long r = sel[i] / sel_div;
    float4 d = deltas[r];
    out[i] = decode_box(__ldg(anchors + r), d, mean, stdv);
    return;
  }
  long long r = i / k;
  float4 d = deltas[i];
  out[i] = decode_box(__ldg(anchors + r), d, mean, stdv);
  if (writeback) deltas[i] = d;
}

__global__ void __launch_bounds__(256) sum_encode_kernel(const float4* __restrict__ anchors, const float4* __restrict__ gt, int N, Vec4 mean,
                                                         Vec4 stdv, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float4 a = __ldg(anchors + i), g = __ldg(gt + i);
  // boxcoder.py:116-120: (gt - anchors - mean) / std
  out[i] = make_float4(__fdiv_rn((g.x - a.x) - mean.v[0], stdv.v[0]), __fdiv_rn((g.y - a.y) - mean.v[1], stdv.v[1]),
                       __fdiv_rn((g.z - a.z) - mean.v[2], stdv.v[2]), __fdiv_rn((g.w - a.w) - mean.v[3], stdv.v[3]));
}

__global__ void __launch_bounds__(256) sum_decode_kernel(const float4* __restrict__ anchors, float4* deltas, int N, Vec4 mean, Vec4 stdv,
                                                         float4* __restrict__ out, int writeback) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float4 a = __ldg(anchors + i), d = deltas[i];
  d.x = d.x * stdv.v[0] + mean.v[0];
  d.y = d.y * stdv.v[1] + mean.v[1];
  d.z = d.z * stdv.v[2] + mean.v[2];
  d.w = d.w * stdv.v[3] + mean.v[3];
  out[i] = make_float4(a.x + d.x, a.y + d.y, a.z + d.z, a.w + d.w);
  if (writeback) deltas[i] = d;
}

// PointCoder.encode, boxcoder.py:132-133: out[g, a] = [p - gt[:2], gt[2:] - p]
__global__ void __launch_bounds__(256) point_encode_kernel(const float2* __restrict__ pts, int A, const float* __restrict__ gt, int gt_ld, int G,
                                                           float4* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)A * G) return;
  int g = (int)(i / A), a = (int)(i % A);
  float2 p = __ldg(pts + a);
  float4 b = load_box<false>(gt, g, gt_ld);
  out[i] = make_float4(p.x - b.x, p.y - b.y, b.z - p.x, b.w - p.y);
}

// PointCoder.encode on matched rows (models/det/fcos.py:268): out[a] = [p_a - gt_a[:2], gt_a[2:] - p_a]
__global__ void __launch_bounds__(256) point_encode_rows_kernel(const float2* __restrict__ pts, const float* __restrict__ gt, int gt_ld, int N,
                                                                float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float2 p = __ldg(pts + i);
  float4 b = load_box<false>(gt, i, gt_ld);
  out[i] = make_float4(p.x - b.x, p.y - b.y, b.z - p.x, b.w - p.y);
}

// PointCoder.decode, boxcoder.py:135-141
__global__ void __launch_bounds__(256) point_decode_kernel(const float2* __restrict__ pts, const float4* __restrict__ deltas, long long total,
                                                           int k, float4* __restrict__ out, const int* __restrict__ sel, int sel_div) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long r, src;
  if (sel) {
    r = sel[i] / sel_div;
    src = r;
  } else {
    r = i / k;
    src = i;
  }
  float2 p = __ldg(pts + r);
  float4 d = __ldg(deltas + src);
  out[i] = make_float4(p.x - d.x, p.y - d.y, p.x + d.z, p.y + d.w);
}

// Boxes.scale then Boxes.clip (boxes.py:193-212, :152-177): F.clip(x, 0, upper) = min(max(x, 0), upper)
__global__ void __launch_bounds__(256) scale_clip_kernel(float4* boxes, int N, float sw, float sh, float cw, float ch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float4 b = boxes[i];
  b.x *= sw;
  b.y *= sh;
  b.z *= sw;
  b.w *= sh;
  if (cw >= 0.f) {
    b.x = fminf(fmaxf(b.x, 0.f), cw);
    b.y = fminf(fmaxf(b.y, 0.f), ch);
    b.z = fminf(fmaxf(b.z, 0.f), cw);
    b.w = fminf(fmaxf(b.w, 0.f), ch);
  }
  boxes[i] = b;
}

// Boxes.filter_by_size, boxes.py:146-150 (keeps the reference's h/w naming swap):
//   h, w = width, height ; keep = (w > sizes[0]) & (h > sizes[1])
__global__ void __launch_bounds__(256) filter_by_size_kernel(const float4* __restrict__ boxes, int N, float s0, float s1,
                                                             uint8_t* __restrict__ keep) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float4 b = __ldg(boxes + i);
  float width = b.z - b.x, height = b.w - b.y;
  keep[i] = (height > s0) && (width > s1);
}

}  // namespace bdet

using namespace bdet;

extern "C" int bdet_box_encode(const float* bbox, const float* gt, int gt_ld, const int* gather_idx, int N,
                               const float* mean_host, const float* std_host, float* out, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0 && gt_ld >= 4, "bad shape");
  if (N == 0) return BDET_OK;
  BDET_REQUIRE(bbox && gt && out, "null argument");
  BDET_REQUIRE(aligned16(bbox) && aligned16(out), "bbox/out must be 16-byte aligned");
  BDET_REQUIRE(gt_ld != 4 || gather_idx || aligned16(gt), "gt must be 16-byte aligned");
  BDET_KERNEL("box_encode_kernel", as_stream(stream), box_encode_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(bbox), gt, gt_ld, gather_idx, N,
                                                                     to_vec4(mean_host, 0.f), to_vec4(std_host, 1.f),
                                                                     reinterpret_cast<float4*>(out)));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_box_decode(const float* anchors, float* deltas, int N, int k, const float* mean_host,
                               const float* std_host, float* out, int writeback, const int* sel_idx, int n_sel,
                               int sel_div, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0 && k >= 1, "bad shape");
  BDET_REQUIRE(!sel_idx || (k == 1 && sel_div >= 1 && n_sel >= 0), "selection decode needs k == 1 and sel_div >= 1");
  long long total = sel_idx ? n_sel : (long long)N * k;
  if (total == 0) return BDET_OK;
  BDET_REQUIRE(anchors && deltas && out, "null argument");
  BDET_REQUIRE(aligned16(anchors) && aligned16(deltas) && aligned16(out), "anchors/deltas/out must be 16-byte aligned");
  BDET_KERNEL("box_decode_kernel", as_stream(stream), box_decode_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(anchors), reinterpret_cast<float4*>(deltas), total, k, to_vec4(mean_host, 0.f),
      to_vec4(std_host, 1.f), reinterpret_cast<float4*>(out), sel_idx ? 0 : writeback, sel_idx, sel_div));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_sum_encode(const float* anchors, const float* gt, int N, const float* mean_host,
                               const float* std_host, float* out, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0, "bad shape");
  if (N == 0) return BDET_OK;
  BDET_REQUIRE(anchors && gt && out, "null argument");
  BDET_REQUIRE(aligned16(anchors) && aligned16(gt) && aligned16(out), "pointers must be 16-byte aligned");
  BDET_KERNEL("sum_encode_kernel", as_stream(stream), sum_encode_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(anchors),
                                                                     reinterpret_cast<const float4*>(gt), N, to_vec4(mean_host, 0.f),
                                                                     to_vec4(std_host, 1.f), reinterpret_cast<float4*>(out)));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_sum_decode(const float* anchors, float* deltas, int N, const float* mean_host,
                               const float* std_host, float* out, int writeback, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0, "bad shape");
  if (N == 0) return BDET_OK;
  BDET_REQUIRE(anchors && deltas && out, "null argument");
  BDET_REQUIRE(aligned16(anchors) && aligned16(deltas) && aligned16(out), "pointers must be 16-byte aligned");
  BDET_KERNEL("sum_decode_kernel", as_stream(stream), sum_decode_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(anchors),
                                                                     reinterpret_cast<float4*>(deltas), N, to_vec4(mean_host, 0.f),
                                                                     to_vec4(std_host, 1.f), reinterpret_cast<float4*>(out), writeback));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_point_encode(const float* points, int A, const float* gt, int gt_ld, int G, float* out,
                                 bdet_stream_t stream) {
  BDET_REQUIRE(A >= 0 && G >= 0 && gt_ld >= 4, "bad shape");
  if (A == 0 || G == 0) return BDET_OK;
  BDET_REQUIRE(points && gt && out, "null argument");
  BDET_REQUIRE(aligned16(out) && (reinterpret_cast<uintptr_t>(points) & 7u) == 0, "alignment");
  BDET_KERNEL("point_encode_kernel", as_stream(stream), point_encode_kernel<<<ceil_div((int64_t)A * G, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(points), A, gt, gt_ld,
                                                                                  G, reinterpret_cast<float4*>(out)));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_point_encode_rows(const float* points, const float* gt, int gt_ld, int N, float* out,
                                      bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0 && gt_ld >= 4, "bad shape");
  if (N == 0) return BDET_OK;
  BDET_REQUIRE(points && gt && out, "null argument");
  BDET_REQUIRE(aligned16(out) && (reinterpret_cast<uintptr_t>(points) & 7u) == 0, "alignment");
  BDET_KERNEL("point_encode_rows_kernel", as_stream(stream),
              point_encode_rows_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(points), gt, gt_ld, N,
                                                                                        reinterpret_cast<float4*>(out)));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_point_decode(const float* points, const float* deltas, int N, int k, float* out,
                                 const int* sel_idx, int n_sel, int sel_div, bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0 && k >= 1, "bad shape");
  BDET_REQUIRE(!sel_idx || (k == 1 && sel_div >= 1 && n_sel >= 0), "selection decode needs k == 1 and sel_div >= 1");
  long long total = sel_idx ? n_sel : (long long)N * k;
  if (total == 0) return BDET_OK;
  BDET_REQUIRE(points && deltas && out, "null argument");
  BDET_REQUIRE(aligned16(deltas) && aligned16(out) && (reinterpret_cast<uintptr_t>(points) & 7u) == 0, "alignment");
  BDET_KERNEL("point_decode_kernel", as_stream(stream), point_decode_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(points),
                                                                           reinterpret_cast<const float4*>(deltas), total, k,
                                                                           reinterpret_cast<float4*>(out), sel_idx, sel_div));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_boxes_scale_clip(float* boxes, int N, float scale_w, float scale_h, float clip_w, float clip_h,
                                     bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0, "bad shape");
  if (N == 0) return BDET_OK;
  BDET_REQUIRE(boxes && aligned16(boxes), "boxes must be a 16-byte aligned device pointer");
  BDET_KERNEL("scale_clip_kernel", as_stream(stream), scale_clip_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(boxes), N, scale_w, scale_h, clip_w, clip_h));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_boxes_filter_by_size(const float* boxes, int N, float size0, float size1, uint8_t* keep_mask,
                                         bdet_stream_t stream) {
  BDET_REQUIRE(N >= 0, "bad shape");
  if (N == 0) return BDET_OK;
  BDET_REQUIRE(boxes && keep_mask && aligned16(boxes), "boxes must be a 16-byte aligned device pointer");
  BDET_KERNEL("filter_by_size_kernel", as_stream(stream), filter_by_size_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(boxes), N, size0, size1, keep_mask));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
