// Batched glue between top-k, NMS and the final detections (the per-image / per-level Python loops of the reference):
//   select_decode_kernel : models/det/retinanet.py:193-196 / fcos.py:204-207 (label = idx % C, box = decode(...)[idx // C])
//                          models/det/rpn.py:149-171 (decode, top-k gather, level ids, clip -> filter_by_size mask)
//                          + level concat of layers/common/post_processing.py:63-67 / rpn.py:163-165
//   finalize_kernel      : layers/common/post_processing.py:96-101 (gather kept, scale to the original image, clip)
//                          and rpn.py:179-183 (rois = [batch, x1, y1, x2, y2])
// Only the <= k selected candidates per (image, level) are decoded -- the reference decodes every anchor and then
// gathers.  One CTA per image keeps the reference's ordering (levels in order, score-descending inside).
#include "common.cuh"
#include "sortnet.cuh"

namespace bdet {

constexpr int kDetThreads = 1024;

struct SelectArgs {
  const float* anchors[BDET_MAX_LEVELS];  // (n_l, 4) boxes or (n_l, 2) points
  const float* deltas[BDET_MAX_LEVELS];   // (B, n_l, 4)
  int n_l[BDET_MAX_LEVELS];
  int hw[BDET_MAX_LEVELS];                // > 0: deltas of this level are the head output (B, A*4, H, W), hw = H*W
  int L, B, k, div, coder, label_mode, filter;
  const int* topk_idx;     // (B, L, k) flat index within the (image, level) segment
  const float* topk_val;   // (B, L, k)
  const int* topk_cnt;     // (B, L)
  const float* im_info;    // (B, info_ld): [h, w, ...] used by the size filter
  int info_ld;
  Vec4 mean, stdv;
  float* boxes;   // (B, L*k, 4)
  float* scores;  // (B, L*k)
  void* labels;   // (B, L*k) int32 (label_mode 0) or fp32 level ids (label_mode 1)
  int* count;     // (B)
  int* run_end;   // (B, L) optional
  uint32_t* flagw;  // workspace: (B, L, nblk, 8) keep ballots of the 256-candidate blocks (filter path)
  int* blkcnt;      // workspace: (B, L, nblk) kept candidates per block
};

__device__ __forceinline__ float4 decode_one(const SelectArgs& p, int l, int b, int a) {
  float4 d;
  if (p.hw[l] > 0) {  // NCHW head output: box a = pos * A + anchor, component c lives at ((anchor*4 + c) * HW + pos)
    const int hw = p.hw[l], na = p.n_l[l] / hw;
    const int pos = a / na, an = a - pos * na;
    const float* q = p.deltas[l] + ((long long)b * na * 4 + an * 4) * hw + pos;
    d = make_float4(__ldg(q), __ldg(q + hw), __ldg(q + 2 * hw), __ldg(q + 3 * hw));
  } else {
    d = ldg4(p.deltas[l] + ((long long)b * p.n_l[l] + a) * 4);
  }
  if (p.coder == 1) {  // PointCoder.decode, structures/boxcoder.py:135-141
    const float2 pt = __ldg(reinterpret_cast<const float2*>(p.anchors[l]) + a);
    return make_float4(pt.x - d.x, pt.y - d.y, pt.x + d.z, pt.y + d.w);
  }
  const float4 an = ldg4(p.anchors[l] + (long long)a * 4);  // BoxCoder.decode, structures/boxcoder.py:75-98
  const float dx = d.x * p.stdv.v[0] + p.mean.v[0], dy = d.y * p.stdv.v[1] + p.mean.v[1];
  const float dw = d.z * p.stdv.v[2] + p.mean.v[2], dh = d.w * p.stdv.v[3] + p.mean.v[3];
  const float aw = an.z - an.x, ah = an.w - an.y;
  const float acx = an.x + 0.5f * aw, acy = an.y + 0.5f * ah;
  const float pcx = acx + dx * aw, pcy = acy + dy * ah;
  const float pw = aw * expf(dw), ph = ah * expf(dh);
  const float hw = 0.5f * pw, hh = 0.5f * ph;
  return make_float4(pcx - hw, pcy - hh, pcx + hw, pcy + hh);
}

__global__ void __launch_bounds__(kDetThreads) select_decode_kernel(const SelectArgs p) {
  __shared__ int warp_cnt[kDetThreads / 32];
  __shared__ int sbase;
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long cap = (long long)p.L * p.k;
  float4* ob = reinterpret_cast<float4*>(p.boxes) + b * cap;
  float* os = p.scores + b * cap;
  if (t == 0) sbase = 0;
  __syncthreads();
  float ih = 0.f, iw = 0.f;
  if (p.filter) {
    ih = __ldg(p.im_info + (long long)b * p.info_ld);
    iw = __ldg(p.im_info + (long long)b * p.info_ld + 1);
  }
  for (int l = 0; l < p.L; ++l) {
    const int seg = b * p.L + l;
    const int cnt = min(p.topk_cnt[seg], p.k);
    for (int j0 = 0; j0 < cnt; j0 += kDetThreads) {
      const int j = j0 + t;
      bool keep = j < cnt;
      float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
      int idx = 0;
      if (keep) {
        idx = __ldg(p.topk_idx + (long long)seg * p.k + j);
        bx = decode_one(p, l, b, idx / p.div);
        if (p.filter) {
          // rpn.py:168-169: Boxes(proposals).clip(im_info[:2]).filter_by_size(): (h > 0) & (w > 0) of the CLIPPED
          // box; the proposal itself keeps its un-clipped coordinates
          const float x1 = fminf(fmaxf(bx.x, 0.f), iw), y1 = fminf(fmaxf(bx.y, 0.f), ih);
          const float x2 = fminf(fmaxf(bx.z, 0.f), iw), y2 = fminf(fmaxf(bx.w, 0.f), ih);
          keep = (y2 - y1 > 0.f) && (x2 - x1 > 0.f);
        }
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) warp_cnt[warp] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
      for (int w = 0; w < kDetThreads / 32; ++w) {
        const int c = warp_cnt[w];
        if (w < warp) before += c;
        total += c;
      }
      const int base = sbase;
      if (keep) {
        const long long o = base + before + __popc(bal & ((1u << lane) - 1u));
        ob[o] = bx;
        os[o] = __ldg(p.topk_val + (long long)seg * p.k + j);
        if (p.label_mode == 0)
          reinterpret_cast<int*>(p.labels)[b * cap + o] = idx % p.div;   // retinanet.py:194
        else
          reinterpret_cast<float*>(p.labels)[b * cap + o] = (float)l;    // rpn.py:160 F.full_like(scores, level)
      }
      __syncthreads();
      if (t == 0) sbase = base + total;
      __syncthreads();
    }
    if (t == 0 && p.run_end) p.run_end[seg] = sbase;
  }
  if (t == 0) p.count[b] = sbase;
}

// No size filter (dense heads): the output slot of candidate j of level l is known from the per-level counts alone,
// so every (image, level, 256 candidates) block works independently -- no per-image serial compaction.
__global__ void __launch_bounds__(256) select_decode_flat_kernel(const SelectArgs p) {
  const int b = blockIdx.z, l = blockIdx.y;
  const int seg = b * p.L + l;
  int before = 0, total = 0;
  for (int q = 0; q < p.L; ++q) {
    const int c = min(p.topk_cnt[b * p.L + q], p.k);
    if (q < l) before += c;
    total += c;
  }
  const int cnt = min(p.topk_cnt[seg], p.k);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (p.run_end) p.run_end[seg] = before + cnt;
    if (l == 0) p.count[b] = total;
  }
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= cnt) return;
  const long long cap = (long long)p.L * p.k;
  const long long o = b * cap + before + j;
  const int idx = __ldg(p.topk_idx + (long long)seg * p.k + j);
  reinterpret_cast<float4*>(p.boxes)[o] = decode_one(p, l, b, idx / p.div);
  p.scores[o] = __ldg(p.topk_val + (long long)seg * p.k + j);
  if (p.label_mode == 0)
    reinterpret_cast<int*>(p.labels)[o] = idx % p.div;   // retinanet.py:194
  else
    reinterpret_cast<float*>(p.labels)[o] = (float)l;    // rpn.py:160
}

// Size filter with a workspace: two fully parallel launches over (256 candidates, level, image) blocks instead of one
// serial CTA per image.  Pass 1 decodes, tests and stores the keep ballots and block counts; pass 2 sums the counts of
// the blocks in front (<= L * nblk values), decodes again (cheaper than parking 16 B per candidate) and writes.
__device__ __forceinline__ bool rpn_keep(const SelectArgs& p, int b, int l, int j, int cnt, float4& bx, int& idx) {
  bool keep = j < cnt;
  bx = make_float4(0.f, 0.f, 0.f, 0.f);
  idx = 0;
  if (keep) {
    idx = __ldg(p.topk_idx + (long long)(b * p.L + l) * p.k + j);
    bx = decode_one(p, l, b, idx / p.div);
    const float ih = __ldg(p.im_info + (long long)b * p.info_ld), iw = __ldg(p.im_info + (long long)b * p.info_ld + 1);
    const float x1 = fminf(fmaxf(bx.x, 0.f), iw), y1 = fminf(fmaxf(bx.y, 0.f), ih);  // rpn.py:168-169, see above
    const float x2 = fminf(fmaxf(bx.z, 0.f), iw), y2 = fminf(fmaxf(bx.w, 0.f), ih);
    keep = (y2 - y1 > 0.f) && (x2 - x1 > 0.f);
  }
  return keep;
}

__global__ void __launch_bounds__(256) select_flags_kernel(const SelectArgs p) {
  __shared__ int wc[8];
  const int b = blockIdx.z, l = blockIdx.y, blk = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nblk = gridDim.x;
  const int cnt = min(p.topk_cnt[b * p.L + l], p.k);
  float4 bx;
  int idx;
  const uint32_t bal = __ballot_sync(0xffffffffu, rpn_keep(p, b, l, blk * 256 + t, cnt, bx, idx));
  const long long e = ((long long)b * p.L + l) * nblk + blk;
  if (lane == 0) {
    p.flagw[e * 8 + warp] = bal;
    wc[warp] = __popc(bal);
  }
  __syncthreads();
  if (t == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += wc[w];
    p.blkcnt[e] = s;
  }
}

__global__ void __launch_bounds__(256) select_write_kernel(const SelectArgs p) {
  __shared__ int sred[8];
  const int b = blockIdx.z, l = blockIdx.y, blk = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nblk = gridDim.x;
  const int mine = l * nblk + blk, total = p.L * nblk;
  const int* bc = p.blkcnt + (long long)b * total;
  int before = 0, all = 0;
  for (int i = t; i < total; i += 256) {
    const int c = bc[i];
    if (i < mine) before += c;
    all += c;
  }
  before = __reduce_add_sync(0xffffffffu, before);
  all = __reduce_add_sync(0xffffffffu, all);
  if (lane == 0) sred[warp] = before;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < 8; ++w) base += sred[w];
  __syncthreads();
  if (lane == 0) sred[warp] = all;
  __syncthreads();
  if (mine == 0 && t == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += sred[w];
    p.count[b] = s;
  }
  if (p.run_end && blk == nblk - 1 && t == 0) p.run_end[b * p.L + l] = base + bc[mine];
  const uint32_t* fw = p.flagw + ((long long)b * total + mine) * 8;
  const uint32_t bal = fw[warp];
  if (!((bal >> lane) & 1u)) return;
  int pos = base + __popc(bal & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) pos += __popc(fw[w]);
  const int cnt = min(p.topk_cnt[b * p.L + l], p.k);
  const int j = blk * 256 + t;
  float4 bx;
  int idx;
  rpn_keep(p, b, l, j, cnt, bx, idx);
  const long long cap = (long long)p.L * p.k, o = b * cap + pos;
  reinterpret_cast<float4*>(p.boxes)[o] = bx;
  p.scores[o] = __ldg(p.topk_val + (long long)(b * p.L + l) * p.k + j);
  if (p.label_mode == 0)
    reinterpret_cast<int*>(p.labels)[o] = idx % p.div;   // retinanet.py:194
  else
    reinterpret_cast<float*>(p.labels)[o] = (float)l;    // rpn.py:160
}

struct FinalArgs {
  const float* boxes;   // (B, N, 4)
  const float* scores;  // (B, N)
  const void* labels;   // (B, N) int32 / fp32
  const int* keep;      // (B, keep_ld)
  const int* keep_count;
  const float* im_info;  // (B, info_ld) [h, w, orig_h, orig_w, ...] or NULL
  int info_ld, N, keep_ld, max_out, mode, labels_float, B;
  float* out;  // mode 0: (B, max_out, 6) [x1,y1,x2,y2,score,label]; mode 1: (B, max_out, 5) [batch,x1,y1,x2,y2]
};

__global__ void __launch_bounds__(256) finalize_kernel(const FinalArgs p) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.max_out) return;
  const int cnt = min(p.keep_count[b], p.max_out);
  const int row = p.mode == 0 ? 6 : 5;
  float* o = p.out + ((long long)b * p.max_out + j) * row;
  if (j >= cnt) {
    for (int q = 0; q < row; ++q) o[q] = 0.f;
    return;
  }
  const int src = p.keep[(long long)b * p.keep_ld + j];
  float4 bx = ldg4(p.boxes + ((long long)b * p.N + src) * 4);
  if (p.mode == 0) {
    if (p.im_info) {
      const float* info = p.im_info + (long long)b * p.info_ld;
      // post_processing.py:99-101: scale ratios (orig / resized) then clip to the original image
      const float sh = __fdiv_rn(info[2], info[0]), sw = __fdiv_rn(info[3], info[1]);
      bx.x *= sw;
      bx.y *= sh;
      bx.z *= sw;
      bx.w *= sh;
      const float ch = info[2], cw = info[3];
      bx.x = fminf(fmaxf(bx.x, 0.f), cw);
      bx.y = fminf(fmaxf(bx.y, 0.f), ch);
      bx.z = fminf(fmaxf(bx.z, 0.f), cw);
      bx.w = fminf(fmaxf(bx.w, 0.f), ch);
    }
    o[0] = bx.x;
    o[1] = bx.y;
    o[2] = bx.z;
    o[3] = bx.w;
    o[4] = p.scores[(long long)b * p.N + src];
    o[5] = p.labels_float ? reinterpret_cast<const float*>(p.labels)[(long long)b * p.N + src]
                          : (float)reinterpret_cast<const int*>(p.labels)[(long long)b * p.N + src];
  } else {
    o[0] = (float)b;  // rpn.py:181-182: F.full((n, 1), bid) ++ proposals
    o[1] = bx.x;
    o[2] = bx.y;
    o[3] = bx.z;
    o[4] = bx.w;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Dense tail: select_decode -> sort -> class-aware NMS -> finalize of a dense head (RetinaNet / FCOS inference,
// retinanet.py:193-209, fcos.py:204-221, post_processing.py:17-47,78-103) as ONE kernel, one CTA per image.
// The four launches it replaces are single-CTA, latency-bound kernels (7-17 us each); here their intermediates never
// leave shared memory:
//   (1) the per-level top-k runs (already score-sorted) are merged by ranking -- own position + a binary search in every
//       other run over the scores staged in shared memory -- and every candidate is decoded straight into its sorted slot;
//   (2) the class offset `label * (max(boxes) + 1)` (post_processing.py:44-46) is applied on the fly;
//   (3) keep-driven NMS over 64-box blocks against the kept list (the nms_fused_kernel scheme), stopping at max_out;
//   (4) the kept boxes are scaled / clipped and written as (max_out, 6) rows, zero padded.
// Decisions are the same predicates on the same fp32 values as the separate kernels: identical detections.
struct TailArgs {
  SelectArgs s;
  float thr;
  int max_out;
  float* dets;     // (B, max_out, 6)
  int* det_count;  // (B)
};

__global__ void __launch_bounds__(kDetThreads) dense_tail_kernel(const TailArgs q) {
  extern __shared__ __align__(16) unsigned char raw[];
  const SelectArgs& p = q.s;
  const int cap = p.L * p.k;
  float4* sbox = reinterpret_cast<float4*>(raw);                 // cap: decoded boxes in sorted order
  float* sscore = reinterpret_cast<float*>(sbox + cap);          // cap: sorted scores
  int* slabel = reinterpret_cast<int*>(sscore + cap);            // cap: sorted labels
  float* rscore = reinterpret_cast<float*>(slabel + cap);        // cap: the runs' scores, level-major (for the ranking)
  float4* kbox = reinterpret_cast<float4*>(rscore + cap);        // max_out: shifted boxes kept so far
  float* karea = reinterpret_cast<float*>(kbox + q.max_out);     // max_out
  int* kidx = reinterpret_cast<int*>(karea + q.max_out);         // max_out: sorted position of the kept boxes
  __shared__ int rend[BDET_MAX_LEVELS + 1];
  __shared__ uint32_t smax;
  __shared__ float4 srow[64];
  __shared__ float sarea[64];
  __shared__ uint64_t sdiag[64];
  __shared__ int skept[64];
  __shared__ int snk;
  __shared__ uint32_t srem[2];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) {
    int s = 0;
    rend[0] = 0;
    for (int l = 0; l < p.L; ++l) {
      s += min(p.topk_cnt[b * p.L + l], p.k);
      rend[l + 1] = s;
    }
    smax = 0u;
  }
  __syncthreads();
  const int n = rend[p.L];
  for (int l = 0; l < p.L; ++l) {
    const int cnt = rend[l + 1] - rend[l];
    const float* v = p.topk_val + (long long)(b * p.L + l) * p.k;
    for (int j = t; j < cnt; j += kDetThreads) rscore[rend[l] + j] = __ldg(v + j);
  }
  __syncthreads();
  // (1) rank + decode.  Order = (score desc, position in the level-major concatenation asc), the key of nms_sort_small.
  float m = -CUDART_INF_F;
  for (int i = t; i < n; i += kDetThreads) {
    int l = 0;
    while (rend[l + 1] <= i) ++l;
    const uint64_t key = make_key(rscore[i], (uint32_t)i);
    int rank = 0;
    for (int r = 0; r < p.L; ++r) {
      int lo = rend[r], hi = rend[r + 1];
      if (r == l) {
        rank += i - lo;
      } else {
        const int start = lo;
        while (lo < hi) {  // lower bound among unique keys
          const int mid = (lo + hi) >> 1;
          if (make_key(rscore[mid], (uint32_t)mid) < key) lo = mid + 1;
          else hi = mid;
        }
        rank += lo - start;
      }
    }
    const int idx = __ldg(p.topk_idx + (long long)(b * p.L + l) * p.k + (i - rend[l]));
    const float4 bx = decode_one(p, l, b, idx / p.div);
    sbox[rank] = bx;
    sscore[rank] = rscore[i];
    slabel[rank] = idx % p.div;                                   // retinanet.py:194
    m = fmaxf(m, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
  }
  {
    const uint32_t w = __reduce_max_sync(0xffffffffu, f2ord(m));
    if (lane == 0 && n > 0) atomicMax(&smax, w);
  }
  __syncthreads();
  const float shift_unit = ord2f(smax) + 1.f;                     // post_processing.py:45: max_coordinate + 1
  auto shifted = [&](int i) {                                     // :46 boxes + offsets
    float4 bx = sbox[i];
    const float off = (float)slabel[i] * shift_unit;
    bx.x += off;
    bx.y += off;
    bx.z += off;
    bx.w += off;
    return bx;
  };
  // (3) keep-driven NMS (see nms_fused_kernel)
  const int max_out = min(q.max_out, cap);
  const int nblk = (n + 63) >> 6;
  int count = 0;
  for (int blk = 0; blk < nblk && count < max_out; ++blk) {
    __syncthreads();
    const int r0 = blk * 64;
    const int nv = min(64, n - r0);
    const uint64_t valid = nv >= 64 ? ~0ull : ((1ull << nv) - 1ull);
    if (t < 64) {
      const float4 bx = t < nv ? shifted(r0 + t) : make_float4(0.f, 0.f, 0.f, 0.f);
      srow[t] = bx;
      sarea[t] = box_area(bx);
      sdiag[t] = 0ull;
    }
    if (t < 2) srem[t] = 0u;
    __syncthreads();
    {
      const int col = ((warp & 1) << 5) | lane, slice = warp >> 1;
      bool sup = false;
      if (col < nv) {
        const float4 c = srow[col];
        const float ca = sarea[col];
        for (int k2 = slice; k2 < count && !sup; k2 += kDetThreads / 64) sup = nms_overlap(kbox[k2], karea[k2], c, ca, q.thr);
      }
      const uint32_t mm = __ballot_sync(0xffffffffu, sup);
      if (lane == 0 && mm) atomicOr(&srem[warp & 1], mm);
    }
    __syncthreads();
    const uint64_t alive = ~(((uint64_t)srem[1] << 32) | srem[0]) & valid;
    if (alive == 0ull) continue;
    {
      const float4 c0 = srow[lane], c1 = srow[lane + 32];
      const float a0 = sarea[lane], a1 = sarea[lane + 32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = warp + 32 * h;
        if (!((alive >> i) & 1ull)) continue;
        const float4 a = srow[i];
        const float sa = sarea[i];
        const bool p0 = (lane > i) && (lane < nv) && nms_overlap(a, sa, c0, a0, q.thr);
        const bool p1 = (lane + 32 > i) && (lane + 32 < nv) && nms_overlap(a, sa, c1, a1, q.thr);
        const uint32_t lo = __ballot_sync(0xffffffffu, p0);
        const uint32_t hi = __ballot_sync(0xffffffffu, p1);
        if (lane == 0) sdiag[i] = ((uint64_t)hi << 32) | lo;
      }
    }
    __syncthreads();
    if (warp == 0) {
      uint64_t cand = alive;
      int nk = 0;
      while (cand != 0ull && count + nk < max_out) {
        const int i = __ffsll((long long)cand) - 1;
        if (lane == 0) skept[nk] = i;
        ++nk;
        cand &= ~sdiag[i];
        cand &= ~(1ull << i);
      }
      if (lane == 0) snk = nk;
    }
    __syncthreads();
    const int nk = snk;
    if (t < nk) {
      const int i = skept[t];
      kbox[count + t] = srow[i];
      karea[count + t] = sarea[i];
      kidx[count + t] = r0 + i;
    }
    count += nk;
  }
  __syncthreads();
  // (4) finalize, post_processing.py:96-101
  float sh = 1.f, sw = 1.f, ch = 0.f, cw = 0.f;
  if (p.im_info) {
    const float* info = p.im_info + (long long)b * p.info_ld;
    sh = __fdiv_rn(info[2], info[0]);
    sw = __fdiv_rn(info[3], info[1]);
    ch = info[2];
    cw = info[3];
  }
  for (int j = t; j < q.max_out; j += kDetThreads) {
    float* o = q.dets + ((long long)b * q.max_out + j) * 6;
    if (j >= count) {
#pragma unroll
      for (int c = 0; c < 6; ++c) o[c] = 0.f;
      continue;
    }
    const int i = kidx[j];
    float4 bx = sbox[i];
    if (p.im_info) {
      bx.x *= sw;
      bx.y *= sh;
      bx.z *= sw;
      bx.w *= sh;
      bx.x = fminf(fmaxf(bx.x, 0.f), cw);
      bx.y = fminf(fmaxf(bx.y, 0.f), ch);
      bx.z = fminf(fmaxf(bx.z, 0.f), cw);
      bx.w = fminf(fmaxf(bx.w, 0.f), ch);
    }
    o[0] = bx.x;
    o[1] = bx.y;
    o[2] = bx.z;
    o[3] = bx.w;
    o[4] = sscore[i];
    o[5] = (float)slabel[i];
  }
  if (t == 0) q.det_count[b] = count;
}

}  // namespace bdet

using namespace bdet;

extern "C" int bdet_select_decode(const float* const* anchors_host, const float* const* deltas_host, const int* n_l_host, int L,
                                  int B, int k, int div, int coder, int label_mode, const int* topk_idx,
                                  const float* topk_val, const int* topk_cnt, const float* mean_host, const float* std_host,
                                  const float* im_info, int info_ld, float* boxes, float* scores, void* labels, int* count,
                                  int* run_end, bdet_stream_t stream) {
  return bdet_select_decode_ws(anchors_host, deltas_host, n_l_host, nullptr, L, B, k, div, coder, label_mode, topk_idx, topk_val,
                               topk_cnt, mean_host, std_host, im_info, info_ld, boxes, scores, labels, count, run_end, nullptr, 0,
                               stream);
}

extern "C" size_t bdet_select_decode_workspace(int L, int B, int k) {
  if (L <= 0 || B <= 0 || k <= 0) return 16;
  return (size_t)B * L * ceil_div(k, 256) * (8 + 1) * 4 + 256;
}

extern "C" int bdet_select_decode_nchw(const float* const* anchors_host, const float* const* deltas_host, const int* n_l_host,
                                       const int* hw_host, int L, int B, int k, int div, int coder, int label_mode,
                                       const int* topk_idx, const float* topk_val, const int* topk_cnt,
                                       const float* mean_host, const float* std_host, const float* im_info, int info_ld,
                                       float* boxes, float* scores, void* labels, int* count, int* run_end,
                                       bdet_stream_t stream) {
  return bdet_select_decode_ws(anchors_host, deltas_host, n_l_host, hw_host, L, B, k, div, coder, label_mode, topk_idx, topk_val,
                               topk_cnt, mean_host, std_host, im_info, info_ld, boxes, scores, labels, count, run_end, nullptr, 0,
                               stream);
}

extern "C" int bdet_select_decode_ws(const float* const* anchors_host, const float* const* deltas_host, const int* n_l_host,
                                     const int* hw_host, int L, int B, int k, int div, int coder, int label_mode,
                                     const int* topk_idx, const float* topk_val, const int* topk_cnt,
                                     const float* mean_host, const float* std_host, const float* im_info, int info_ld,
                                     float* boxes, float* scores, void* labels, int* count, int* run_end, void* workspace,
                                     size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(L >= 1 && L <= BDET_MAX_LEVELS && B >= 0 && B <= 65535 && k >= 0 && div >= 1, "bad sizes");
  BDET_REQUIRE(coder == 0 || coder == 1, "coder must be 0 (BoxCoder) or 1 (PointCoder)");
  BDET_REQUIRE(label_mode == 0 || label_mode == 1, "label_mode must be 0 (idx % div) or 1 (level id)");
  if (B == 0) return BDET_OK;
  BDET_REQUIRE(count, "null count");
  cudaStream_t st = as_stream(stream);
  if (k == 0) {
    BDET_CUDA(cudaMemsetAsync(count, 0, (size_t)B * 4, st));
    if (run_end) BDET_CUDA(cudaMemsetAsync(run_end, 0, (size_t)B * L * 4, st));
    return BDET_OK;
  }
  BDET_REQUIRE(anchors_host && deltas_host && n_l_host && topk_idx && topk_val && topk_cnt && boxes && scores && labels,
               "null argument");
  BDET_REQUIRE(aligned16(boxes), "boxes must be 16-byte aligned");
  BDET_REQUIRE(!im_info || info_ld >= 2, "im_info rows need at least [h, w]");
  SelectArgs a;
  for (int l = 0; l < L; ++l) {
    const int hw = hw_host ? hw_host[l] : 0;
    BDET_REQUIRE(anchors_host[l] && deltas_host[l] && (hw > 0 || aligned16(deltas_host[l])), "null / unaligned level pointer");
    BDET_REQUIRE(hw >= 0 && (hw == 0 || n_l_host[l] % hw == 0), "NCHW level: n_l must be H*W*A");
    a.anchors[l] = anchors_host[l];
    a.deltas[l] = deltas_host[l];
    a.n_l[l] = n_l_host[l];
    a.hw[l] = hw;
  }
  a.L = L;
  a.B = B;
  a.k = k;
  a.div = div;
  a.coder = coder;
  a.label_mode = label_mode;
  a.filter = im_info != nullptr;
  a.topk_idx = topk_idx;
  a.topk_val = topk_val;
  a.topk_cnt = topk_cnt;
  a.im_info = im_info;
  a.info_ld = info_ld;
  for (int i = 0; i < 4; ++i) {
    a.mean.v[i] = mean_host ? mean_host[i] : 0.f;
    a.stdv.v[i] = std_host ? std_host[i] : 1.f;
  }
  a.boxes = boxes;
  a.scores = scores;
  a.labels = labels;
  a.count = count;
  a.run_end = run_end;
  a.flagw = nullptr;
  a.blkcnt = nullptr;
  if (a.filter && workspace) {
    const size_t need = bdet_select_decode_workspace(L, B, k);
    if (workspace_bytes < need) return set_error(BDET_EWORKSPACE, "bdet_select_decode_ws: workspace needs %zu bytes", need);
    BDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 3u) == 0, "workspace must be 4-byte aligned");
    const int nblk = ceil_div(k, 256);
    a.flagw = reinterpret_cast<uint32_t*>(workspace);
    a.blkcnt = reinterpret_cast<int*>(a.flagw + (size_t)B * L * nblk * 8);
    const dim3 grid(nblk, L, B);
    BDET_KERNEL("select_decode_kernel", st, select_flags_kernel<<<grid, 256, 0, st>>>(a));
    BDET_KERNEL("select_decode_kernel", st, select_write_kernel<<<grid, 256, 0, st>>>(a));
  } else if (a.filter)
    BDET_KERNEL("select_decode_kernel", st, select_decode_kernel<<<B, kDetThreads, 0, st>>>(a));
  else
    BDET_KERNEL("select_decode_kernel", st, select_decode_flat_kernel<<<dim3(ceil_div(k, 256), L, B), 256, 0, st>>>(a));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_finalize_detections(const float* boxes, const float* scores, const void* labels, int labels_is_float, int N,
                                        const int* keep, int keep_ld, const int* keep_count, const float* im_info, int info_ld,
                                        int B, int max_out, int mode, float* out, bdet_stream_t stream) {
  BDET_REQUIRE(B >= 0 && max_out >= 0 && N >= 0 && (mode == 0 || mode == 1), "bad arguments");
  if (B == 0 || max_out == 0) return BDET_OK;
  BDET_REQUIRE(boxes && keep && keep_count && out && (mode == 1 || (scores && labels)), "null argument");
  BDET_REQUIRE(aligned16(boxes), "boxes must be 16-byte aligned");
  BDET_REQUIRE(!im_info || info_ld >= 4, "im_info rows need [h, w, orig_h, orig_w]");
  BDET_REQUIRE(B <= 65535, "B > 65535");
  FinalArgs a{boxes, scores, labels, keep, keep_count, im_info, info_ld, N, keep_ld, max_out, mode, labels_is_float, B, out};
  BDET_KERNEL("finalize_kernel", as_stream(stream),
              finalize_kernel<<<dim3(ceil_div(max_out, 256), B), 256, 0, as_stream(stream)>>>(a));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" size_t bdet_dense_tail_smem(int L, int k, int max_out) {
  const size_t cap = (size_t)L * k;
  return cap * (16 + 4 + 4 + 4) + (size_t)max_out * (16 + 4 + 4);
}

extern "C" int bdet_dense_tail(const float* const* anchors_host, const float* const* deltas_host, const int* n_l_host,
                               const int* hw_host, int L, int B, int k, int div, int coder, const int* topk_idx,
                               const float* topk_val, const int* topk_cnt, const float* mean_host, const float* std_host,
                               const float* im_info, int info_ld, float iou_thresh, int max_out, float* dets, int* det_count,
                               bdet_stream_t stream) {
  BDET_REQUIRE(L >= 1 && L <= BDET_MAX_LEVELS && B >= 0 && B <= 65535 && k >= 1 && div >= 1 && max_out >= 1, "bad sizes");
  BDET_REQUIRE(coder == 0 || coder == 1, "coder must be 0 (BoxCoder) or 1 (PointCoder)");
  if (B == 0) return BDET_OK;
  BDET_REQUIRE(anchors_host && deltas_host && n_l_host && topk_idx && topk_val && topk_cnt && dets && det_count, "null argument");
  BDET_REQUIRE(!im_info || info_ld >= 4, "im_info rows need [h, w, orig_h, orig_w]");
  const size_t smem = bdet_dense_tail_smem(L, k, max_out);
  if (smem > 200 * 1024)
    return set_error(BDET_EUNSUPPORTED, "bdet_dense_tail: L * k = %d candidates do not fit shared memory (use the separate kernels)", L * k);
  TailArgs q;
  SelectArgs& a = q.s;
  for (int l = 0; l < L; ++l) {
    const int hw = hw_host ? hw_host[l] : 0;
    BDET_REQUIRE(anchors_host[l] && deltas_host[l] && (hw > 0 || aligned16(deltas_host[l])), "null / unaligned level pointer");
    BDET_REQUIRE(hw >= 0 && (hw == 0 || n_l_host[l] % hw == 0), "NCHW level: n_l must be H*W*A");
    a.anchors[l] = anchors_host[l];
    a.deltas[l] = deltas_host[l];
    a.n_l[l] = n_l_host[l];
    a.hw[l] = hw;
  }
  a.L = L;
  a.B = B;
  a.k = k;
  a.div = div;
  a.coder = coder;
  a.label_mode = 0;
  a.filter = 0;
  a.topk_idx = topk_idx;
  a.topk_val = topk_val;
  a.topk_cnt = topk_cnt;
  a.im_info = im_info;
  a.info_ld = info_ld;
  for (int i = 0; i < 4; ++i) {
    a.mean.v[i] = mean_host ? mean_host[i] : 0.f;
    a.stdv.v[i] = std_host ? std_host[i] : 1.f;
  }
  a.boxes = nullptr;
  a.scores = nullptr;
  a.labels = nullptr;
  a.count = nullptr;
  a.run_end = nullptr;
  a.flagw = nullptr;
  a.blkcnt = nullptr;
  q.thr = iou_thresh;
  q.max_out = max_out;
  q.dets = dets;
  q.det_count = det_count;
  cudaStream_t st = as_stream(stream);
  BDET_CUDA(cudaFuncSetAttribute(dense_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BDET_KERNEL("dense_tail_kernel", st, dense_tail_kernel<<<B, kDetThreads, smem, st>>>(q));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
