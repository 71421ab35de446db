// a13/a14: multi-level ROIAlign forward / backward and FPN level assignment.
// Reference: basedet/layers/common/roi_pool.py:12-78 -> F.nn.roi_align(mode="average", sample_points=2,
// aligned=True); MegDNN roi_align semantics restated in oracle ASSUMED-6 (zero padding for taps outside the map,
// lerp as a + (b - a) * t, average = sum / S^2).
//
// forward : one CTA per ROI.  The S*P sample coordinates per axis (floor index, fraction) are computed once
//           into shared memory; threads then walk the (channel, bin) outputs in memory order, so the 49-float
//           output rows are written fully coalesced and the 16 taps of a bin hit L1/L2 (the ROI footprint is a
//           few KB per channel).  All levels in one launch; output directly in the original ROI order.
// backward: (no workspace; the faster one, ~L2 atomic throughput) scatter form: per ROI and channel the separable
//           Wy / Wx weight tables give every footprint pixel its total weight, accumulated in registers and flushed
//           with ONE red.global.add.f32 per touched pixel;
//           (workspace given) atomics-free tile gather -- ROIs are binned per feature-map tile, one CTA accumulates its
//           tile x channel chunk in shared memory and writes every dfeat element exactly once: bit-deterministic.
#include <algorithm>

#include "roi_common.cuh"
#include "sortnet.cuh"

namespace bdet {

template <int TPH, int TPW, int TS>
__global__ void __launch_bounds__(kRoiThreads) roi_align_fwd_kernel(const RoiArgs p) {
  __shared__ SampleTab ty, tx;
  const int k = roi_of_cta(p, blockIdx.x), t = threadIdx.x;
  const int PH = TPH ? TPH : p.PH, PW = TPW ? TPW : p.PW, SH = TS ? TS : p.SH, SW = TS ? TS : p.SW;
  const RoiGeom g = roi_geom(p, k);
  if (!g.valid) {  // out-of-range batch / level index: defined as zeros (the reference would read out of bounds)
    float* out = p.out + (long long)k * p.C * PH * PW;
    for (int o = t; o < p.C * PH * PW; o += kRoiThreads) out[o] = 0.f;
    return;
  }
  fill_axis(ty, PH, SH, g.start_h, g.bin_h);
  fill_axis(tx, PW, SW, g.start_w, g.bin_w);
  __syncthreads();
  roi_fwd_direct<TPH, TPW, TS>(p, k, g, ty, tx, kRoiThreads);
}

// ---------------------------------------------------------------------------------------------------------------
// Backward, gather form (chosen by passing a workspace).  The feature maps are cut into kTH x kTW pixel tiles; ROIs are binned per tile
// (count -> scan -> fill); one CTA owns (tile, chunk of kCC channels), accumulates every ROI of its list into a
// shared-memory tile and writes each dfeat element exactly once: no global atomics, no memset, deterministic.
// Per pixel the contributing samples form a contiguous range along each axis (sample coordinates are monotone), so
//   dfeat[y][x] += sum_{sy in R(y)} sum_{sx in C(x)}  (dout[ph(sy)][pw(sx)] / S^2) * (wy(sy, y) * wx(sx, x))
// with the same products as the scatter form (oracle: g * ((1-ly)(1-lx)) ...); only the summation order differs.
constexpr int kTH = 16, kTW = 32, kCC = 32;
constexpr int kGatherThreads = 256;
constexpr int kMaxList = 384;  // (static shared memory: 3 CTAs of the gather kernel per SM)

struct TileGrid {
  int tiles_y[BDET_MAX_LEVELS], tiles_x[BDET_MAX_LEVELS];
  int base[BDET_MAX_LEVELS + 1];  // first tile id of each level; base[n_levels] = total
};

struct BinArgs {
  RoiArgs r;
  TileGrid g;
  int* count;   // (n_tiles)
  int* offset;  // (n_tiles + 1)
  int* fill;    // (n_tiles)
  int* list;    // (cap)
};

// footprint rows / cols of ROI k on its level (same fp32 expressions as fill_axis), clipped to the map
__device__ __forceinline__ bool roi_footprint(const RoiArgs& p, const RoiGeom& g, int& y_lo, int& y_hi, int& x_lo, int& x_hi) {
  const float fy0 = __fdiv_rn(0.5f, (float)p.SH), fy1 = __fdiv_rn((float)(p.SH - 1) + 0.5f, (float)p.SH);
  const float fx0 = __fdiv_rn(0.5f, (float)p.SW), fx1 = __fdiv_rn((float)(p.SW - 1) + 0.5f, (float)p.SW);
  const float ya = g.start_h + g.bin_h * (0.f + fy0), yb = g.start_h + g.bin_h * ((float)(p.PH - 1) + fy1);
  const float xa = g.start_w + g.bin_w * (0.f + fx0), xb = g.start_w + g.bin_w * ((float)(p.PW - 1) + fx1);
  y_lo = max((int)floorf(ya), 0);
  y_hi = min((int)floorf(yb) + 1, g.H - 1);
  x_lo = max((int)floorf(xa), 0);
  x_hi = min((int)floorf(xb) + 1, g.W - 1);
  return y_lo <= y_hi && x_lo <= x_hi;
}

template <bool FILL>
__global__ void __launch_bounds__(256) roi_bin_kernel(const BinArgs b) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= b.r.K) return;
  const RoiGeom g = roi_geom(b.r, k);
  if (!g.valid) return;
  int y_lo, y_hi, x_lo, x_hi;
  if (!roi_footprint(b.r, g, y_lo, y_hi, x_lo, x_hi)) return;
  const int tx_n = b.g.tiles_x[g.lvl], ty_n = b.g.tiles_y[g.lvl];
  const int t0 = b.g.base[g.lvl] + g.n * ty_n * tx_n;
  for (int ty = y_lo / kTH; ty <= y_hi / kTH; ++ty)
    for (int tx = x_lo / kTW; tx <= x_hi / kTW; ++tx) {
      const int tile = t0 + ty * tx_n + tx;
      if (FILL)
        b.list[b.offset[tile] + atomicAdd(&b.fill[tile], 1)] = k;
      else
        atomicAdd(&b.count[tile], 1);
    }
}

__global__ void __launch_bounds__(1024) roi_bin_scan_kernel(const BinArgs b) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n = b.g.base[b.r.lv.n_levels];
  if (t == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + t;
    const int v = i < n ? b.count[i] : 0;
    int incl = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= s) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int add = carry;
    for (int w = 0; w < warp; ++w) add += warp_tot[w];
    if (i < n) {
      b.offset[i] = add + incl - v;
      b.fill[i] = 0;
    }
    __syncthreads();
    if (t == 1023) carry = add + incl;
    __syncthreads();
  }
  if (t == 0) b.offset[n] = carry;
}

// Per-axis weight of one sample coordinate c on pixel `pix`: the bilinear taps are floor(c) and floor(c) + 1.
__device__ __forceinline__ float tap_weight(float c, int pix) {
  const float fl = floorf(c);
  const int i0 = (int)fl;
  const float fr = c - fl;
  return i0 == pix ? 1.f - fr : (i0 + 1 == pix ? fr : 0.f);
}

// FAST: 7 x 7 bins, 2 x 2 samples (every shipped config).  The ROI's contribution is computed as in the scatter kernel --
// a lane owns one tile column of one channel (narrow footprints pack 4 / 2 channels into a warp), T[ph] = sum_pw Wx[x][pw]
// dout[ph][pw] / 4 in registers, then one 7-term dot product with Wy[row][:] per footprint row -- and added to the
// shared-memory tile by its owner lane (no atomics: a warp works on its own channels, ROIs are visited one after another).
// Tables are padded to 8 floats per row (128-bit shared loads); the tile columns are XOR-swizzled by the channel so that
// the channels packed into a warp fall on different banks.
__device__ __forceinline__ int gather_col(int c, int x) { return x ^ ((c & 3) << 3); }

template <bool FAST>
__global__ void __launch_bounds__(kGatherThreads, FAST ? 3 : 1) roi_align_bwd_gather_kernel(const BinArgs b, int accumulate) {
  extern __shared__ __align__(16) float gsm[];
  const RoiArgs& p = b.r;
  const int PH = p.PH, PW = p.PW, bins = PH * PW;
  const int PHs = FAST ? 8 : PH, PWs = FAST ? 8 : PW;  // table row pitch
  float* acc = gsm;                          // kCC * kTH * kTW
  float* sdout = acc + kCC * kTH * kTW;      // kCC * PH * PWs  (dout / S^2 of the current ROI)
  float* wy = sdout + kCC * PH * PWs;        // kTH * PHs       Wy[r][ph] = sum of the bin's sample weights on row r
  float* wx = wy + kTH * PHs;                // kTW * PWs
  __shared__ int rng[4];                     // FAST: rows / columns of the tile the current ROI touches
  __shared__ int rlo[kTH], rhi[kTH], clo[kTW], chi[kTW];
  __shared__ int sids[kMaxList];
  const int tile = blockIdx.x, c0 = blockIdx.y * kCC, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nc = min(kCC, p.C - c0);
  int l = 0;
#pragma unroll
  for (int q = 1; q < BDET_MAX_LEVELS; ++q)
    if (q < p.lv.n_levels && tile >= b.g.base[q]) l = q;
  const int H = p.lv.H[l], W = p.lv.W[l];
  const int tx_n = b.g.tiles_x[l], ty_n = b.g.tiles_y[l];
  int rel = tile - b.g.base[l];
  const int n = rel / (ty_n * tx_n);
  rel -= n * ty_n * tx_n;
  const int y0t = (rel / tx_n) * kTH, x0t = (rel % tx_n) * kTW;
  const float cnt = (float)(p.SH * p.SW);
  for (int i = t; i < kCC * kTH * kTW; i += kGatherThreads) acc[i] = 0.f;
  const int beg = b.offset[tile], end = b.offset[tile + 1];
  // The bin lists are filled with atomics (arbitrary order); visiting ROIs in ascending id makes the fp32 sums
  // reproducible.  Rank sort in shared memory (ids are unique); longer lists stay in fill order.
  const int nlist = end - beg;
  const bool sorted = nlist <= kMaxList;
  if (sorted) {
    for (int i = t; i < nlist; i += kGatherThreads) {
      const int id = b.list[beg + i];
      int rank = 0;
      for (int j = 0; j < nlist; ++j) rank += b.list[beg + j] < id;
      sids[rank] = id;
    }
  }
  for (int e = beg; e < end; ++e) {
    __syncthreads();  // previous ROI's tables are free; (first iteration) acc zero-fill and sids are complete
    const int k = sorted ? sids[e - beg] : b.list[e];
    const RoiGeom g = roi_geom(p, k);
    if (FAST) {
      // warp 0: Wy of the 16 tile rows, warp 1: Wx of the 32 tile columns; zero outside the map / the ROI's footprint
      if (warp < 2) {
        const bool rows = warp == 0;
        const int pix = rows ? y0t + lane : x0t + lane;
        const bool live = rows ? (lane < kTH && pix < H) : pix < W;
        const float start = rows ? g.start_h : g.start_w, bin = rows ? g.bin_h : g.bin_w;
        float* tab = rows ? wy : wx;
        bool nz = false;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          float w = 0.f;
          if (live) w = tap_weight(start + bin * ((float)q + 0.25f), pix) + tap_weight(start + bin * ((float)q + 0.75f), pix);
          if (rows ? lane < kTH : true) tab[lane * 8 + q] = w;
          nz |= w != 0.f;
        }
        if (rows ? lane < kTH : true) tab[lane * 8 + 7] = 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) {
          rng[rows ? 0 : 2] = m ? __ffs(m) - 1 : 0;
          rng[rows ? 1 : 3] = m ? 31 - __clz(m) : -1;
        }
      }
      const float* dk = p.dout + ((long long)k * p.C + c0) * 49;
      for (int i = t; i < nc * 49; i += kGatherThreads) {
        const int c = i / 49, q = i - c * 49, ph = q / 7;
        sdout[c * 56 + ph * 8 + (q - ph * 7)] = __ldg(dk + i) * 0.25f;  // / (2 * 2) exactly
      }
      __syncthreads();
      const int fr0 = rng[0], fr1 = rng[1], fc0 = rng[2], fc1 = rng[3];
      if (fr1 < fr0 || fc1 < fc0) continue;  // CTA-uniform
      const int ncol = fc1 - fc0 + 1;
      const int lpc = ncol <= 8 ? 8 : (ncol <= 16 ? 16 : 32), G = 32 / lpc;
      const int sub = lane / lpc, xl = lane - sub * lpc, x = fc0 + xl;
      const bool xok = xl < ncol;
      float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
      if (xok) {
        wa = *reinterpret_cast<const float4*>(wx + x * 8);
        wb = *reinterpret_cast<const float4*>(wx + x * 8 + 4);
      }
      for (int cb = warp * G; cb < nc; cb += (kGatherThreads / 32) * G) {
        const int c = cb + sub;
        const bool ok = xok && c < nc;
        const float* dc = sdout + (c < nc ? c : nc - 1) * 56;
        float T[7];
#pragma unroll
        for (int ph = 0; ph < 7; ++ph) {
          const float4 a = *reinterpret_cast<const float4*>(dc + ph * 8);
          const float4 d2 = *reinterpret_cast<const float4*>(dc + ph * 8 + 4);
          float v = wa.x * a.x;
          v = __fmaf_rn(wa.y, a.y, v);
          v = __fmaf_rn(wa.z, a.z, v);
          v = __fmaf_rn(wa.w, a.w, v);
          v = __fmaf_rn(wb.x, d2.x, v);
          v = __fmaf_rn(wb.y, d2.y, v);
          v = __fmaf_rn(wb.z, d2.z, v);
          T[ph] = v;
        }
        float* ac = acc + c * (kTH * kTW) + gather_col(c, x);
        for (int r = fr0; r <= fr1; ++r) {
          const float4 u = *reinterpret_cast<const float4*>(wy + r * 8);
          const float4 w2 = *reinterpret_cast<const float4*>(wy + r * 8 + 4);
          float sum = u.x * T[0];
          sum = __fmaf_rn(u.y, T[1], sum);
          sum = __fmaf_rn(u.z, T[2], sum);
          sum = __fmaf_rn(u.w, T[3], sum);
          sum = __fmaf_rn(w2.x, T[4], sum);
          sum = __fmaf_rn(w2.y, T[5], sum);
          sum = __fmaf_rn(w2.z, T[6], sum);
          if (ok) ac[r * kTW] += sum;
        }
      }
      continue;
    }
    // rows: thread r < kTH builds Wy[r][:] and its non-zero bin range; columns: threads 32 .. 32 + kTW
    if (t < kTH) {
      const int y = y0t + t;
      int lo = PH, hi = -1;
      for (int ph = 0; ph < PH; ++ph) {
        float w = 0.f;
        if (y < H)
          for (int iy = 0; iy < p.SH; ++iy)
            w += tap_weight(g.start_h + g.bin_h * ((float)ph + __fdiv_rn((float)iy + 0.5f, (float)p.SH)), y);
        wy[t * PH + ph] = w;
        if (w != 0.f) {
          lo = min(lo, ph);
          hi = ph;
        }
      }
      rlo[t] = lo;
      rhi[t] = hi;
    } else if (t >= 32 && t < 32 + kTW) {
      const int xx = t - 32, x = x0t + xx;
      int lo = PW, hi = -1;
      for (int pw = 0; pw < PW; ++pw) {
        float w = 0.f;
        if (x < W)
          for (int ix = 0; ix < p.SW; ++ix)
            w += tap_weight(g.start_w + g.bin_w * ((float)pw + __fdiv_rn((float)ix + 0.5f, (float)p.SW)), x);
        wx[xx * PW + pw] = w;
        if (w != 0.f) {
          lo = min(lo, pw);
          hi = pw;
        }
      }
      clo[xx] = lo;
      chi[xx] = hi;
    }
    const float* dk = p.dout + ((long long)k * p.C + c0) * bins;
    for (int i = t; i < nc * bins; i += kGatherThreads) sdout[i] = __fdiv_rn(__ldg(dk + i), cnt);
    __syncthreads();
    // footprint of this ROI inside the tile: rows / columns with a non-empty bin range (contiguous)
    int fr0 = kTH, fr1 = -1, fc0 = kTW, fc1 = -1;
    for (int r = 0; r < kTH; ++r)
      if (rlo[r] <= rhi[r]) {
        fr0 = min(fr0, r);
        fr1 = r;
      }
    for (int x = 0; x < kTW; ++x)
      if (clo[x] <= chi[x]) {
        fc0 = min(fc0, x);
        fc1 = x;
      }
    if (fr1 < fr0 || fc1 < fc0) continue;  // CTA-uniform (tables are shared)
    const int ncol = fc1 - fc0 + 1, nrow = fr1 - fr0 + 1;
    int wshift = 0;
    while ((1 << wshift) < ncol) ++wshift;          // lanes: x = lane & (2^wshift - 1), row sub-index = lane >> wshift
    const int rows_per_it = 32 >> wshift;
    const int lx = lane & ((1 << wshift) - 1), lr = lane >> wshift;
    const int x = fc0 + lx;
    const bool xok = lx < ncol;
    const int tlo = xok ? clo[x] : 1, thi = xok ? chi[x] : 0;
    for (int c = warp; c < nc; c += kGatherThreads / 32) {
      const float* dc = sdout + c * bins;
      float* ac = acc + c * kTH * kTW;
      for (int r0 = 0; r0 < nrow; r0 += rows_per_it) {
        const int r = fr0 + r0 + lr;
        if (r > fr1 || tlo > thi) continue;
        const int slo = rlo[r], shi = rhi[r];
        float sum = 0.f;
        for (int ph = slo; ph <= shi; ++ph) {
          const float wyv = wy[r * PH + ph];
          for (int pw = tlo; pw <= thi; ++pw) sum += dc[ph * PW + pw] * (wyv * wx[x * PW + pw]);
        }
        ac[r * kTW + x] += sum;  // one lane per (c, r, x): no conflict
      }
    }
  }
  __syncthreads();
  float* df = p.lv.dfeat[l] + ((long long)n * p.C + c0) * H * W;
  for (int i = t; i < nc * kTH * kTW; i += kGatherThreads) {
    const int c = i / (kTH * kTW), rem = i - c * (kTH * kTW);
    const int y = y0t + rem / kTW, x = x0t + rem % kTW;
    if (y < H && x < W) {
      float* o = df + ((long long)c * H + y) * W + x;
      const float v = FAST ? acc[c * (kTH * kTW) + (rem / kTW) * kTW + gather_col(c, rem % kTW)] : acc[i];
      *o = accumulate ? (*o + v) : v;
    }
  }
}

// Backward, scatter form (no workspace): one CTA per ROI.  The separable bilinear weights of the ROI are tabulated
// once per footprint chunk -- Wy[row][ph], Wx[col][pw] = summed tap weights of the bin's samples -- and shared by all
// C channels; a warp takes a channel, a lane one footprint column, and every (channel, y, x) of the footprint gets
// ONE red.global.add.f32 (coalesced along x): footprint x C reds per ROI instead of 16 x P^2 x C, no shared atomics,
// no zero / flush passes.  Accumulation order across ROIs is arbitrary (fp32, 1e-5 gate).
constexpr int kFpChunk = 64;  // footprint rows / cols tabulated at a time

__global__ void __launch_bounds__(kRoiThreads) roi_align_bwd_kernel(const RoiArgs p, unsigned tma_level_mask) {
  extern __shared__ __align__(16) float bsm[];
  const int PH = p.PH, PW = p.PW;
  const int PHs = (PH + 3) & ~3, PWs = (PW + 3) & ~3;  // table rows padded to 16 bytes (128-bit shared loads)
  float* wy = bsm;                              // kFpChunk * PHs
  float* wx = wy + kFpChunk * PHs;              // kFpChunk * PWs
  float* sd = wx + kFpChunk * PWs;              // (warps) * PH * PWs: dout / S^2 of the warp's current channel
  __shared__ int rlo[kFpChunk], rhi[kFpChunk], clo[kFpChunk], chi[kFpChunk];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int k = roi_of_cta(p, blockIdx.x);
  const int Cn = p.C, cbeg = 0;
  const RoiGeom g = roi_geom(p, k);
  if (!g.valid) return;
  if (tma_level_mask) {  // ROIs the TMA kernel (roi_tma.cu, launched just before) has taken
    FwdPlan unused;
    if (roi_bwd_takes_tma(p, g, tma_level_mask, &unused)) return;
  }
  int y_lo, y_hi, x_lo, x_hi;
  if (!roi_footprint(p, g, y_lo, y_hi, x_lo, x_hi)) return;  // ROI entirely outside the map
  const int H = g.H, W = g.W;
  const float cnt = (float)(p.SH * p.SW);
  const int icnt = p.SH * p.SW;
  const bool cnt_pow2 = (icnt & (icnt - 1)) == 0;  // x / 2^k == x * 2^-k exactly: skips the IEEE division routine
  const float inv_cnt = __fdiv_rn(1.f, cnt);
  float* dfeat = p.lv.dfeat[g.lvl] + ((long long)g.n * p.C + cbeg) * H * W;
  const float* dout = p.dout + ((long long)k * p.C + cbeg) * PH * PW;
  float* sdw = sd + warp * (PH == 7 && PW == 7 ? 4 * 56 : PH * PWs);  // 7 x 7: up to 4 packed channels per warp
  for (int fy = y_lo; fy <= y_hi; fy += kFpChunk) {
    for (int fx = x_lo; fx <= x_hi; fx += kFpChunk) {
      const int nr = min(kFpChunk, y_hi - fy + 1), ncol = min(kFpChunk, x_hi - fx + 1);
      __syncthreads();
      if (t < nr) {
        const int y = fy + t;
        int lo = PH, hi = -1;
        for (int ph = 0; ph < PHs; ++ph) {
          float w = 0.f;
          if (ph < PH)
            for (int iy = 0; iy < p.SH; ++iy)
              w += tap_weight(g.start_h + g.bin_h * ((float)ph + __fdiv_rn((float)iy + 0.5f, (float)p.SH)), y);
          wy[t * PHs + ph] = w;
          if (w != 0.f) {
            lo = min(lo, ph);
            hi = ph;
          }
        }
        rlo[t] = lo;
        rhi[t] = hi;
      } else if (t >= 128 && t < 128 + ncol) {
        const int xx = t - 128, x = fx + xx;
        int lo = PW, hi = -1;
        for (int pw = 0; pw < PWs; ++pw) {
          float w = 0.f;
          if (pw < PW)
            for (int ix = 0; ix < p.SW; ++ix)
              w += tap_weight(g.start_w + g.bin_w * ((float)pw + __fdiv_rn((float)ix + 0.5f, (float)p.SW)), x);
          wx[xx * PWs + pw] = w;
          if (w != 0.f) {
            lo = min(lo, pw);
            hi = pw;
          }
        }
        clo[xx] = lo;
        chi[xx] = hi;
      }
      __syncthreads();
      if (PH == 7 && PW == 7) {
        // 7 x 7 bins (every shipped config): two separable passes per channel, all in registers.  A lane owns one
        // footprint column: T[ph] = sum_pw Wx[x][pw] * dout[ph][pw], then per footprint row (warp-uniform weights,
        // broadcast loads) sum_ph Wy[y][ph] * T[ph] and one coalesced red.  Dense 7-term sums: weights outside a
        // bin's range are zero, so there is no data-dependent branch.
        // Narrow footprints pack G = 4 / 2 channels into a warp (8 / 16 lanes per channel), so that a 6-pixel-wide ROI
        // does not leave 26 lanes idle; G consecutive channels are 49 * G contiguous floats of dout.
        const int lpc = ncol <= 8 ? 8 : (ncol <= 16 ? 16 : 32);
        const int G = 32 / lpc;
        const int sub = lane / lpc, xl = lane - sub * lpc;
        const int nd = 49 * G;                                   // dout floats per warp pass
        float* sdg = sdw + sub * 56;
        // where this lane's j-th dout float of a pass goes in the padded table (the same for every pass): -1 = none
        int dofs[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const int i = lane + 32 * j;
          const int ch = i / 49, q = i - ch * 49;
          dofs[j] = i < nd ? ch * 56 + (q / 7) * 8 + (q % 7) : -1;
        }
        for (int x0 = 0; x0 < ncol; x0 += 32) {
          const int xx = x0 + xl;
          const bool xok = xx < ncol;
          float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
          if (xok) {
            wa = *reinterpret_cast<const float4*>(wx + xx * 8);
            wb = *reinterpret_cast<const float4*>(wx + xx * 8 + 4);
          }
          // dout of the next channel group is fetched while the current one is processed (the load is the only long
          // latency of the loop)
          float dr[7];
          auto fetch = [&](int c) {
            const float* src = dout + (long long)c * 49;
            const int n = min(nd, (Cn - c) * 49);
#pragma unroll
            for (int j = 0; j < 7; ++j) {
              const int i = lane + 32 * j;
              dr[j] = (j * 32 < nd && i < n) ? __ldg(src + i) : 0.f;
            }
          };
          if (warp * G < Cn) fetch(warp * G);
          for (int c = warp * G; c < Cn; c += (kRoiThreads / 32) * G) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 7; ++j)
              if (dofs[j] >= 0) sdw[dofs[j]] = cnt_pow2 ? dr[j] * inv_cnt : __fdiv_rn(dr[j], cnt);
            __syncwarp();
            const int cn = c + (kRoiThreads / 32) * G;
            if (cn < Cn) fetch(cn);
            float T[7];
#pragma unroll
            for (int ph = 0; ph < 7; ++ph) {
              const float4 a = *reinterpret_cast<const float4*>(sdg + ph * 8);
              const float4 b = *reinterpret_cast<const float4*>(sdg + ph * 8 + 4);
              float v = wa.x * a.x;
              v = __fmaf_rn(wa.y, a.y, v);
              v = __fmaf_rn(wa.z, a.z, v);
              v = __fmaf_rn(wa.w, a.w, v);
              v = __fmaf_rn(wb.x, b.x, v);
              v = __fmaf_rn(wb.y, b.y, v);
              v = __fmaf_rn(wb.z, b.z, v);
              T[ph] = v;
            }
            const bool ok = xok && c + sub < Cn;
            float* gp = dfeat + (long long)(c + sub) * H * W + (long long)fy * W + (fx + xx);
#pragma unroll 4
            for (int r = 0; r < nr; ++r, gp += W) {
              const float4 u = *reinterpret_cast<const float4*>(wy + r * 8);
              const float4 w2 = *reinterpret_cast<const float4*>(wy + r * 8 + 4);
              float sum = u.x * T[0];
              sum = __fmaf_rn(u.y, T[1], sum);
              sum = __fmaf_rn(u.z, T[2], sum);
              sum = __fmaf_rn(u.w, T[3], sum);
              sum = __fmaf_rn(w2.x, T[4], sum);
              sum = __fmaf_rn(w2.y, T[5], sum);
              sum = __fmaf_rn(w2.z, T[6], sum);
              red_add_if(gp, sum, ok && sum != 0.f);  // predicated: no divergent branch around the reduction
            }
          }
        }
        continue;
      }
      const int wcols = min(ncol, 32);
      int wshift = 0;
      while ((1 << wshift) < wcols) ++wshift;  // lanes: x = lane & (2^wshift - 1), row sub-index = lane >> wshift
      const int rows_per_it = 32 >> wshift;
      const int lx = lane & ((1 << wshift) - 1), lr = lane >> wshift;
      for (int c = warp; c < Cn; c += kRoiThreads / 32) {
        __syncwarp();
        for (int i = lane; i < PH * PW; i += 32)
          sdw[(i / PW) * PWs + (i % PW)] = __fdiv_rn(__ldg(dout + (long long)c * PH * PW + i), cnt);
        __syncwarp();
        float* gc = dfeat + (long long)c * H * W;
        for (int x0 = 0; x0 < ncol; x0 += 32) {
          const int xx = x0 + lx;
          const bool xok = lx < wcols && xx < ncol;
          const int tlo = xok ? clo[xx] : 1, thi = xok ? chi[xx] : 0;
          if (tlo > thi) continue;  // lane idle for this column block (no warp-level sync inside)
          for (int r0 = 0; r0 < nr; r0 += rows_per_it) {
            const int r = r0 + lr;
            if (r >= nr) continue;
            const int slo = rlo[r], shi = rhi[r];
            float sum = 0.f;
            for (int ph = slo; ph <= shi; ++ph) {
              const float wyv = wy[r * PHs + ph];
              for (int pw = tlo; pw <= thi; ++pw) sum += sdw[ph * PWs + pw] * (wyv * wx[xx * PWs + pw]);
            }
            if (sum != 0.f) atomicAdd(gc + (long long)(fy + r) * W + (fx + xx), sum);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Max ROI pooling, layers/common/roi_pool.py:62-63 -> F.nn.roi_pooling(mode="max", scale) (MegDNN ROIPooling, the Caffe
// rule, oracle ASSUMED-13, pinned by the reference's known-answer test tests/layers/test_roi_pool.py:48-61): roi corners
// rounded to pixels, size = end - start + 1 (>= 1), bin [floor(p * size / P), ceil((p + 1) * size / P)) clipped to the
// map, maximum over the bin (0 for an empty bin); the argmax is kept for the backward.
__global__ void __launch_bounds__(256) roi_maxpool_fwd_kernel(const RoiArgs p, int* __restrict__ argmax) {
  const long long total = (long long)p.K * p.C * p.PH * p.PW;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(o % p.PW), ph = (int)((o / p.PW) % p.PH);
    const int c = (int)((o / ((long long)p.PW * p.PH)) % p.C), k = (int)(o / ((long long)p.PW * p.PH * p.C));
    const float* r = p.rois + (long long)k * 5;
    const int n = (int)__ldg(r);
    const int lvl = p.levels ? __ldg(p.levels + k) : 0;
    float best = 0.f;
    int bi = -1;
    if (n >= 0 && n < p.B && lvl >= 0 && lvl < p.lv.n_levels) {
      const int H = p.lv.H[lvl], W = p.lv.W[lvl];
      const float sc = p.lv.scale[lvl];
      const int x0 = (int)roundf(__ldg(r + 1) * sc), y0 = (int)roundf(__ldg(r + 2) * sc);
      const int x1 = (int)roundf(__ldg(r + 3) * sc), y1 = (int)roundf(__ldg(r + 4) * sc);
      const int rw = max(x1 - x0 + 1, 1), rh = max(y1 - y0 + 1, 1);
      const float bh = __fdiv_rn((float)rh, (float)p.PH), bw = __fdiv_rn((float)rw, (float)p.PW);
      int hs = (int)floorf((float)ph * bh), he = (int)ceilf((float)(ph + 1) * bh);
      int ws = (int)floorf((float)pw * bw), we = (int)ceilf((float)(pw + 1) * bw);
      hs = min(max(hs + y0, 0), H);
      he = min(max(he + y0, 0), H);
      ws = min(max(ws + x0, 0), W);
      we = min(max(we + x0, 0), W);
      if (he > hs && we > ws) {
        const float* f = p.lv.feat[lvl] + ((long long)n * p.C + c) * H * W;
        best = -3.402823466e+38f;
        for (int y = hs; y < he; ++y)
          for (int x = ws; x < we; ++x) {
            const float v = __ldg(f + y * W + x);
            if (v > best) {
              best = v;
              bi = y * W + x;
            }
          }
      }
    }
    p.out[o] = best;
    if (argmax) argmax[o] = bi;
  }
}

__global__ void __launch_bounds__(256) roi_maxpool_bwd_kernel(const RoiArgs p, const int* __restrict__ argmax) {
  const long long total = (long long)p.K * p.C * p.PH * p.PW;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int a = argmax[o];
    if (a < 0) continue;
    const int c = (int)((o / ((long long)p.PW * p.PH)) % p.C), k = (int)(o / ((long long)p.PW * p.PH * p.C));
    const int n = (int)__ldg(p.rois + (long long)k * 5);
    const int lvl = p.levels ? __ldg(p.levels + k) : 0;
    if (n < 0 || n >= p.B || lvl < 0 || lvl >= p.lv.n_levels) continue;
    atomicAdd(p.lv.dfeat[lvl] + ((long long)n * p.C + c) * p.lv.H[lvl] * p.lv.W[lvl] + a, __ldg(p.dout + o));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Processing order.  CTAs that run at the same time should touch the same part of the pyramid: a ROI reads (updates) its
// footprint in all C planes of its level, and one image's planes (91 MB in config 3) do not stay in L2 when the ROIs of
// the image arrive in random spatial order (ncu: the forward read 2.26 GB from DRAM for 0.9 GB of distinct footprints).
// CTA b sorts the ROIs of image b (CTA B: the ROIs no kernel processes) by (level, 32 x 32-pixel tile row, tile column, index)
// and writes them behind the ROIs of the images before it; the ROI kernels walk that permutation (outputs stay in the
// original ROI order).  Every CTA scans all K batch indices (20 B stride, L2 hits), so no count / scan kernel precedes it.
constexpr int kOrderMax = 16384;
__global__ void __launch_bounds__(1024) roi_order_kernel(const RoiArgs p, int* __restrict__ perm) {
  extern __shared__ __align__(16) unsigned char oraw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(oraw);
  __shared__ int s_before, s_mine;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) s_before = s_mine = 0;
  __syncthreads();
  int before = 0;
  for (int i = threadIdx.x; i < p.K; i += blockDim.x) {
    const RoiGeom g = roi_geom(p, i);
    const int img = g.valid ? g.n : p.B;
    before += img < b;
    if (img == b) {
      const float cy = g.start_h + 0.5f * g.bin_h * (float)p.PH, cx = g.start_w + 0.5f * g.bin_w * (float)p.PW;
      const uint32_t ty = (uint32_t)min(max((int)(cy * (1.f / 32.f)), 0), 1023);
      const uint32_t tx = (uint32_t)min(max((int)(cx * (1.f / 32.f)), 0), 1023);
      const uint32_t l = g.valid ? (uint32_t)g.lvl & 7u : 0u;
      keys[atomicAdd(&s_mine, 1)] = (((uint64_t)l << 10 | ty) << 10 | tx) << 32 | (uint32_t)i;
    }
  }
  for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if ((threadIdx.x & 31) == 0 && before) atomicAdd(&s_before, before);
  __syncthreads();
  const int mine = s_mine;
  if (mine == 0) return;
  int P = 2;
  while (P < mine) P <<= 1;
  for (int i = mine + threadIdx.x; i < P; i += blockDim.x) keys[i] = ~0ull;
  __syncthreads();
  bitonic_sort_smem(keys, P);
  for (int i = threadIdx.x; i < mine; i += blockDim.x) perm[s_before + i] = (int)(uint32_t)keys[i];
}

static void make_tile_grid(TileGrid* g, int n_levels, const int* hw, int B, int* max_per_image) {
  int base = 0, mx = 1;
  for (int l = 0; l < n_levels; ++l) {
    g->tiles_y[l] = (hw[2 * l] + kTH - 1) / kTH;
    g->tiles_x[l] = (hw[2 * l + 1] + kTW - 1) / kTW;
    g->base[l] = base;
    base += B * g->tiles_y[l] * g->tiles_x[l];
    mx = max(mx, g->tiles_y[l] * g->tiles_x[l]);
  }
  for (int l = n_levels; l <= BDET_MAX_LEVELS; ++l) g->base[l] = base;
  if (max_per_image) *max_per_image = mx;
}

// assign_rois, roi_pool.py:19-25: clamp(floor(4 + log(sqrt(area) / 224) / ln 2), lo, hi) - lo
__global__ void __launch_bounds__(256) roi_assign_levels_kernel(const float* __restrict__ rois, int K, int lo, int hi, float ln2,
                                                                int* __restrict__ levels) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float* r = rois + (long long)k * 5;
  float area = (__ldg(r + 3) - __ldg(r + 1)) * (__ldg(r + 4) - __ldg(r + 2));
  float v = floorf(4.f + __fdiv_rn(logf(__fdiv_rn(sqrtf(area), 224.f)), ln2));
  // float -> int32 of NaN / -inf on the reference's x86 path is INT_MIN; the clamp then yields `lo`
  int l = (v == v && v > -2.0e9f) ? (v < 2.0e9f ? (int)v : 0x7fffffff) : (int)0x80000000;
  l = max(min(l, hi), lo);
  levels[k] = l - lo;
}

static int fill_roi_args(RoiArgs* a, const void* const* feats, bool bwd, int n_levels, const int* hw, const float* scale, int B,
                         int C, const float* rois, const int* levels, int K, int PH, int PW, int sh, int sw, int aligned) {
  if (n_levels < 1 || n_levels > BDET_MAX_LEVELS) return set_error(BDET_EINVAL, "roi_align: n_levels must be in [1, %d]", BDET_MAX_LEVELS);
  if (PH < 1 || PW < 1 || sh < 1 || sw < 1) return set_error(BDET_EINVAL, "roi_align: pool shape / sample points must be >= 1");
  if (PH * sh > kMaxSamples || PW * sw > kMaxSamples) return set_error(BDET_EUNSUPPORTED, "roi_align: more than %d samples per axis", kMaxSamples);
  if (B < 0 || C < 0 || K < 0) return set_error(BDET_EINVAL, "roi_align: negative size");
  a->lv.n_levels = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    if (!feats[l]) return set_error(BDET_EINVAL, "roi_align: null feature pointer");
    if (bwd) {
      a->lv.dfeat[l] = reinterpret_cast<float*>(const_cast<void*>(feats[l]));
      a->lv.feat[l] = nullptr;
    } else {
      a->lv.feat[l] = reinterpret_cast<const float*>(feats[l]);
      a->lv.dfeat[l] = nullptr;
    }
    a->lv.H[l] = hw[2 * l];
    a->lv.W[l] = hw[2 * l + 1];
    a->lv.scale[l] = scale[l];
    if (a->lv.H[l] < 1 || a->lv.W[l] < 1) return set_error(BDET_EINVAL, "roi_align: empty feature map");
  }
  a->rois = rois;
  a->levels = levels;
  a->B = B;
  a->C = C;
  a->K = K;
  a->PH = PH;
  a->PW = PW;
  a->SH = sh;
  a->SW = sw;
  a->offset = aligned ? 0.5f : 0.f;
  a->dout = nullptr;
  a->out = nullptr;
  a->bwd_cap = 0;
  a->perm = nullptr;
  return BDET_OK;
}

}  // namespace bdet

using namespace bdet;

extern "C" int bdet_roi_assign_levels(const float* rois, int K, int min_level, int max_level, int* levels,
                                      bdet_stream_t stream) {
  BDET_REQUIRE(K >= 0 && min_level <= max_level, "bad arguments");
  if (K == 0) return BDET_OK;
  BDET_REQUIRE(rois && levels, "null argument");
  BDET_KERNEL("roi_assign_levels_kernel", as_stream(stream), roi_assign_levels_kernel<<<ceil_div(K, 256), 256, 0, as_stream(stream)>>>(rois, K, min_level, max_level, (float)0.6931471805599453,
                                                                          levels));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_roi_align_fwd(const float* const* feats_host, int n_levels, const int* hw_host,
                                  const float* scale_host, int B, int C, const float* rois, const int* levels, int K,
                                  int PH, int PW, int sample_h, int sample_w, int aligned, float* out,
                                  bdet_stream_t stream) {
  return bdet_roi_align_fwd_perm(feats_host, n_levels, hw_host, scale_host, B, C, rois, levels, K, PH, PW, sample_h, sample_w,
                                 aligned, out, nullptr, stream);
}

extern "C" int bdet_roi_order(int n_levels, const int* hw_host, const float* scale_host, int B, const float* rois,
                              const int* levels, int K, int PH, int PW, int aligned, int* perm, bdet_stream_t stream) {
  BDET_REQUIRE(hw_host && scale_host && K >= 0, "bad arguments");
  if (K == 0) return BDET_OK;
  BDET_REQUIRE(rois && perm, "null argument");
  if (K > kOrderMax) return set_error(BDET_EUNSUPPORTED, "bdet_roi_order: more than %d rois", kOrderMax);
  RoiArgs a;
  const void* dummy[BDET_MAX_LEVELS];
  for (int l = 0; l < BDET_MAX_LEVELS; ++l) dummy[l] = rois;  // geometry only: the feature pointers are not touched
  int rc = fill_roi_args(&a, dummy, false, n_levels, hw_host, scale_host, B, 1, rois, levels, K, PH, PW, 1, 1, aligned);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const int P = next_pow2(K < 2 ? 2 : K);
  BDET_CUDA(cudaFuncSetAttribute(roi_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kOrderMax * 8));
  BDET_KERNEL("roi_order_kernel", st, roi_order_kernel<<<B + 1, 1024, (size_t)P * 8, st>>>(a, perm));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_roi_align_fwd_perm(const float* const* feats_host, int n_levels, const int* hw_host,
                                       const float* scale_host, int B, int C, const float* rois, const int* levels, int K,
                                       int PH, int PW, int sample_h, int sample_w, int aligned, float* out, const int* perm,
                                       bdet_stream_t stream) {
  BDET_REQUIRE(feats_host && hw_host && scale_host, "null argument");
  RoiArgs a;
  int rc = fill_roi_args(&a, reinterpret_cast<const void* const*>(feats_host), false, n_levels, hw_host, scale_host, B, C, rois,
                         levels, K, PH, PW, sample_h, sample_w, aligned);
  if (rc) return rc;
  if (K == 0 || C == 0) return BDET_OK;
  BDET_REQUIRE(rois && out, "null argument");
  a.out = out;
  cudaStream_t st = as_stream(stream);
  a.perm = perm;
  rc = roi_fwd_tma_launch(a, st);  // TMA kernel (roi_tma.cu) when the shape / alignment qualifies
  if (rc < 0) return rc;
  if (rc == 1) {
    BDET_LAUNCH_CHECK();
    return BDET_OK;
  }
  if (PH == 7 && PW == 7 && sample_h == 2 && sample_w == 2)
    BDET_KERNEL("roi_align_fwd_kernel", st, roi_align_fwd_kernel<7, 7, 2><<<K, kRoiThreads, 0, st>>>(a));
  else
    BDET_KERNEL("roi_align_fwd_kernel", st, roi_align_fwd_kernel<0, 0, 0><<<K, kRoiThreads, 0, st>>>(a));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_roi_maxpool_fwd(const float* const* feats_host, int n_levels, const int* hw_host, const float* scale_host,
                                    int B, int C, const float* rois, const int* levels, int K, int PH, int PW, float* out,
                                    int* argmax, bdet_stream_t stream) {
  BDET_REQUIRE(feats_host && hw_host && scale_host, "null argument");
  RoiArgs a;
  int rc = fill_roi_args(&a, reinterpret_cast<const void* const*>(feats_host), false, n_levels, hw_host, scale_host, B, C, rois,
                         levels, K, PH, PW, 1, 1, 0);
  if (rc) return rc;
  if (K == 0 || C == 0) return BDET_OK;
  BDET_REQUIRE(rois && out, "null argument");
  a.out = out;
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)K * C * PH * PW;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
  BDET_KERNEL("roi_maxpool_fwd_kernel", st, roi_maxpool_fwd_kernel<<<grid, 256, 0, st>>>(a, argmax));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_roi_maxpool_bwd(float* const* dfeats_host, int n_levels, const int* hw_host, int B, int C, const float* rois,
                                    const int* levels, int K, int PH, int PW, const float* dout, const int* argmax,
                                    int accumulate, bdet_stream_t stream) {
  BDET_REQUIRE(dfeats_host && hw_host, "null argument");
  float ones[BDET_MAX_LEVELS];
  for (int l = 0; l < BDET_MAX_LEVELS; ++l) ones[l] = 1.f;
  RoiArgs a;
  int rc = fill_roi_args(&a, reinterpret_cast<const void* const*>(dfeats_host), true, n_levels, hw_host, ones, B, C, rois, levels,
                         K, PH, PW, 1, 1, 0);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (B == 0 || C == 0) return BDET_OK;
  if (!accumulate)
    for (int l = 0; l < n_levels; ++l)
      BDET_CUDA(cudaMemsetAsync(dfeats_host[l], 0, (size_t)B * C * hw_host[2 * l] * hw_host[2 * l + 1] * 4, st));
  if (K == 0) return BDET_OK;
  BDET_REQUIRE(rois && dout && argmax, "null argument");
  a.dout = dout;
  const long long total = (long long)K * C * PH * PW;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
  BDET_KERNEL("roi_maxpool_bwd_kernel", st, roi_maxpool_bwd_kernel<<<grid, 256, 0, st>>>(a, argmax));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" size_t bdet_roi_align_bwd_workspace(int n_levels, const int* hw_host, int B, int K) {
  if (n_levels < 1 || n_levels > BDET_MAX_LEVELS || !hw_host || B <= 0 || K <= 0) return 16;
  TileGrid g;
  int mx = 1;
  make_tile_grid(&g, n_levels, hw_host, B, &mx);
  const size_t n_tiles = (size_t)g.base[n_levels];
  // a ROI can touch at most every tile of its (image, level) map
  return align_up((n_tiles + 1) * 4, 256) * 3 + align_up((size_t)K * mx * 4, 256) + 256;
}

extern "C" int bdet_roi_align_bwd(float* const* dfeats_host, int n_levels, const int* hw_host, const float* scale_host,
                                  int B, int C, const float* rois, const int* levels, int K, int PH, int PW,
                                  int sample_h, int sample_w, int aligned, const float* dout, int accumulate,
                                  void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  return bdet_roi_align_bwd_perm(dfeats_host, n_levels, hw_host, scale_host, B, C, rois, levels, K, PH, PW, sample_h, sample_w,
                                 aligned, dout, accumulate, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int bdet_roi_align_bwd_perm(float* const* dfeats_host, int n_levels, const int* hw_host, const float* scale_host,
                                       int B, int C, const float* rois, const int* levels, int K, int PH, int PW,
                                       int sample_h, int sample_w, int aligned, const float* dout, int accumulate,
                                       const int* perm, void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(dfeats_host && hw_host && scale_host, "null argument");
  RoiArgs a;
  int rc = fill_roi_args(&a, reinterpret_cast<const void* const*>(dfeats_host), true, n_levels, hw_host, scale_host, B, C, rois,
                         levels, K, PH, PW, sample_h, sample_w, aligned);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (B == 0 || C == 0) return BDET_OK;
  BDET_REQUIRE(K == 0 || (rois && dout), "null argument");
  a.dout = dout;
  a.perm = perm;
  if (workspace) {
    // gather form: every dfeat element is written once (zero where no ROI reaches), no atomics
    const size_t need = bdet_roi_align_bwd_workspace(n_levels, hw_host, B, K > 0 ? K : 1);
    if (workspace_bytes < need) return set_error(BDET_EWORKSPACE, "bdet_roi_align_bwd: workspace needs %zu bytes", need);
    BDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 3u) == 0, "workspace must be 4-byte aligned");
    BinArgs b;
    b.r = a;
    make_tile_grid(&b.g, n_levels, hw_host, B, nullptr);
    const int n_tiles = b.g.base[n_levels];
    char* w = reinterpret_cast<char*>(workspace);
    const size_t seg = align_up(((size_t)n_tiles + 1) * 4, 256);
    b.count = reinterpret_cast<int*>(w);
    b.offset = reinterpret_cast<int*>(w + seg);
    b.fill = reinterpret_cast<int*>(w + 2 * seg);
    b.list = reinterpret_cast<int*>(w + 3 * seg);
    BDET_CUDA(cudaMemsetAsync(b.count, 0, (size_t)n_tiles * 4, st));
    if (K > 0) BDET_KERNEL("roi_bin_kernel", st, roi_bin_kernel<false><<<ceil_div(K, 256), 256, 0, st>>>(b));
    BDET_KERNEL("roi_bin_scan_kernel", st, roi_bin_scan_kernel<<<1, 1024, 0, st>>>(b));
    if (K > 0) BDET_KERNEL("roi_bin_kernel", st, roi_bin_kernel<true><<<ceil_div(K, 256), 256, 0, st>>>(b));
    const bool fast = PH == 7 && PW == 7 && sample_h == 2 && sample_w == 2;
    const int PHs = fast ? 8 : PH, PWs = fast ? 8 : PW;
    const size_t smem = ((size_t)kCC * kTH * kTW + (size_t)kCC * PH * PWs + (size_t)kTH * PHs + (size_t)kTW * PWs) * 4;
    if (smem > 200 * 1024) return set_error(BDET_EUNSUPPORTED, "bdet_roi_align_bwd: pool shape too large for the gather kernel");
    auto kern = fast ? roi_align_bwd_gather_kernel<true> : roi_align_bwd_gather_kernel<false>;
    BDET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chunks = ceil_div(C, kCC);
    if (chunks > 65535) return set_error(BDET_EUNSUPPORTED, "bdet_roi_align_bwd: too many channels");
    BDET_KERNEL("roi_align_bwd_gather_kernel", st, kern<<<dim3(n_tiles, chunks), kGatherThreads, smem, st>>>(b, accumulate));
    BDET_LAUNCH_CHECK();
    return BDET_OK;
  }
  // scatter form (no workspace): shared-memory footprint accumulation + red.global.add flush
  if (!accumulate) {
    for (int l = 0; l < n_levels; ++l)
      BDET_CUDA(cudaMemsetAsync(dfeats_host[l], 0, (size_t)B * C * hw_host[2 * l] * hw_host[2 * l + 1] * 4, st));
  }
  if (K == 0) return BDET_OK;
  const int PHs = (PH + 3) & ~3, PWs = (PW + 3) & ~3;
  const size_t smem = ((size_t)kFpChunk * (PHs + PWs) + (size_t)(kRoiThreads / 32) * (PH == 7 && PW == 7 ? 4 * 56 : PH * PWs)) * 4;
  if (smem > 200 * 1024) return set_error(BDET_EUNSUPPORTED, "bdet_roi_align_bwd: pool shape too large");
  if (smem > 40 * 1024)
    BDET_CUDA(cudaFuncSetAttribute(roi_align_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned tma_mask = 0;
  rc = roi_bwd_tma_launch(a, st, &tma_mask);  // smem accumulation + cp.reduce.async.bulk.tensor (roi_tma.cu)
  if (rc < 0) return rc;
  // the direct scatter kernel takes what is left (levels whose rows are not 16-byte multiples, oversized footprints);
  // when every level has tensor maps it only finds work for ROIs wider than 64 feature pixels
  BDET_KERNEL("roi_align_bwd_kernel", st, roi_align_bwd_kernel<<<K, kRoiThreads, smem, st>>>(a, tma_mask));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
