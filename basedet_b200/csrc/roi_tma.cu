// a13/a14 on the Tensor Memory Accelerator: ROIAlign 7x7 (2x2 samples) forward and backward for feature levels whose
// rows are 16-byte multiples (W % 4 == 0); other levels / pool shapes take the direct-load kernels of roi_align.cu.
// Reference: basedet/layers/common/roi_pool.py:35-78 -> F.nn.roi_align(mode="average", sample_points=2, aligned=True)
// (MegDNN semantics restated in oracle ASSUMED-6: taps outside the map read 0, lerp as a + (b - a) * t, mean of 4).
//
// forward : one CTA per ROI (7 compute warps + 1 DMA warp, 4 CTAs per SM).  The ROI's footprint (rows / columns
//           floor(first sample) .. floor(last sample) + 1) is fetched per channel chunk as [rows x 8 channels x BW columns]
//           boxes with cp.async.bulk.tensor (BW = 12, 20, ..., 60 floats; boxes of 8 / 4 / 2 rows; one tensor map per
//           (level, BW, rows)) into a ring of stages behind full / empty mbarriers.  Out-of-bounds box elements arrive as
//           zeros -- exactly the reference's border rule, so the inner loop has no bounds tests.  A thread owns (channel,
//           bin column): it walks the bin rows, keeps the horizontal lerps of the current two pixel rows of both sample
//           columns in registers and combines the 2x2 samples of a bin in the reference's order.  Outputs are staged in
//           shared memory and leave as one bulk copy per chunk (49 * channels floats are contiguous in (K, C, 7, 7)).
//           What bounds it (measured, profiles/r02_notes.md): NOT HBM -- the time is the same with the whole pyramid
//           resident in L2 -- but the latency of each CTA's load -> lerp -> store chain; hence 4 small CTAs per SM rather
//           than 2 large ones (-20% on small ROIs), and the op-by-op variants that cut instructions but also ILP lost.
// backward: same footprint boxes in the other direction: per chunk the footprint gradient tile is accumulated in
//           shared memory (no atomics inside the CTA: one thread owns a pixel) and flushed with
//           cp.reduce.async.bulk.tensor (element-wise fp32 add in L2, out-of-bounds elements dropped).
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "roi_common.cuh"

namespace bdet {

// Forward: boxes of 8, 4 and 2 rows per (level, width), so that a footprint of fh rows is fetched as fh rounded up to even
// rows (8 + 2 for 10 rows, not 16): less shared memory per channel, so more channels per stage and fewer stages per ROI.
// 10.5 KB of kernel parameters (CUDA >= 12.1: up to 32 KB).
constexpr int kHClasses = 3;
__host__ __device__ constexpr int box_height(int hc) { return kBoxH >> hc; }
struct alignas(64) RoiFwdMaps {
  CUtensorMap m[kTmaLevels][kWClasses][kHClasses];
};

struct RoiTmaArgs {
  RoiArgs r;
  unsigned level_mask;  // levels that have tensor maps
  int chunk_bytes;      // forward: target size of a footprint stage
};

// Sample-row program of one bin row (CTA-uniform, shared memory): for each of its two sample rows the footprint row r0 of
// the upper tap (<< 2; rows are `pitch` bytes apart) with the action in the low two bits
// (0 = both pixel rows are new, 1 = one row further: the lower lerp becomes the upper one, 2 = same rows as the previous
// sample) and the vertical fraction.
// (An op-list formulation -- ADVANCE row / EMIT sample, visiting every pixel row once -- executes 40% fewer instructions
// and measured 25% SLOWER: one op at a time leaves a warp a single dependent LDS -> FADD -> FMUL -> FADD chain, while this
// loop keeps the four lerps of a bin row's two sample rows in flight.)
struct __align__(16) RowProg {
  int o0;
  float f0;
  int o1;
  float f1;
};

__device__ __forceinline__ float lerp_at(const char* q, float t) {
  const float l = *reinterpret_cast<const float*>(q), r = *reinterpret_cast<const float*>(q + 4);
  return l + (r - l) * t;
}

// One forward task: channel c of the stage, bin column pw, bin rows [ph0, ph1).  ta / tb point at the channel's first
// footprint row, columns x0(2 pw) / x0(2 pw + 1).  The thread keeps the horizontal lerps of the current two pixel rows of
// both sample columns in registers, so a tap is read once per pixel row; the 2 x 2 samples of a bin are combined in the
// reference's order: ((v(0,0) + v(0,1)) + v(1,0)) + v(1,1), accumulated from 0.
__device__ __forceinline__ void fwd_task(const char* __restrict__ ta, const char* __restrict__ tb, float lxa, float lxb, int pitch,
                                         const RowProg* __restrict__ prog, int ph0, int ph1, float* __restrict__ oc) {
  float ua = 0.f, la = 0.f, ub = 0.f, lb = 0.f;  // upper / lower pixel row, sample columns a / b
  for (int ph = ph0; ph < ph1; ++ph) {
    const int4 pr = *reinterpret_cast<const int4*>(prog + ph);
    float v[2][2];
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
      const int o = iy ? pr.z : pr.x;
      const float fy = __int_as_float(iy ? pr.w : pr.y);
      const int mode = (iy == 0 && ph == ph0) ? 0 : (o & 3);
      if (mode != 2) {
        const int off = (o >> 2) * pitch;
        if (mode == 1) {
          ua = la;
          ub = lb;
        } else {
          ua = lerp_at(ta + off, lxa);
          ub = lerp_at(tb + off, lxb);
        }
        la = lerp_at(ta + off + pitch, lxa);
        lb = lerp_at(tb + off + pitch, lxb);
      }
      v[iy][0] = ua + (la - ua) * fy;
      v[iy][1] = ub + (lb - ub) * fy;
    }
    float acc = 0.f + v[0][0];
    acc = acc + v[0][1];
    acc = acc + v[1][0];
    acc = acc + v[1][1];
    oc[ph * 7] = acc * 0.25f;  // == acc / 4 exactly
  }
}

// Forward CTA = 7 compute warps + 1 DMA warp (lane 0 issues every bulk operation, so its per-thread bulk groups cover them
// all).  224 compute lanes = 32 channels x 7 bin columns: a pass over a 32-channel stage leaves no lane idle; stages of
// 16 / 8 channels are split into 2 / 4 runs of bin rows.
constexpr int kFwdLanes = 224;
// Backward (opt-in) CTA = 8 compute warps + 1 DMA warp.
constexpr int kComputeThreads = 256;
constexpr int kTmaThreads = kComputeThreads + 32;
// named barrier ids (0 = __syncthreads)
constexpr int kBarReady0 = 1, kBarFree0 = 3;  // + stage

constexpr int kFwdThreads = kFwdLanes + 32;
constexpr int kRingBytes = 36 * 1024;              // ring of footprint stages
constexpr int kFwdChunkBytes = 18 * 1024;          // stage size at most when the footprint allows: >= 2 stages in the ring
constexpr int kMaxC = 32;                          // channels per stage at most (= one pass of the 224 lanes)
constexpr int kOutStageBytes = kMaxC * 49 * 4;     // two out stages
constexpr int kFwdSmem = kRingBytes + 2 * kOutStageBytes;
__global__ void __launch_bounds__(kFwdThreads, 4)
roi_align_fwd_tma_kernel(const RoiTmaArgs a, const __grid_constant__ RoiFwdMaps maps) {
  // dynamic shared memory: [ring of footprint stages][out stage 0][out stage 1]; 1024-byte aligned by declaration
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ SampleTab14 ty, tx;
  __shared__ __align__(8) uint64_t full_bar[kMaxSlots], empty_bar[kMaxSlots];
  __shared__ FwdPlan plan;
  __shared__ RowProg prog[8];
  const RoiArgs& p = a.r;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int k = roi_of_cta(p, blockIdx.x);
  const int Cn = p.C;
  const RoiGeom g = roi_geom(p, k);
  float* out = p.out + (long long)k * p.C * 49;
  if (!g.valid) {
    for (int o = t; o < Cn * 49; o += kFwdThreads) out[o] = 0.f;
    return;
  }
  fill_axis(ty, 7, 2, g.start_h, g.bin_h, kFwdThreads);
  fill_axis(tx, 7, 2, g.start_w, g.bin_w, kFwdThreads);
  if (t == 32) {  // (a slot is consumed by one group: kFwdLanes / 32 arrivals free it)
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kFwdLanes / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (t == 0) {
    FwdPlan pl;
    pl.xs = tx.i0[0] & ~3;  // TMA: the innermost box coordinate must be a multiple of 16 bytes (measured: faults otherwise)
    pl.ys = ty.i0[0];
    const long long fw = (long long)tx.i0[13] + 2 - pl.xs, fh = (long long)ty.i0[13] + 2 - pl.ys;
    // finite, ordered coordinates only (NaN / inf rois take the direct path, which tests every tap)
    const bool sane = fabsf(g.start_w) < 1e8f && fabsf(g.start_h) < 1e8f && g.bin_w < 1e7f && g.bin_h < 1e7f && fw >= 2 && fh >= 2;
    pl.cls = -1;
    pl.nrb = 0;
    pl.ccs = 0;
    if (sane && ((a.level_mask >> g.lvl) & 1u) && fw <= kMaxBoxWidth && fh <= kFwdMaxRows) {
      pl.cls = width_class(fw);
      pl.nrb = (int)((fh + 1) & ~1ll);  // (forward: ROWS fetched, boxes of 8 / 4 / 2)
      const int per_c = pl.nrb * box_width(pl.cls) * 4;  // bytes per channel
      // stages of <= chunk_bytes keep two or more of them in the ring; footprints too large for that take the whole ring
      // per stage (load and compute alternate; the SM's other three CTAs fill the gaps)
      int fit = a.chunk_bytes / per_c;
      if (fit < 8) fit = kRingBytes / per_c;  // (a single stage may fill the ring: load and compute alternate)
      int ccs = fit >= 32 ? 32 : (fit >= 16 ? 16 : (fit >= 8 ? 8 : 0));
      while (ccs > 8 && ccs > Cn) ccs >>= 1;  // 8 <= Cn (multiple of 8); a tail chunk may still be partial
      pl.ccs = ccs;
      if (ccs < kBoxC) pl.cls = -1;  // footprint too large for one stage
    }
    plan = pl;
  }
  if (t >= 32 && t < 39) {
    // row program (offsets in ROWS here; scaled by the pitch once the width class is known)
    const int b = t - 32, ys = ty.i0[0];
    RowProg rp;
    const int ra = ty.i0[2 * b] - ys, rb = ty.i0[2 * b + 1] - ys;
    const int da = b > 0 ? ra - (ty.i0[2 * b - 1] - ys) : 2, db = rb - ra;
    rp.o0 = ra << 2 | (da == 0 ? 2 : (da == 1 ? 1 : 0));
    rp.o1 = rb << 2 | (db == 0 ? 2 : (db == 1 ? 1 : 0));
    rp.f0 = ty.frac[2 * b];
    rp.f1 = ty.frac[2 * b + 1];
    prog[b] = rp;
  }
  __syncthreads();
  if (plan.cls < 0) {
    roi_fwd_direct<7, 7, 2>(p, k, g, ty, tx, kFwdThreads);
    return;
  }
  // ---- TMA path
  unsigned char* const stage0 = smem_raw;
  float* const ostage0 = reinterpret_cast<float*>(smem_raw + kRingBytes);
  const int BW = box_width(plan.cls), rows = plan.nrb, CCS = plan.ccs, xs = plan.xs, ys = plan.ys;
  const int ncb = CCS / kBoxC;                    // channel boxes per stage
  // a stage is [channel box][row][8 channels][BW]: footprint rows of a channel are 8 * BW floats apart (a multiple of
  // 128 bytes: every row is a legal TMA destination), whatever boxes brought them
  const int pitch = kBoxC * BW * 4, cb_bytes = rows * pitch;
  const int n_chunks = (Cn + CCS - 1) / CCS;
  const int z0 = g.n * p.C;
  // ring of footprint stages: slot = one chunk (rounded up to 1 KB), as many slots as fit (2 .. kMaxSlots)
  const int slot_bytes = min((ncb * cb_bytes + 1023) & ~1023, kRingBytes);
  const int nslots = min(kMaxSlots, kRingBytes / slot_bytes);
  if (warp == kFwdLanes / 32) {
    // ---------------- DMA warp: footprint boxes in, finished output chunks out
    const CUtensorMap* map = &maps.m[g.lvl][plan.cls][0];  // + height class
    int ld_slot = 0, ld_round = 0;  // slot / round of the next chunk to load (chunks are loaded in order)
    auto load = [&](int chunk) {
      const int s = ld_slot, c0 = chunk * CCS;
      if (ld_round >= 1) mbar_wait_sleep(&empty_bar[s], (uint32_t)((ld_round - 1) & 1));  // the compute warps are done with the slot
      if (lane == 0) {
        const int cbs = min(ncb, (Cn - c0 + kBoxC - 1) / kBoxC);
        mbar_expect_tx(&full_bar[s], (uint32_t)(cbs * cb_bytes));
        for (int cb = 0; cb < cbs; ++cb) {
          unsigned char* dst = stage0 + s * slot_bytes + cb * cb_bytes;
          const int z = z0 + c0 + cb * kBoxC;
          int r = 0;
          for (; r + 8 <= rows; r += 8) tma_load_3d(dst + r * pitch, map, xs, z, ys + r, &full_bar[s]);
          if (rows - r >= 4) {
            tma_load_3d(dst + r * pitch, map + 1, xs, z, ys + r, &full_bar[s]);
            r += 4;
          }
          if (rows - r >= 2) tma_load_3d(dst + r * pitch, map + 2, xs, z, ys + r, &full_bar[s]);
        }
      }
      if (++ld_slot == nslots) {
        ld_slot = 0;
        ++ld_round;
      }
    };
    for (int c = 0; c < nslots - 1 && c < n_chunks; ++c) load(c);
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const int s = chunk & 1, c0 = chunk * CCS;
      if (chunk + nslots - 1 < n_chunks) load(chunk + nslots - 1);
      bar_sync(kBarReady0 + s, kFwdLanes + 32);  // the out stage holds chunk `chunk`
      if (lane == 0) {
        bulk_store(out + (size_t)c0 * 49, ostage0 + s * (kOutStageBytes / 4), (uint32_t)min(CCS, Cn - c0) * 49 * 4);
        bulk_commit();
        bulk_wait_read<0>();  // shared memory is read out: the stage may be rewritten (and must outlive the copy)
      }
      __syncwarp();
      if (chunk + 2 < n_chunks) bar_arrive(kBarFree0 + s, kFwdLanes + 32);
    }
    return;
  }

  // ---------------- compute warps: lane -> (channel, run of bin rows, bin column), the same for every chunk
  const int tg = t;
  int c_l, seg, pw, ph0, ph1;  // (divisions by constants)
  if (CCS >= 32) {
    c_l = tg / 7, pw = tg - c_l * 7, seg = 0, ph0 = 0, ph1 = 7;
  } else if (CCS >= 16) {
    c_l = tg / 14;
    const int rem = tg - c_l * 14;
    seg = rem / 7, pw = rem - seg * 7, ph0 = seg * 4, ph1 = min(7, ph0 + 4);
  } else {
    c_l = tg / 28;
    const int rem = tg - c_l * 28;
    seg = rem / 7, pw = rem - seg * 7, ph0 = seg * 2, ph1 = min(7, ph0 + 2);
  }
  const float lxa = tx.frac[2 * pw], lxb = tx.frac[2 * pw + 1];
  const int xa = (tx.i0[2 * pw] - xs) * 4, xb = (tx.i0[2 * pw + 1] - xs) * 4;

  int slot = 0, round = 0;
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int s = chunk & 1, c0 = chunk * CCS;
    const int nc = min(CCS, Cn - c0);
    if (chunk >= 2) bar_sync(kBarFree0 + s, kFwdLanes + 32);  // the bulk store of chunk - 2 has read the out stage
    mbar_wait_sleep(&full_bar[slot], (uint32_t)(round & 1));
    const char* tile = reinterpret_cast<const char*>(stage0 + slot * slot_bytes);
    float* os = ostage0 + s * (kOutStageBytes / 4);
    if (c_l < nc) {
      const char* tc = tile + (c_l >> 3) * cb_bytes + (c_l & 7) * (BW * 4);
      fwd_task(tc + xa, tc + xb, lxa, lxb, pitch, prog, ph0, ph1, os + c_l * 49 + pw);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[slot]);  // this warp no longer reads the stage
    fence_proxy_async_smem();                   // its out-stage writes are visible to the bulk copy
    bar_arrive(kBarReady0 + s, kFwdLanes + 32);
    if (++slot == nslots) {
      slot = 0;
      ++round;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward.  Whether a ROI is taken by the TMA kernel is a pure function of its geometry and the level mask
// (roi_bwd_takes_tma, roi_common.cuh), so the direct scatter kernel of roi_align.cu (launched right after with the same
// mask) skips exactly those ROIs.
//
// Per channel chunk the gradient tile of the footprint is built in shared memory and added to dfeat with one
// cp.reduce.async.bulk.tensor per [8 rows][8 channels][BW] box.  A lane owns one footprint column of one channel
// (8 / 16 / 32 lanes per channel, so small footprints pack 4 / 2 channels into a warp): T[ph] = sum_pw Wx[x][pw] / 4 *
// dout[ph][pw] in registers, then per box row the 7-term dot product with Wy[row][:] (the separable weights of the
// direct kernel) and one store.  No atomics inside the CTA.  The DMA warp brings the dout chunks in (1-D bulk copies)
// and sends the finished tiles out.
constexpr int kBwdTmaDefaultCls = -1;  // measured (profiles/): the direct scatter kernel is faster on every ROI mix tried
constexpr int kBwdRawBytes = kBwdMaxC * 49 * 4;    // dout chunk as it lies in memory
constexpr int kBwdWxBytes = 8 * kWClasses * 8 * 4;  // Wx[x][pw] / 4
constexpr int kBwdWyBytes = kBwdMaxRows * 8 * 4;   // Wy[row][ph]
constexpr int kBwdSmem = 2 * kBwdStageBytes + 2 * kBwdRawBytes + kBwdWxBytes + kBwdWyBytes;
// the backward's tensor maps: box widths 8, 16, ..., 56 floats x box heights 8 / 4 / 2 rows
struct alignas(64) RoiBwdMaps {
  CUtensorMap m[kTmaLevels][kWClasses][kHClasses];
};

__global__ void __launch_bounds__(kTmaThreads, 3)
roi_align_bwd_tma_kernel(const RoiTmaArgs a, const __grid_constant__ RoiBwdMaps maps) {
  // dynamic shared memory: [tile stage 0][tile stage 1][raw 0][raw 1][wx][wy]; 1024-byte aligned by declaration
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ SampleTab14 ty, tx;
  __shared__ __align__(8) uint64_t full_bar[kMaxSlots], empty_bar[kMaxSlots];
  __shared__ float trash[32];  // lanes beyond the box width store here (keeps the stores branch-free)
  const RoiArgs& p = a.r;
  const int k = roi_of_cta(p, blockIdx.x), t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const RoiGeom g = roi_geom(p, k);
  if (!g.valid) return;
  FwdPlan plan;
  if (roi_bwd_takes_tma(p, g, a.level_mask, &plan) != 1) return;  // the direct kernel takes this ROI, or nothing to do
  float* const stage0 = reinterpret_cast<float*>(smem_raw);
  float* const raw0 = reinterpret_cast<float*>(smem_raw + 2 * kBwdStageBytes);
  float* const wx = reinterpret_cast<float*>(smem_raw + 2 * kBwdStageBytes + 2 * kBwdRawBytes);
  float* const wy = reinterpret_cast<float*>(smem_raw + 2 * kBwdStageBytes + 2 * kBwdRawBytes + kBwdWxBytes);
  fill_axis(ty, 7, 2, g.start_h, g.bin_h, kTmaThreads);
  fill_axis(tx, 7, 2, g.start_w, g.bin_w, kTmaThreads);
  const int BW = 8 * (plan.cls + 1), rows = plan.nrb, CCS = plan.ccs, xs = plan.xs, ys = plan.ys;
  // a tile stage is [channel box][row][8 channels][BW]: rows are 8 * BW floats apart (a multiple of 128 bytes, so any row
  // can start a box of 8 / 4 / 2 rows)
  const int pitch = kBoxC * BW, cb_floats = rows * pitch;
  const int n_chunks = (p.C + CCS - 1) / CCS;
  const int z0 = g.n * p.C;
  const float* dout = p.dout + (size_t)k * p.C * 49;
  // ring of dout chunks (CCS * 196 bytes each): as many slots as fit, so that the loads run several chunks ahead
  const int raw_floats = CCS * 49;
  const int nslots = min(min(kMaxSlots, n_chunks), (2 * kBwdRawBytes) / (raw_floats * 4));
  if (t == 0) {
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kComputeThreads / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();  // tables, barriers
  // Wx[x][pw] / 4: summed tap weights of bin pw's two sample columns on footprint column x (zero beyond the footprint);
  // the factor 1 / (samples per bin) is folded in here (a power of two: the products are unchanged)
  for (int i = t; i < BW * 8; i += kTmaThreads) {
    const int x = i >> 3, pw = i & 7;
    float w = 0.f;
    if (pw < 7) {
#pragma unroll
      for (int ix = 0; ix < 2; ++ix) {
        const int sxi = 2 * pw + ix, rel = tx.i0[sxi] - xs;
        const float fr = tx.frac[sxi];
        w += rel == x ? 1.f - fr : (rel + 1 == x ? fr : 0.f);
      }
    }
    wx[i] = w * 0.25f;
  }
  // Wy[r][ph]: the same for the rows; every row of the tile is filled, so a row no sample touches gets zero weights and the
  // tile needs no separate zero fill (the reduce adds whole boxes)
  for (int i = t; i < rows * 8; i += kTmaThreads) {
    const int r = i >> 3, ph = i & 7;
    float w = 0.f;
    if (ph < 7) {
#pragma unroll
      for (int iy = 0; iy < 2; ++iy) {
        const int syi = 2 * ph + iy, rel = ty.i0[syi] - ys;
        const float fr = ty.frac[syi];
        w += rel == r ? 1.f - fr : (rel + 1 == r ? fr : 0.f);
      }
    }
    wy[i] = w;
  }
  __syncthreads();

  if (warp == kComputeThreads / 32) {
    // ---------------- DMA warp: dout chunks in, finished tiles out (reduce-add into dfeat)
    const CUtensorMap* map = &maps.m[g.lvl][plan.cls][0];  // + height class
    int ld_slot = 0, ld_round = 0;
    auto load = [&](int chunk) {
      const int s = ld_slot;
      if (ld_round >= 1) mbar_wait_sleep(&empty_bar[s], (uint32_t)((ld_round - 1) & 1));
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)min(CCS, p.C - chunk * CCS) * 196;
        mbar_expect_tx(&full_bar[s], bytes);
        bulk_load(raw0 + s * raw_floats, dout + (size_t)chunk * CCS * 49, bytes, &full_bar[s]);
      }
      if (++ld_slot == nslots) {
        ld_slot = 0;
        ++ld_round;
      }
    };
    for (int c = 0; c < nslots - 1; ++c) load(c);
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const int s = chunk & 1, c0 = chunk * CCS;
      if (chunk + nslots - 1 < n_chunks) load(chunk + nslots - 1);
      bar_sync(kBarReady0 + s, kTmaThreads);  // tile stage s holds chunk `chunk`
      if (lane == 0) {
        const int cbs = (min(CCS, p.C - c0) + kBoxC - 1) / kBoxC;
        const float* tile = stage0 + s * (kBwdStageBytes / 4);
        for (int cb = 0; cb < cbs; ++cb) {
          const float* src = tile + cb * cb_floats;
          const int z = z0 + c0 + cb * kBoxC;
          int r = 0;
          for (; r + 8 <= rows; r += 8) tma_reduce_add_3d(map, xs, z, ys + r, src + r * pitch);
          if (rows - r >= 4) {
            tma_reduce_add_3d(map + 1, xs, z, ys + r, src + r * pitch);
            r += 4;
          }
          if (rows - r >= 2) tma_reduce_add_3d(map + 2, xs, z, ys + r, src + r * pitch);
        }
        bulk_commit();
        bulk_wait_read<0>();  // the tile is read out: the stage may be rewritten (and must outlive the reduce)
      }
      __syncwarp();
      if (chunk + 2 < n_chunks) bar_arrive(kBarFree0 + s, kTmaThreads);
    }
    return;
  }

  // ---------------- compute warps
  // lanes per channel: 8 / 16 / 32 (two column passes when BW > 32)
  const int lpc = BW <= 8 ? 8 : (BW <= 16 ? 16 : 32);
  const int cpw = 32 / lpc;                       // channels per warp pass
  const int xl = lane & (lpc - 1), csub = lane / lpc;
  const int xpasses = (BW + 31) / 32;
  int slot = 0, round = 0;

  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int s = chunk & 1, c0 = chunk * CCS;
    const int nc = min(CCS, p.C - c0);
    if (chunk >= 2) bar_sync(kBarFree0 + s, kTmaThreads);  // the reduce of chunk - 2 has read tile stage s
    mbar_wait_sleep(&full_bar[slot], (uint32_t)(round & 1));
    float* tile = stage0 + s * (kBwdStageBytes / 4);
    const float* raw = raw0 + slot * raw_floats;
    for (int cb = warp * cpw; cb < nc; cb += (kComputeThreads / 32) * cpw) {
      const int c = cb + csub;
      const bool cok = c < nc;
      const float* dc = raw + (cok ? c : nc - 1) * 49;
      for (int xp = 0; xp < xpasses; ++xp) {
        const int x = xp * 32 + xl;
        const bool ok = cok && x < BW;
        float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
        if (x < BW) {
          wa = *reinterpret_cast<const float4*>(wx + x * 8);
          wb = *reinterpret_cast<const float4*>(wx + x * 8 + 4);
        }
        float T[7];
#pragma unroll
        for (int ph = 0; ph < 7; ++ph) {
          const float* d = dc + ph * 7;
          float v = wa.x * d[0];
          v = __fmaf_rn(wa.y, d[1], v);
          v = __fmaf_rn(wa.z, d[2], v);
          v = __fmaf_rn(wa.w, d[3], v);
          v = __fmaf_rn(wb.x, d[4], v);
          v = __fmaf_rn(wb.y, d[5], v);
          v = __fmaf_rn(wb.z, d[6], v);
          T[ph] = v;
        }
        // 32-bit shared addresses; a lane beyond the box width is redirected to its trash slot (branch-free stores)
        uint32_t ad = ok ? smem_u32(tile + (c >> 3) * cb_floats + (c & 7) * BW + x) : smem_u32(trash + lane);
        const uint32_t row_pitch = ok ? (uint32_t)pitch * 4u : 0u;
#pragma unroll 2
        for (int r = 0; r < rows; ++r, ad += row_pitch) {
          const float4 u = *reinterpret_cast<const float4*>(wy + r * 8);
          const float4 w2 = *reinterpret_cast<const float4*>(wy + r * 8 + 4);
          float sum = u.x * T[0];
          sum = __fmaf_rn(u.y, T[1], sum);
          sum = __fmaf_rn(u.z, T[2], sum);
          sum = __fmaf_rn(u.w, T[3], sum);
          sum = __fmaf_rn(w2.x, T[4], sum);
          sum = __fmaf_rn(w2.y, T[5], sum);
          sum = __fmaf_rn(w2.z, T[6], sum);
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(ad), "f"(sum) : "memory");
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[slot]);  // this warp no longer reads its dout slot
    fence_proxy_async_smem();                   // its tile writes are visible to the reduce
    bar_arrive(kBarReady0 + s, kTmaThreads);
    if (++slot == nslots) {
      slot = 0;
      ++round;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Host: tensor maps.  cuTensorMapEncodeTiled is fetched through the runtime (no link-time libcuda dependency); maps are
// pure functions of (pointer, H, W, B * C, box width) and are cached per thread.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  int H, W;
  long long BC;
  int bwd;  // width family: forward 12, 20, ..., backward 8, 16, ...
  bool operator==(const MapKey& o) const { return ptr == o.ptr && H == o.H && W == o.W && BC == o.BC && bwd == o.bwd; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (size_t)k.H * 1000003u ^ (size_t)k.W * 10007u ^ (size_t)k.BC * 31u;
  }
};
struct LevelMaps {
  CUtensorMap m[kWClasses][kHClasses];
};

// Builds (or finds) the kWClasses maps of one level.  Returns false when the level cannot be described (alignment).
static bool level_maps(const void* ptr, int H, int W, long long BC, bool bwd, const LevelMaps** out) {
  static thread_local std::unordered_map<MapKey, LevelMaps, MapKeyHash>* cache = nullptr;
  if (!cache) cache = new std::unordered_map<MapKey, LevelMaps, MapKeyHash>();
  if ((W & 3) || (reinterpret_cast<uintptr_t>(ptr) & 15u) || BC < 1 || BC > 0x7fffffffll) return false;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  MapKey key{ptr, H, W, BC, bwd ? 1 : 0};
  auto it = cache->find(key);
  if (it == cache->end()) {
    if (cache->size() > 256) cache->clear();
    LevelMaps lm;
    for (int c = 0; c < kWClasses; ++c)
      for (int h = 0; h < kHClasses; ++h) {
        // dimension order (x, channel, y): a box lands in shared memory as [rows][8 channels][BW], so the lanes of a
        // warp, which read one pixel row of neighbouring channels, spread over the banks (channel stride = BW floats)
        cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)BC, (cuuint64_t)H};
        cuuint64_t strides[2] = {(cuuint64_t)W * H * 4, (cuuint64_t)W * 4};
        cuuint32_t box[3] = {(cuuint32_t)(bwd ? 8 * (c + 1) : box_width(c)), (cuuint32_t)kBoxC, (cuuint32_t)box_height(h)};
        cuuint32_t es[3] = {1, 1, 1};
        alignas(64) CUtensorMap m;
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
        lm.m[c][h] = m;
      }
    it = cache->emplace(key, lm).first;
  }
  *out = &it->second;
  return true;
}

static int tma_mode() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("BDET_ROI_TMA");
    on = e ? atoi(e) : 1;
  }
  return on;
}
static bool tma_enabled() { return tma_mode() != 0; }
// widest footprint class (box width 8 * (cls + 1)) the TMA backward takes; wider footprints go to the direct scatter
// kernel.  BDET_ROI_BWD_TMA_CLS overrides (-1 = TMA backward off, 6 = every width).
static int bwd_max_cls() {
  static int v = -100;
  if (v == -100) {
    const char* e = getenv("BDET_ROI_BWD_TMA_CLS");
    v = e ? atoi(e) : kBwdTmaDefaultCls;
    if (v >= kWClasses) v = kWClasses - 1;
  }
  return v;
}
// Forward launch through the TMA kernel when at least one level qualifies.  Returns 1 if launched, 0 if the caller
// should use the direct kernel, < 0 on error.
int roi_fwd_tma_launch(const RoiArgs& a, cudaStream_t st) {
  if (!tma_enabled() || a.PH != 7 || a.PW != 7 || a.SH != 2 || a.SW != 2) return 0;
  if (a.lv.n_levels > kTmaLevels || (a.C & 7) || a.C < 8) return 0;
  if ((reinterpret_cast<uintptr_t>(a.out) & 15u)) return 0;
  static thread_local RoiFwdMaps* maps = nullptr;  // filled per call, passed by value
  if (!maps) maps = new RoiFwdMaps();
  RoiTmaArgs ta;
  ta.r = a;
  ta.level_mask = 0;
  for (int l = 0; l < a.lv.n_levels; ++l) {
    const LevelMaps* lm = nullptr;
    if (level_maps(a.lv.feat[l], a.lv.H[l], a.lv.W[l], (long long)a.B * a.C, false, &lm)) {
      for (int c = 0; c < kWClasses; ++c)
        for (int h = 0; h < kHClasses; ++h) maps->m[l][c][h] = lm->m[c][h];
      ta.level_mask |= 1u << l;
    }
  }
  if (!ta.level_mask) return 0;
  ta.chunk_bytes = kFwdChunkBytes;
  if (cudaFuncSetAttribute(roi_align_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem) != cudaSuccess)
    return set_error(BDET_ECUDA, "roi_align_fwd: cannot reserve %d bytes of shared memory", kFwdSmem);
  BDET_KERNEL("roi_align_fwd_tma_kernel", st, roi_align_fwd_tma_kernel<<<a.K, kFwdThreads, kFwdSmem, st>>>(ta, *maps));
  return 1;
}

// Backward launch: the TMA kernel takes the ROIs `roi_bwd_takes_tma` accepts; *level_mask_out tells the direct kernel
// which levels have tensor maps (0 = nothing was launched and the direct kernel takes every ROI).
int roi_bwd_tma_launch(const RoiArgs& a, cudaStream_t st, unsigned* level_mask_out) {
  *level_mask_out = 0;
  if (!tma_enabled() || a.PH != 7 || a.PW != 7 || a.SH != 2 || a.SW != 2) return 0;
  if (a.lv.n_levels > kTmaLevels || (a.C & 7) || a.C < 8) return 0;
  if ((reinterpret_cast<uintptr_t>(a.dout) & 15u)) return 0;
  static thread_local RoiBwdMaps* maps = nullptr;
  if (!maps) maps = new RoiBwdMaps();
  RoiTmaArgs ta;
  ta.r = a;
  ta.level_mask = 0;
  for (int l = 0; l < a.lv.n_levels; ++l) {
    const LevelMaps* lm = nullptr;
    if (level_maps(a.lv.dfeat[l], a.lv.H[l], a.lv.W[l], (long long)a.B * a.C, true, &lm)) {
      for (int c = 0; c < kWClasses; ++c)
        for (int h = 0; h < kHClasses; ++h) maps->m[l][c][h] = lm->m[c][h];
      ta.level_mask |= 1u << l;
    }
  }
  if (!ta.level_mask) return 0;
  const int max_cls = bwd_max_cls();
  if (max_cls < 0) return 0;
  ta.level_mask |= (unsigned)max_cls << 16;  // travels with the mask to both kernels (bwd_plan)
  if (cudaFuncSetAttribute(roi_align_bwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem) != cudaSuccess)
    return set_error(BDET_ECUDA, "roi_align_bwd: cannot reserve %d bytes of shared memory", kBwdSmem);
  BDET_KERNEL("roi_align_bwd_tma_kernel", st, roi_align_bwd_tma_kernel<<<a.K, kTmaThreads, kBwdSmem, st>>>(ta, *maps));
  *level_mask_out = ta.level_mask;
  return 1;
}

}  // namespace bdet
