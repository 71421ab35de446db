// a13/a14 on the Tensor Memory Accelerator: ROIAlign 7x7 (2x2 samples) forward and backward for feature levels whose
// rows are 16-byte multiples (W % 4 == 0); other levels / pool shapes take the direct-load kernels of roi_align.cu.
// Reference: basedet/layers/common/roi_pool.py:35-78 -> F.nn.roi_align(mode="average", sample_points=2, aligned=True)
// (MegDNN semantics restated in oracle ASSUMED-6: taps outside the map read 0, lerp as a + (b - a) * t, mean of 4).
//
// forward : one CTA per ROI.  The ROI's footprint (rows / columns floor(first sample) .. floor(last sample) + 1) is
//           fetched per channel chunk as [8 channels x 8 rows x BW columns] boxes with cp.async.bulk.tensor (BW = the
//           footprint width rounded up to 8; one tensor map per (level, BW)), double-buffered behind an mbarrier while
//           the previous chunk is computed.  Out-of-bounds box elements arrive as zeros -- exactly the reference's
//           border rule, so the inner loop has no bounds tests.  A thread owns (channel, sample column): it walks the 14
//           sample rows, keeps the two horizontal lerps of the current pixel rows in registers and never re-reads a tap
//           (2 shared loads per footprint row); the 2x2 samples of a bin are combined in the reference's order with one
//           shuffle.  Outputs are staged in shared memory and leave as one bulk copy per chunk (49 * channels floats are
//           contiguous in (K, C, 7, 7)).
// backward: same footprint boxes in the other direction: per chunk the footprint gradient tile is accumulated in
//           shared memory (no atomics inside the CTA: one thread owns a pixel) and flushed with
//           cp.reduce.async.bulk.tensor (element-wise fp32 add in L2, out-of-bounds elements dropped).
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "roi_common.cuh"

namespace bdet {

struct alignas(64) RoiTmaMaps {
  CUtensorMap m[kTmaLevels][kWClasses];
};

struct RoiTmaArgs {
  RoiArgs r;
  unsigned level_mask;  // levels that have tensor maps
  const RoiTmaMaps* gmaps;  // debug (BDET_ROI_TMA=2): descriptors read from global memory instead of the parameters
};

template <int kThreads>
__global__ void __launch_bounds__(kThreads, 2)
roi_align_fwd_tma_kernel(const RoiTmaArgs a, const __grid_constant__ RoiTmaMaps maps) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ SampleTab ty, tx;
  __shared__ __align__(8) uint64_t full_bar[2];
  __shared__ FwdPlan plan;
  const RoiArgs& p = a.r;
  const int k = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const RoiGeom g = roi_geom(p, k);
  float* out = p.out + (long long)k * p.C * 49;
  if (!g.valid) {
    for (int o = t; o < p.C * 49; o += kThreads) out[o] = 0.f;
    return;
  }
  fill_axis(ty, 7, 2, g.start_h, g.bin_h, kThreads);
  fill_axis(tx, 7, 2, g.start_w, g.bin_w, kThreads);
  if (t == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
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
    if (sane && ((a.level_mask >> g.lvl) & 1u) && fw <= 8 * kWClasses) {
      pl.cls = (int)((fw + 7) / 8) - 1;
      pl.nrb = (int)((fh + kBoxH - 1) / kBoxH);
      const long long per_c = (long long)pl.nrb * kBoxH * 8 * (pl.cls + 1) * 4;  // bytes per channel
      long long ccs = (kStageBytes / per_c) & ~7ll;
      if (ccs > kMaxCCS) ccs = kMaxCCS;
      if (ccs > p.C) ccs = p.C;
      pl.ccs = (int)ccs;
      if (ccs < kBoxC) pl.cls = -1;  // footprint too tall for one stage
    }
    plan = pl;
  }
  __syncthreads();
  if (plan.cls < 0) {
    roi_fwd_direct<7, 7, 2>(p, k, g, ty, tx, kThreads);
    return;
  }
  // ---- TMA path
  const uintptr_t base = (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023;
  auto stage = [&](int s) { return reinterpret_cast<float*>(base + (uintptr_t)s * kStageBytes); };
  auto ostage = [&](int s) { return reinterpret_cast<float*>(base + 2 * kStageBytes + (uintptr_t)s * kOutStageBytes); };
  const int BW = 8 * (plan.cls + 1), nrb = plan.nrb, CCS = plan.ccs, xs = plan.xs, ys = plan.ys;
  const int ncb = CCS / kBoxC;                    // channel boxes per stage
  const int box_floats = kBoxC * kBoxH * BW;      // one box: [8 ch][8 rows][BW]
  const int rb_stride = ncb * box_floats;         // floats between row boxes of a stage
  const int n_chunks = (p.C + CCS - 1) / CCS;
  const CUtensorMap* map = a.gmaps ? &a.gmaps->m[g.lvl][plan.cls] : &maps.m[g.lvl][plan.cls];
  const int z0 = g.n * p.C;

  // per-thread copies of the (CTA-uniform) sample-row program: relative pixel row and lerp fraction of the 14 sample rows
  int yr[14];
  float fy[14];
#pragma unroll
  for (int s = 0; s < 14; ++s) {
    yr[s] = ty.i0[s] - ys;
    fy[s] = ty.frac[s];
  }

  auto issue = [&](int chunk) {  // warp 0: one box per lane
    const int s = chunk & 1, c0 = chunk * CCS;
    const int cbs = min(ncb, (p.C - c0 + kBoxC - 1) / kBoxC);
    const int nbox = nrb * cbs;
    if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)nbox * box_floats * 4);
    __syncwarp();
    for (int b = lane; b < nbox; b += 32) {
      const int rb = b / cbs, cb = b - rb * cbs;
      tma_load_3d(stage(s) + rb * rb_stride + cb * box_floats, map, xs, ys + rb * kBoxH, z0 + c0 + cb * kBoxC,
                  &full_bar[s]);
    }
  };

  if (warp == 0) issue(0);
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int s = chunk & 1, c0 = chunk * CCS;
    const int nc = min(CCS, p.C - c0);
    if (t == 0) bulk_wait_read<1>();  // the bulk store that last read ostage[s] (chunk - 2) has drained
    __syncthreads();                  // everyone is done with stage[s ^ 1] (chunk - 1) and may write ostage[s]
    if (warp == 0 && chunk + 1 < n_chunks) issue(chunk + 1);
    mbar_wait(&full_bar[s], (uint32_t)((chunk >> 1) & 1));
    const float* tile = stage(s);
    float* os = ostage(s);
    const int total = nc * 14;
    for (int base_task = warp * 32; base_task < total; base_task += kThreads) {
      const int task = base_task + lane;
      const bool live = task < total;
      const int tk = live ? task : total - 1;
      const int c = tk / 14, sx = tk - c * 14;
      const float* tc = tile + (c >> 3) * box_floats + (c & 7) * (kBoxH * BW) + (tx.i0[sx] - xs);
      const float lx = tx.frac[sx];
      auto hlerp = [&](int r) -> float {
        const float* q = tc + (r >> 3) * rb_stride + (r & 7) * BW;
        const float l = q[0], rr = q[1];
        return l + (rr - l) * lx;
      };
      float v[14];
      int rcur = -0x40000000;
      float ha = 0.f, hb = 0.f;
#pragma unroll
      for (int sy = 0; sy < 14; ++sy) {
        const int r0 = yr[sy];  // CTA-uniform: the branches below do not diverge
        if (r0 != rcur) {
          if (r0 == rcur + 1) {
            ha = hb;
          } else {
            ha = hlerp(r0);
          }
          hb = hlerp(r0 + 1);
          rcur = r0;
        }
        v[sy] = ha + (hb - ha) * fy[sy];
      }
      // bin (ph, pw): ((v(0,0) + v(0,1)) + v(1,0)) + v(1,1), the reference's accumulation order from 0
      float o[7];
#pragma unroll
      for (int ph = 0; ph < 7; ++ph) {
        const float p0 = __shfl_down_sync(0xffffffffu, v[2 * ph], 1);
        const float p1 = __shfl_down_sync(0xffffffffu, v[2 * ph + 1], 1);
        float acc = 0.f + v[2 * ph];
        acc = acc + p0;
        acc = acc + v[2 * ph + 1];
        acc = acc + p1;
        o[ph] = acc * 0.25f;  // == acc / 4 exactly
      }
      if (live && !(sx & 1)) {
        float* oc = os + c * 49 + (sx >> 1);
#pragma unroll
        for (int ph = 0; ph < 7; ++ph) oc[ph * 7] = o[ph];
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (t == 0) {
      bulk_store(out + (size_t)c0 * 49, os, (uint32_t)nc * 49 * 4);
      bulk_commit();
    }
  }
  if (t == 0) bulk_wait_read<0>();  // shared memory must outlive the last bulk store's reads
}

// ---------------------------------------------------------------------------------------------------------------
// Backward.  Whether a ROI is taken by the TMA kernel is a pure function of its geometry and the level mask, so the
// direct scatter kernel of roi_align.cu (launched right after with the same mask) skips exactly those ROIs.
constexpr int kBwdRawBytes = kMaxCCS * 49 * 4;     // dout chunk as it lies in memory
constexpr int kBwdPadBytes = kMaxCCS * 56 * 4;     // dout chunk / 4, rows padded to 8 floats
constexpr int kBwdWxBytes = 8 * kWClasses * 8 * 4;  // Wx[x][pw]
constexpr int kBwdSmem = 2 * kBwdStageBytes + kBwdRawBytes + kBwdPadBytes + kBwdWxBytes + 1024;

template <int kThreads>
__global__ void __launch_bounds__(kThreads, 2)
roi_align_bwd_tma_kernel(const RoiTmaArgs a, const __grid_constant__ RoiTmaMaps maps) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ SampleTab ty, tx;
  __shared__ __align__(8) uint64_t raw_bar;
  const RoiArgs& p = a.r;
  const int k = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const RoiGeom g = roi_geom(p, k);
  if (!g.valid) return;
  FwdPlan plan;
  if (roi_bwd_takes_tma(p, g, a.level_mask, &plan) != 1) return;  // the direct kernel takes this ROI, or nothing to do
  const uintptr_t base = (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023;
  auto stage = [&](int s) { return reinterpret_cast<float*>(base + (uintptr_t)s * kBwdStageBytes); };
  float* raw = reinterpret_cast<float*>(base + 2 * kBwdStageBytes);
  float* pad = reinterpret_cast<float*>(base + 2 * kBwdStageBytes + kBwdRawBytes);
  float* wx = reinterpret_cast<float*>(base + 2 * kBwdStageBytes + kBwdRawBytes + kBwdPadBytes);
  fill_axis(ty, 7, 2, g.start_h, g.bin_h, kThreads);
  fill_axis(tx, 7, 2, g.start_w, g.bin_w, kThreads);
  if (t == 0) {
    mbar_init(&raw_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int BW = 8 * (plan.cls + 1), nrb = plan.nrb, CCS = plan.ccs, xs = plan.xs, ys = plan.ys;
  const int ncb = CCS / kBoxC;
  const int box_floats = kBoxC * kBoxH * BW;
  const int rb_stride = ncb * box_floats;
  const int n_chunks = (p.C + CCS - 1) / CCS;
  const int rows = nrb * kBoxH;
  const CUtensorMap* map = a.gmaps ? &a.gmaps->m[g.lvl][plan.cls] : &maps.m[g.lvl][plan.cls];
  const int z0 = g.n * p.C;
  const float* dout = p.dout + (size_t)k * p.C * 49;
  if (t == 0) {
    mbar_expect_tx(&raw_bar, (uint32_t)min(CCS, p.C) * 196);
    bulk_load(raw, dout, (uint32_t)min(CCS, p.C) * 196, &raw_bar);
  }
  // Wx[x][pw]: summed tap weights of bin pw's two sample columns on footprint column x (zero beyond the footprint)
  for (int i = t; i < BW * 8; i += kThreads) {
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
    wx[i] = w;
  }
  int yr[14];
  float fy[14];
#pragma unroll
  for (int s = 0; s < 14; ++s) {
    yr[s] = ty.i0[s] - ys;
    fy[s] = ty.frac[s];
  }
  // lanes per channel: 8 / 16 / 32 (two column passes when BW > 32)
  const int lpc = BW <= 8 ? 8 : (BW <= 16 ? 16 : 32);
  const int cpw = 32 / lpc;                       // channels per warp pass
  const int xl = lane & (lpc - 1), csub = lane / lpc;
  const int xpasses = (BW + 31) / 32;
  __syncthreads();

  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int s = chunk & 1, c0 = chunk * CCS;
    const int nc = min(CCS, p.C - c0);
    mbar_wait(&raw_bar, (uint32_t)(chunk & 1));
    for (int i = t; i < nc * 49; i += kThreads) {   // dout / 4 into rows of 8 floats
      const int c = i / 49, j = i - c * 49;
      const int ph = j / 7, pw = j - ph * 7;
      pad[c * 56 + ph * 8 + pw] = raw[i] * 0.25f;
    }
    if (warp == 0) bulk_wait_read<1>();             // bulk groups are per thread: the lanes that issued the reduces of
                                                    // chunk - 2 (from stage[s]) wait for their shared-memory reads
    __syncthreads();
    if (t == 0 && chunk + 1 < n_chunks) {
      const int nn = min(CCS, p.C - c0 - CCS);
      mbar_expect_tx(&raw_bar, (uint32_t)nn * 196);
      bulk_load(raw, dout + (size_t)(c0 + CCS) * 49, (uint32_t)nn * 196, &raw_bar);
    }
    float* tile = stage(s);
    for (int cb = warp * cpw; cb < nc; cb += (kThreads / 32) * cpw) {
      const int c = cb + csub;
      const bool cok = c < nc;
      const float* dc = pad + (cok ? c : nc - 1) * 56;
      for (int xp = 0; xp < xpasses; ++xp) {
        const int x = xp * 32 + xl;
        const bool xok = x < BW;
        float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
        if (xok) {
          wa = *reinterpret_cast<const float4*>(wx + x * 8);
          wb = *reinterpret_cast<const float4*>(wx + x * 8 + 4);
        }
        float T[7];
#pragma unroll
        for (int ph = 0; ph < 7; ++ph) {
          const float4 u = *reinterpret_cast<const float4*>(dc + ph * 8);
          const float4 w2 = *reinterpret_cast<const float4*>(dc + ph * 8 + 4);
          float v = wa.x * u.x;
          v = __fmaf_rn(wa.y, u.y, v);
          v = __fmaf_rn(wa.z, u.z, v);
          v = __fmaf_rn(wa.w, u.w, v);
          v = __fmaf_rn(wb.x, w2.x, v);
          v = __fmaf_rn(wb.y, w2.y, v);
          v = __fmaf_rn(wb.z, w2.z, v);
          T[ph] = v;
        }
        float* tcol = tile + (c >> 3) * box_floats + (c & 7) * (kBoxH * BW) + x;
        auto put = [&](int r, float v) {
          if (cok && xok && r >= 0) tcol[(r >> 3) * rb_stride + (r & 7) * BW] = v;  // r < 0: rows above the map
        };
        // walk the 14 sample rows (CTA-uniform program): rows r0, r0 + 1 accumulate in registers, finished rows are
        // stored once, rows no sample touches are stored as zeros (the reduce adds the whole box)
        int rcur = yr[0];
        float ra = 0.f, rb2 = 0.f;
#pragma unroll
        for (int sy = 0; sy < 14; ++sy) {
          const int r0 = yr[sy];
          if (r0 != rcur) {
            put(rcur, ra);
            if (r0 == rcur + 1) {
              ra = rb2;
            } else {
              put(rcur + 1, rb2);
              for (int r = rcur + 2; r < r0; ++r) put(r, 0.f);
              ra = 0.f;
            }
            rb2 = 0.f;
            rcur = r0;
          }
          const float tv = T[sy >> 1];
          ra = __fmaf_rn(1.f - fy[sy], tv, ra);
          rb2 = __fmaf_rn(fy[sy], tv, rb2);
        }
        put(rcur, ra);
        put(rcur + 1, rb2);
        for (int r = rcur + 2; r < rows; ++r) put(r, 0.f);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      const int cbs = (nc + kBoxC - 1) / kBoxC;
      const int nbox = nrb * cbs;
      for (int b = lane; b < nbox; b += 32) {
        const int rb = b / cbs, cbx = b - rb * cbs;
        tma_reduce_add_3d(map, xs, ys + rb * kBoxH, z0 + c0 + cbx * kBoxC, tile + rb * rb_stride + cbx * box_floats);
      }
      bulk_commit();
    }
  }
  if (warp == 0) bulk_wait_read<0>();
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
  bool operator==(const MapKey& o) const { return ptr == o.ptr && H == o.H && W == o.W && BC == o.BC; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (size_t)k.H * 1000003u ^ (size_t)k.W * 10007u ^ (size_t)k.BC * 31u;
  }
};
struct LevelMaps {
  CUtensorMap m[kWClasses];
};

// Builds (or finds) the kWClasses maps of one level.  Returns false when the level cannot be described (alignment).
static bool level_maps(const void* ptr, int H, int W, long long BC, const CUtensorMap** out) {
  static thread_local std::unordered_map<MapKey, LevelMaps, MapKeyHash>* cache = nullptr;
  if (!cache) cache = new std::unordered_map<MapKey, LevelMaps, MapKeyHash>();
  if ((W & 3) || (reinterpret_cast<uintptr_t>(ptr) & 15u) || BC < 1 || BC > 0x7fffffffll) return false;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  MapKey key{ptr, H, W, BC};
  auto it = cache->find(key);
  if (it == cache->end()) {
    if (cache->size() > 256) cache->clear();
    LevelMaps lm;
    for (int c = 0; c < kWClasses; ++c) {
      cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)BC};
      cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
      cuuint32_t box[3] = {(cuuint32_t)(8 * (c + 1)), (cuuint32_t)kBoxH, (cuuint32_t)kBoxC};
      cuuint32_t es[3] = {1, 1, 1};
      alignas(64) CUtensorMap m;
      CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return false;
      lm.m[c] = m;
    }
    it = cache->emplace(key, lm).first;
  }
  *out = it->second.m;
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
static const RoiTmaMaps* debug_global_maps(const RoiTmaMaps* host, cudaStream_t st) {
  if (tma_mode() != 2) return nullptr;
  RoiTmaMaps* d = nullptr;
  if (cudaMalloc(&d, sizeof(RoiTmaMaps)) != cudaSuccess) return nullptr;  // debug only: leaked
  cudaMemcpyAsync(d, host, sizeof(RoiTmaMaps), cudaMemcpyHostToDevice, st);
  return d;
}

// Forward launch through the TMA kernel when at least one level qualifies.  Returns 1 if launched, 0 if the caller
// should use the direct kernel, < 0 on error.
int roi_fwd_tma_launch(const RoiArgs& a, cudaStream_t st) {
  if (!tma_enabled() || a.PH != 7 || a.PW != 7 || a.SH != 2 || a.SW != 2) return 0;
  if (a.lv.n_levels > kTmaLevels || (a.C & 7) || a.C < 8) return 0;
  if ((reinterpret_cast<uintptr_t>(a.out) & 15u)) return 0;
  static thread_local RoiTmaMaps* maps = nullptr;  // 4 KB: filled per call, passed by value
  if (!maps) maps = new RoiTmaMaps();
  RoiTmaArgs ta;
  ta.r = a;
  ta.level_mask = 0;
  for (int l = 0; l < a.lv.n_levels; ++l) {
    const CUtensorMap* lm = nullptr;
    if (level_maps(a.lv.feat[l], a.lv.H[l], a.lv.W[l], (long long)a.B * a.C, &lm)) {
      for (int c = 0; c < kWClasses; ++c) maps->m[l][c] = lm[c];
      ta.level_mask |= 1u << l;
    }
  }
  if (!ta.level_mask) return 0;
  ta.gmaps = debug_global_maps(maps, st);
  if (cudaFuncSetAttribute(roi_align_fwd_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem) != cudaSuccess)
    return set_error(BDET_ECUDA, "roi_align_fwd: cannot reserve %d bytes of shared memory", kFwdSmem);
  BDET_KERNEL("roi_align_fwd_tma_kernel", st, roi_align_fwd_tma_kernel<256><<<a.K, 256, kFwdSmem, st>>>(ta, *maps));
  return 1;
}

// Backward launch: the TMA kernel takes the ROIs `roi_bwd_takes_tma` accepts; *level_mask_out tells the direct kernel
// which levels have tensor maps (0 = nothing was launched and the direct kernel takes every ROI).
int roi_bwd_tma_launch(const RoiArgs& a, cudaStream_t st, unsigned* level_mask_out) {
  *level_mask_out = 0;
  if (!tma_enabled() || a.PH != 7 || a.PW != 7 || a.SH != 2 || a.SW != 2) return 0;
  if (a.lv.n_levels > kTmaLevels || (a.C & 7) || a.C < 8) return 0;
  if ((reinterpret_cast<uintptr_t>(a.dout) & 15u)) return 0;
  static thread_local RoiTmaMaps* maps = nullptr;
  if (!maps) maps = new RoiTmaMaps();
  RoiTmaArgs ta;
  ta.r = a;
  ta.level_mask = 0;
  for (int l = 0; l < a.lv.n_levels; ++l) {
    const CUtensorMap* lm = nullptr;
    if (level_maps(a.lv.dfeat[l], a.lv.H[l], a.lv.W[l], (long long)a.B * a.C, &lm)) {
      for (int c = 0; c < kWClasses; ++c) maps->m[l][c] = lm[c];
      ta.level_mask |= 1u << l;
    }
  }
  if (!ta.level_mask) return 0;
  ta.gmaps = debug_global_maps(maps, st);
  if (cudaFuncSetAttribute(roi_align_bwd_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem) != cudaSuccess)
    return set_error(BDET_ECUDA, "roi_align_bwd: cannot reserve %d bytes of shared memory", kBwdSmem);
  BDET_KERNEL("roi_align_bwd_tma_kernel", st, roi_align_bwd_tma_kernel<256><<<a.K, 256, kBwdSmem, st>>>(ta, *maps));
  *level_mask_out = ta.level_mask;
  return 1;
}

}  // namespace bdet
