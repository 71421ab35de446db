// Shared by roi_align.cu (direct-load kernels, level assignment, host entry points) and roi_tma.cu (TMA kernels):
// ROI geometry and sample tables of F.nn.roi_align(mode="average", aligned) as restated in oracle ASSUMED-6.
#pragma once
#include "common.cuh"

namespace bdet {

constexpr int kRoiThreads = 256;
constexpr int kMaxSamples = 256;  // P * S per axis

struct RoiLevels {
  const float* feat[BDET_MAX_LEVELS];
  float* dfeat[BDET_MAX_LEVELS];
  int H[BDET_MAX_LEVELS], W[BDET_MAX_LEVELS];
  float scale[BDET_MAX_LEVELS];
  int n_levels;
};

struct RoiArgs {
  RoiLevels lv;
  const float* rois;   // (K, 5)
  const int* levels;   // (K) or nullptr
  const float* dout;   // backward
  float* out;          // forward
  int B, C, K, PH, PW, SH, SW;
  float offset;
  int bwd_cap;         // floats of shared accumulation buffer (backward)
  const int* perm;     // optional (K): CTA i works on ROI perm[i] (roi_order_kernel: image / level / tile order for L2)
};
__device__ __forceinline__ int roi_of_cta(const RoiArgs& p, int i) { return p.perm ? __ldg(p.perm + i) : i; }

template <int N>
struct SampleTabT {
  int i0[N];
  float frac[N];
};
using SampleTab = SampleTabT<kMaxSamples>;
using SampleTab14 = SampleTabT<16>;  // 7 bins x 2 samples (the TMA kernels: 4 CTAs / SM need the shared memory)

// sample coordinate table for one axis: coord = start + bin * (p + (i + 0.5) / S)
template <class Tab>
__device__ __forceinline__ void fill_axis(Tab& tab, int P, int S, float start, float bin, int nthreads = kRoiThreads) {
  for (int s = threadIdx.x; s < P * S; s += nthreads) {
    int pidx = s / S, i = s - pidx * S;
    float f = __fdiv_rn((float)i + 0.5f, (float)S);
    float c = start + bin * ((float)pidx + f);
    float fl = floorf(c);
    tab.i0[s] = (int)fl;
    tab.frac[s] = c - fl;
  }
}

struct RoiGeom {
  int n, lvl, H, W;
  float start_w, start_h, bin_w, bin_h;
  bool valid;
};

__device__ __forceinline__ RoiGeom roi_geom(const RoiArgs& p, int k) {
  RoiGeom g;
  const float* r = p.rois + (long long)k * 5;
  g.n = (int)__ldg(r);
  g.lvl = p.levels ? __ldg(p.levels + k) : 0;
  g.valid = g.n >= 0 && g.n < p.B && g.lvl >= 0 && g.lvl < p.lv.n_levels;
  if (!g.valid) g.lvl = 0;
  g.H = p.lv.H[g.lvl];
  g.W = p.lv.W[g.lvl];
  const float sc = p.lv.scale[g.lvl];
  g.start_w = __ldg(r + 1) * sc - p.offset;
  g.start_h = __ldg(r + 2) * sc - p.offset;
  float end_w = __ldg(r + 3) * sc - p.offset;
  float end_h = __ldg(r + 4) * sc - p.offset;
  float roi_w = fmaxf(end_w - g.start_w, 0.f);
  float roi_h = fmaxf(end_h - g.start_h, 0.f);
  g.bin_h = __fdiv_rn(roi_h, (float)p.PH);
  g.bin_w = __fdiv_rn(roi_w, (float)p.PW);
  return g;
}

// Direct-load forward of one ROI (all channels) by the calling CTA: every tap is a predicated __ldg (used for pool
// shapes / levels / ROIs the TMA kernel does not take).  ty / tx must hold the sample tables of the ROI.
template <int TPH, int TPW, int TS, class Tab>
__device__ __forceinline__ void roi_fwd_direct(const RoiArgs& p, int k, const RoiGeom& g, const Tab& ty, const Tab& tx,
                                               int nthreads) {
  const int t = threadIdx.x;
  const int PH = TPH ? TPH : p.PH, PW = TPW ? TPW : p.PW, SH = TS ? TS : p.SH, SW = TS ? TS : p.SW;
  float* out = p.out + (long long)k * p.C * PH * PW;
  const int total = p.C * PH * PW;
  const int H = g.H, W = g.W;
  const float* feat = p.lv.feat[g.lvl] + (long long)g.n * p.C * H * W;
  const float inv_cnt = (float)(SH * SW);
  const int bins = PH * PW;
  for (int o = t; o < total; o += nthreads) {
    const int c = o / bins, bin = o - c * bins;
    const int ph = bin / PW, pw = bin - ph * PW;
    const float* f = feat + (long long)c * H * W;
    float acc = 0.f;
#pragma unroll
    for (int iy = 0; iy < (TS ? TS : 1); ++iy) {
      for (int iy2 = 0; iy2 < (TS ? 1 : SH); ++iy2) {
        const int sy = ph * SH + (TS ? iy : iy2);
        const int y0 = ty.i0[sy], y1 = y0 + 1;
        const float ly = ty.frac[sy];
        const bool y0ok = y0 >= 0 && y0 < H, y1ok = y1 >= 0 && y1 < H;
#pragma unroll
        for (int ix = 0; ix < (TS ? TS : 1); ++ix) {
          for (int ix2 = 0; ix2 < (TS ? 1 : SW); ++ix2) {
            const int sx = pw * SW + (TS ? ix : ix2);
            const int x0 = tx.i0[sx], x1 = x0 + 1;
            const float lx = tx.frac[sx];
            const bool x0ok = x0 >= 0 && x0 < W, x1ok = x1 >= 0 && x1 < W;
            const float tl = (y0ok && x0ok) ? __ldg(f + y0 * W + x0) : 0.f;
            const float tr = (y0ok && x1ok) ? __ldg(f + y0 * W + x1) : 0.f;
            const float bl = (y1ok && x0ok) ? __ldg(f + y1 * W + x0) : 0.f;
            const float br = (y1ok && x1ok) ? __ldg(f + y1 * W + x1) : 0.f;
            const float top = tl + (tr - tl) * lx;
            const float bot = bl + (br - bl) * lx;
            acc += top + (bot - top) * ly;
          }
        }
      }
    }
    out[o] = __fdiv_rn(acc, inv_cnt);
  }
}

// roi_tma.cu: launches the TMA forward when the call qualifies (returns 1), 0 = use the direct kernel, < 0 = error
int roi_fwd_tma_launch(const RoiArgs& a, cudaStream_t st);

// ---- CTA order.  A ROI reads (or updates) its footprint in every channel plane, so ROIs processed at the same time
// should be neighbours: the kernels take ROI perm[blockIdx.x] when the caller passes the permutation of bdet_roi_order
// (roi_of_cta).  (A channel-quarter-major order with 4 CTAs per ROI was measured and dropped: the 4 x per-CTA set-up cost
// more than the locality returned -- forward 0.75 -> 0.98 ms, backward 1.27 -> 1.40 ms; profiles/r02_notes.md.)

// ---- which ROIs the TMA kernels (roi_tma.cu) take: a pure function of the ROI geometry, shared with the direct
// backward kernel, which skips exactly those ROIs -----------------------------------------------------------------
constexpr int kBoxH = 8, kBoxC = 8;     // rows / channels of one TMA box
constexpr int kWClasses = 7;            // box widths 12, 20, ..., 60 floats (4 levels x 7 maps + args < 4 KB of params)
// Box width of class c.  Odd multiples of 4 floats: a box is [8 rows][8 channels][BW], so the 8 channels of a box row start
// BW floats apart = 8 different bank groups (12 c mod 32 = 0, 12, 24, 4, 16, 28, 8, 20); with multiples of 8 floats the
// channels c and c + 4 (or c + 2, or all of them) shared their banks and 60% of the shared-memory wavefronts were replays
// (profiles/r02_notes.md).
__host__ __device__ constexpr int box_width(int cls) { return 8 * cls + 12; }
__host__ __device__ constexpr int width_class(long long fw) { return fw <= 12 ? 0 : (int)((fw - 12 + 7) / 8); }
constexpr int kMaxBoxWidth = box_width(kWClasses - 1);
constexpr int kTmaLevels = 4;
constexpr int kMaxSlots = 8;
constexpr int kFwdMaxRows = 64;
constexpr int kBwdMaxRows = 64;        // footprint rows of a forward stage at most
constexpr int kMaxCCS = 64;             // channels per stage at most

constexpr int kBwdStageBytes = 24 * 1024;  // one gradient tile stage (two per CTA; 3 CTAs per SM)
constexpr int kBwdMaxC = 32;                // channels per stage at most

struct FwdPlan {
  int xs, ys;       // footprint origin (may be negative / beyond the map: TMA fills zeros)
  int cls;          // width class: BW = box_width(cls)
  int nrb;          // rows fetched / reduced (rounded up to even: boxes of 8 / 4 / 2 rows)
  int ccs;          // channels per stage (multiple of 8)
};

// Backward plan.  Returns 0: the direct kernel takes the ROI, 1: the TMA kernel does, 2: the TMA kernel "takes" it
// and has nothing to do (footprint entirely outside the map).  TMA reduces fault on negative box coordinates (loads do
// not; elements beyond the upper bounds are clipped by both -- measured, scripts/tma_probe), so the box origin is clamped
// to the map and the taps in front of it, which the reference drops anyway, never enter the tile.
__device__ __forceinline__ int bwd_plan(const RoiArgs& p, const RoiGeom& g, unsigned level_mask, int x_first, int x_last,
                                        int y_first, int y_last, FwdPlan* pl) {
  const int max_cls = (int)(level_mask >> 16);  // widest box class the TMA backward takes (bits 16..): tuned, see DESIGN.md
  level_mask &= 0xffffu;
  pl->cls = -1;
  pl->nrb = 0;
  pl->ccs = 0;
  const long long fw_full = (long long)x_last + 2 - x_first, fh_full = (long long)y_last + 2 - y_first;
  const bool sane = fabsf(g.start_w) < 1e8f && fabsf(g.start_h) < 1e8f && g.bin_w < 1e7f && g.bin_h < 1e7f && fw_full >= 2 &&
                    fh_full >= 2;
  if (!(sane && ((level_mask >> g.lvl) & 1u))) return 0;
  pl->xs = max(x_first & ~3, 0);  // innermost box coordinate: a multiple of 16 bytes, not negative
  pl->ys = max(y_first, 0);
  const long long fw = (long long)x_last + 2 - pl->xs, fh = (long long)y_last + 2 - pl->ys;
  if (fw < 1 || fh < 1 || pl->xs >= g.W || pl->ys >= g.H) return 2;
  // the backward has its own box widths, 8, 16, ..., 56 floats: its lanes are footprint columns, and 8 / 16 lanes per channel
  // pack 4 / 2 channels into a warp; rows: the footprint rounded up to even (boxes of 8 / 4 / 2 rows)
  if (fw > 8 * (max_cls + 1) || fw > 8 * kWClasses || fh > kBwdMaxRows) return 0;
  const int cls = (int)((fw + 7) / 8) - 1;
  const int nrb = (int)((fh + 1) & ~1ll);  // ROWS
  const long long per_c = (long long)nrb * 8 * (cls + 1) * 4;
  long long ccs = (kBwdStageBytes / per_c) & ~7ll;
  if (ccs > kBwdMaxC) ccs = kBwdMaxC;
  if (ccs > p.C) ccs = p.C;
  if (ccs < kBoxC) return 0;
  pl->cls = cls;
  pl->nrb = nrb;
  pl->ccs = (int)ccs;
  return 1;
}

// first / last sample coordinate of an axis, the same fp32 expressions as fill_axis (s = 0 and s = 13)
__device__ __forceinline__ void axis_ends(float start, float bin, int* first, int* last) {
  const float c0 = start + bin * (0.f + __fdiv_rn(0.5f, 2.f));
  const float c1 = start + bin * (6.f + __fdiv_rn(1.5f, 2.f));
  *first = (int)floorf(c0);
  *last = (int)floorf(c1);
}

__device__ __forceinline__ int roi_bwd_takes_tma(const RoiArgs& p, const RoiGeom& g, unsigned level_mask, FwdPlan* pl) {
  int xf, xl, yf, yl;
  axis_ends(g.start_w, g.bin_w, &xf, &xl);
  axis_ends(g.start_h, g.bin_h, &yf, &yl);
  return bwd_plan(p, g, level_mask, xf, xl, yf, yl, pl);
}

int roi_bwd_tma_launch(const RoiArgs& a, cudaStream_t st, unsigned* level_mask_out);

// non-returning fp32 reduction under a predicate (no branch)
__device__ __forceinline__ void red_add_if(float* addr, float v, bool pred) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p red.global.add.f32 [%0], %1;\n}\n" ::"l"(addr), "f"(v), "r"((int)pred) : "memory");
}

// ---- TMA / mbarrier primitives (sm_90+ PTX; SASS: UTMALDG / UTMAREDG / UBLKCP / SYNCS) ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barriers: `n` threads in total (whole warps) take part, by bar_sync (waits) or bar_arrive (does not)
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// the same with a suspend-time hint: a warp that waits for data sleeps in the barrier unit instead of spinning through the
// issue slots the working warps need (the spin loop was 17% of the forward kernel's instructions)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
// box load of a rank-3 tensor map: coordinates (x, y, z) = (fastest, ..., slowest); out-of-bounds elements arrive as 0
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* map, int x, int y, int z, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
               : "memory");
}
// element-wise fp32 add of a shared-memory box into a rank-3 tensor map (out-of-bounds elements are dropped)
__device__ __forceinline__ void tma_reduce_add_3d(const void* map, int x, int y, int z, const void* smem_src) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(x), "r"(y), "r"(z)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// generic-proxy writes to shared memory must be fenced before the async proxy (TMA store / reduce) reads them
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace bdet
