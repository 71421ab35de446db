// a11: class-aware batched NMS.
// Reference: basedet/layers/common/post_processing.py:17-47 (batched_nms: offset trick) -> F.vision.nms
// (MegEngine NMSKeep: argsort by score, 64-bit overlap masks, serial sweep; oracle ASSUMED-3/5).
//
// Per image (all B images in the same launches):
//   1. sort  : (score desc, index asc) via unique 64-bit keys.  N <= 16384: one CTA; if the caller marks the input as
//              back-to-back runs that are already sorted (the per-level top-k output, bdet_nms_runs) the order comes
//              from a merge by ranking (own position + binary searches in the other runs, promise verified on the
//              device), otherwise from a shared-memory bitonic network; tiled global bitonic network above 16384.
//              The same pass applies the reference's fp32 class offset  boxes + idxs * (max(boxes) + 1)  and gathers
//              the boxes in sorted order.
//   2+3 fused (nms_fused_kernel) when max_output x N is small (every detector head): one CTA per image walks the
//              sorted boxes 64 at a time and tests them against the list of boxes kept so far (shared memory).
//   otherwise, in row chunks of 1024 sorted boxes (nms_chunk_kernel + nms_sweep_kernel per chunk):
//   2. pull  : the chunk's columns are tested against the boxes KEPT by the earlier chunks (a list in global
//              memory) -> removed bits of the chunk; mask: 64x64 tiles of the chunk's own upper triangle; a warp
//              takes a row box, its lanes two column boxes each, two __ballot_sync build the 64-bit suppression
//              word (IoU > thr, IEEE division as the reference).  Nothing runs once max_output boxes are kept.
//   3. sweep : per image, warp 0 resolves one 64-box block at a time from the diagonal words held in registers
//              (only un-suppressed boxes are visited), then the CTA ORs the kept rows' chunk-local words into the
//              `removed` bitmap of the chunk and appends the kept boxes to the list; stops at max_output.
// Rated in pair tests/s (issue bound), not HBM GB/s.
#include "common.cuh"
#include "sortnet.cuh"

namespace bdet {

constexpr int kSmallSortMax = 16384;
constexpr int kSortTile = 4096;
constexpr int kMaxRuns = 32;
constexpr int kChunkBlocks = 16;   // NMS mask/sweep row chunk: 1024 sorted boxes

struct NmsArgs {
  const float* boxes;   // (B, Nmax, 4)
  const float* scores;  // (B, Nmax)
  const void* idxs;     // (B, Nmax) int32 / fp32 or nullptr
  const int* n_dev;     // (B) or nullptr
  const int* run_end;   // (B, n_runs) or nullptr: the image's list is n_runs back-to-back runs, each already in
  int n_runs;           //   (score desc, index asc) order; run r ends (exclusive) at run_end[b][r]
  int idxs_is_float, Nmax, nwords, P;
  float thr;
  int max_out, keep_ld;
  int* order;           // (B, Nmax)
  float4* sboxes;       // (B, Nmax)
  uint32_t* maxc;       // (B) order-encoded max coordinate (large path)
  uint64_t* keys;       // (B, P) (large path)
  uint64_t* mask;       // (B, Nmax, nwords)
  uint64_t* remv_g;     // (B, nwords) columns suppressed by boxes kept in earlier chunks (pull CTAs)
  float4* kbox;         // (B, kcap) boxes kept so far, in keep order
  int kcap, pull_slices, mwords;  // mask rows hold only the chunk-local words: (B, Nmax, mwords)
  int* keep;
  int* keep_count;
};

__device__ __forceinline__ int nms_n(const NmsArgs& p, int b) { return p.n_dev ? max(0, min(p.n_dev[b], p.Nmax)) : p.Nmax; }

__device__ __forceinline__ float class_offset(const NmsArgs& p, long long i, float maxc) {
  if (!p.idxs) return 0.f;
  float c = p.idxs_is_float ? __ldg(reinterpret_cast<const float*>(p.idxs) + i)
                            : (float)__ldg(reinterpret_cast<const int*>(p.idxs) + i);
  return c * (maxc + 1.f);  // post_processing.py:45: idxs * (max_coordinate + 1)
}

__device__ __forceinline__ float4 shifted_box(const NmsArgs& p, long long i, float maxc) {
  float4 bx = ldg4(p.boxes + i * 4);
  if (p.idxs) {
    float off = class_offset(p, i, maxc);
    bx.x += off;  // post_processing.py:46: boxes + offsets.reshape(-1, 1)
    bx.y += off;
    bx.z += off;
    bx.w += off;
  }
  return bx;
}

// ---- small path: one CTA per image does max-reduce, key generation, sort, gather ----------------------
__global__ void __launch_bounds__(1024) nms_sort_small_kernel(const NmsArgs p) {
  extern __shared__ __align__(16) unsigned char raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(raw);
  __shared__ uint32_t smax;
  const int b = blockIdx.x, t = threadIdx.x;
  const int n = nms_n(p, b);
  if (n == 0) return;
  const long long base = (long long)b * p.Nmax;
  if (t == 0) smax = 0u;
  __syncthreads();
  float mc = 0.f;
  if (p.idxs) {
    float m = -CUDART_INF_F;
    for (int i = t; i < n; i += 1024) {
      float4 bx = ldg4(p.boxes + (base + i) * 4);
      m = fmaxf(m, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
    }
    uint32_t w = __reduce_max_sync(0xffffffffu, f2ord(m));
    if ((t & 31) == 0) atomicMax(&smax, w);
    __syncthreads();
    mc = ord2f(smax);
  }
  if (p.run_end) {
    // Pre-sorted runs (the per-level top-k output): rank every element by counting -- its position in its own run
    // plus a binary search in each other run -- instead of sorting.  The promise is verified; a violated one takes
    // the sorting network below.
    __shared__ int rend[kMaxRuns + 1];
    if (t <= p.n_runs) rend[t] = t == 0 ? 0 : __ldg(p.run_end + (long long)b * p.n_runs + t - 1);
    for (int i = t; i < n; i += 1024) keys[i] = make_key(__ldg(p.scores + base + i), (uint32_t)i);
    __syncthreads();
    bool bad = false;
    if (t < p.n_runs) bad = rend[t + 1] < rend[t] || (t == p.n_runs - 1 && rend[t + 1] != n);
    bad = __syncthreads_or(bad);
    if (!bad) {
      for (int i = t; i < n; i += 1024) {
        int r = 0;
        while (rend[r + 1] <= i) ++r;
        if (i > rend[r] && !(keys[i - 1] < keys[i])) bad = true;
      }
      bad = __syncthreads_or(bad);
    }
    if (!bad) {
      // gridDim.y CTAs share an image: each holds all keys (cheap to rebuild) and ranks / gathers its slice of the
      // elements, so the binary searches of one image run on several SMs instead of one
      for (int i = blockIdx.y * 1024 + t; i < n; i += 1024 * gridDim.y) {
        const uint64_t key = keys[i];
        int rank = 0;
        for (int q = 0; q < p.n_runs; ++q) {
          int lo = rend[q], hi = rend[q + 1];
          if (i >= lo && i < hi) {
            rank += i - lo;
          } else {
            const int start = lo;
            while (lo < hi) {  // lower bound: keys are unique
              const int mid = (lo + hi) >> 1;
              if (keys[mid] < key) lo = mid + 1;
              else hi = mid;
            }
            rank += lo - start;
          }
        }
        p.order[base + rank] = i;
        p.sboxes[base + rank] = shifted_box(p, base + i, mc);
      }
      return;
    }
    __syncthreads();
  }
  if (blockIdx.y != 0) return;  // the sorting network is a one-CTA job (every slice CTA reached the same verdict)
  int P = 2;
  while (P < n) P <<= 1;
  for (int i = t; i < P; i += 1024) keys[i] = i < n ? make_key(__ldg(p.scores + base + i), (uint32_t)i) : ~0ull;
  bitonic_sort_smem(keys, P);
  for (int i = t; i < n; i += 1024) {
    int src = (int)(uint32_t)keys[i];
    p.order[base + i] = src;
    p.sboxes[base + i] = shifted_box(p, base + src, mc);
  }
}

// ---- large path ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nms_maxcoord_kernel(const NmsArgs p) {
  const int b = blockIdx.y;
  const int n = nms_n(p, b);
  const long long base = (long long)b * p.Nmax;
  float m = -CUDART_INF_F;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    float4 bx = ldg4(p.boxes + (base + i) * 4);
    m = fmaxf(m, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
  }
  uint32_t w = __reduce_max_sync(0xffffffffu, f2ord(m));
  if ((threadIdx.x & 31) == 0) atomicMax(&p.maxc[b], w);
}

__global__ void __launch_bounds__(1024) nms_tile_sort_kernel(const NmsArgs p) {
  __shared__ uint64_t a[kSortTile];
  const int b = blockIdx.y, t = threadIdx.x;
  const int n = nms_n(p, b);
  const long long tb = (long long)blockIdx.x * kSortTile;
  const long long base = (long long)b * p.Nmax;
  uint64_t* keys = p.keys + (long long)b * p.P;
  for (int i = t; i < kSortTile; i += 1024) {
    long long g = tb + i;
    a[i] = g < n ? make_key(__ldg(p.scores + base + g), (uint32_t)g) : ~0ull;
  }
  for (int size = 2; size <= kSortTile; size <<= 1) bitonic_merge_tail_smem(a, kSortTile, tb, size, size >> 1);
  for (int i = t; i < kSortTile; i += 1024) keys[tb + i] = a[i];
}

__global__ void __launch_bounds__(256) nms_global_step_kernel(const NmsArgs p, long long size, long long stride) {
  const int b = blockIdx.y;
  uint64_t* keys = p.keys + (long long)b * p.P;
  long long i = blockIdx.x * 256ll + threadIdx.x;
  if (i >= (p.P >> 1)) return;
  long long lo = 2 * i - (i & (stride - 1));
  long long hi = lo + stride;
  bool up = ((lo & size) == 0);
  uint64_t x = keys[lo], y = keys[hi];
  if ((x > y) == up) {
    keys[lo] = y;
    keys[hi] = x;
  }
}

__global__ void __launch_bounds__(1024) nms_tile_tail_kernel(const NmsArgs p, long long size) {
  __shared__ uint64_t a[kSortTile];
  const int b = blockIdx.y, t = threadIdx.x;
  const long long tb = (long long)blockIdx.x * kSortTile;
  uint64_t* keys = p.keys + (long long)b * p.P;
  for (int i = t; i < kSortTile; i += 1024) a[i] = keys[tb + i];
  bitonic_merge_tail_smem(a, kSortTile, tb, size, kSortTile >> 1);
  for (int i = t; i < kSortTile; i += 1024) keys[tb + i] = a[i];
}

__global__ void __launch_bounds__(256) nms_gather_kernel(const NmsArgs p) {
  const int b = blockIdx.y;
  const int n = nms_n(p, b);
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const long long base = (long long)b * p.Nmax;
  int src = (int)(uint32_t)p.keys[(long long)b * p.P + i];
  float mc = p.idxs ? ord2f(p.maxc[b]) : 0.f;
  p.order[base + i] = src;
  p.sboxes[base + i] = shifted_box(p, base + src, mc);
}

// ---- suppression mask: nms_overlap (common.cuh) ----------------------------------------------------------------

// One launch per row chunk [blk0, blk0 + cb_n) x 64 sorted boxes, two kinds of CTAs (blockIdx.y):
//   y <  cb_n : PULL  -- column block blk0 + x of the chunk against a slice of the boxes kept by the EARLIER chunks
//               (kept list in global memory); suppressed columns are ORed into remv_g.  A kept box that is never
//               followed by a live column costs nothing, and nothing is computed once max_output boxes are kept.
//   y >= cb_n : MASK  -- 64 x 64 tile (rb = blk0 + y - cb_n, cb = blk0 + x >= rb) of the chunk's own upper triangle.
// Pair tests per image: N x (kept so far) / 1 + chunk^2 / 2 per chunk, instead of (alive rows) x N.
constexpr int kPullSlice = 4096;  // kept boxes per pull CTA

__global__ void __launch_bounds__(256) nms_chunk_kernel(const NmsArgs p, int blk0, int cb_n) {
  const int b = blockIdx.z, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n = nms_n(p, b);
  const int cb = blk0 + blockIdx.x;
  if (cb * 64 >= n) return;
  const int max_out = p.max_out > 0 ? min(p.max_out, p.keep_ld) : p.keep_ld;
  const int count = p.keep_count[b];
  if (count >= max_out) return;
  const float4* sb = p.sboxes + (long long)b * p.Nmax;
  if ((int)blockIdx.y < p.pull_slices) {  // ---- pull
    const int q0 = blockIdx.y * kPullSlice, q1 = min(count, q0 + kPullSlice);
    if (q0 >= q1) return;
    __shared__ uint32_t srem[2];
    if (t < 2) srem[t] = 0u;
    __syncthreads();
    const int col = cb * 64 + (((warp & 1) << 5) | lane), sub = warp >> 1;  // 4 sub-slices of the kept range
    bool sup = false;
    if (col < n) {
      const float4 c = sb[col];
      const float ca = box_area(c);
      const float4* kb = p.kbox + (long long)b * p.kcap;
      for (int q = q0 + sub; q < q1 && !sup; q += 4) {
        const float4 k = kb[q];
        sup = nms_overlap(k, box_area(k), c, ca, p.thr);
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, sup);
    if (lane == 0 && m) atomicOr(&srem[warp & 1], m);
    __syncthreads();
    if (t == 0) {
      const uint64_t w = ((uint64_t)srem[1] << 32) | srem[0];
      if (w) atomicOr(reinterpret_cast<unsigned long long*>(p.remv_g + (long long)b * p.nwords + cb), (unsigned long long)w);
    }
    return;
  }
  // ---- mask tile inside the chunk
  const int rb = blk0 + (int)blockIdx.y - p.pull_slices;
  if (cb < rb || rb * 64 >= n) return;
  __shared__ float4 srow[64];
  __shared__ float sarea[64];
  if (t < 64) {
    int i = rb * 64 + t;
    float4 bx = i < n ? sb[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    srow[t] = bx;
    sarea[t] = box_area(bx);
  }
  const int c0 = cb * 64 + lane, c1 = c0 + 32;
  const float4 b0 = c0 < n ? sb[c0] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b1 = c1 < n ? sb[c1] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float a0 = box_area(b0), a1 = box_area(b1);
  __syncthreads();
  uint64_t* mrow = p.mask + ((long long)b * p.Nmax + rb * 64) * p.mwords + (cb - blk0);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int lr = warp * 8 + r;
    const int i = rb * 64 + lr;
    if (i >= n) break;  // warp-uniform
    const float4 a = srow[lr];
    const float sa = sarea[lr];
    bool p0 = (c0 < n) && (c0 > i) && nms_overlap(a, sa, b0, a0, p.thr);
    bool p1 = (c1 < n) && (c1 > i) && nms_overlap(a, sa, b1, a1, p.thr);
    uint32_t lo = __ballot_sync(0xffffffffu, p0);
    uint32_t hi = __ballot_sync(0xffffffffu, p1);
    if (lane == 0) mrow[(long long)lr * p.mwords] = ((uint64_t)hi << 32) | lo;
  }
}

// ---- sweep -----------------------------------------------------------------------------------------------
constexpr int kSweepThreads = 256;

// Blocks [blk0, blk0 + cb_n) of every image: the chunk's removed words come from the pull CTAs, suppression inside
// the chunk from the chunk-local mask words; kept boxes are appended to the global kept list for the later chunks.
__global__ void __launch_bounds__(kSweepThreads) nms_sweep_kernel(const NmsArgs p, int blk0, int cb_n, int staged) {
  extern __shared__ __align__(16) unsigned char raw[];
  uint64_t* remv = reinterpret_cast<uint64_t*>(raw);  // cb_n words of this chunk (+ the staged mask)
  __shared__ int skept[64];
  __shared__ int snk;
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31;
  const int n = nms_n(p, b);
  const int nblk = (n + 63) >> 6;
  const long long base = (long long)b * p.Nmax;
  const uint64_t* mask = p.mask + base * p.mwords;
  const int max_out = p.max_out > 0 ? min(p.max_out, p.keep_ld) : p.keep_ld;
  int count = p.keep_count[b];
  if (count >= max_out || blk0 >= nblk) return;
  const int blk_end = min(nblk, blk0 + cb_n);
  for (int w = t; w < blk_end - blk0; w += kSweepThreads) remv[w] = p.remv_g[(long long)b * p.nwords + blk0 + w];
  if (staged) {
    // the chunk-local mask (<= 1024 rows x 16 words = 128 KB) is pulled into shared memory once: the block-by-block
    // loop below is a serial chain, and every global access in it would cost an L2 round trip
    uint64_t* sm = remv + cb_n;
    const int rows = min(n, blk_end * 64) - blk0 * 64;
    const uint64_t* src = mask + (long long)blk0 * 64 * p.mwords;
    for (int i = t; i < rows * p.mwords; i += kSweepThreads) sm[i] = src[i];
    mask = sm - (long long)blk0 * 64 * p.mwords;  // same indexing as the global array
  }
  for (int blk = blk0; blk < blk_end && count < max_out; ++blk) {
    __syncthreads();
    const int wl = blk - blk0;
    if (t < 32) {
      const int r0 = blk * 64 + lane, r1 = r0 + 32;
      const uint64_t d0 = r0 < n ? mask[(long long)r0 * p.mwords + wl] : 0ull;
      const uint64_t d1 = r1 < n ? mask[(long long)r1 * p.mwords + wl] : 0ull;
      const int nv = min(64, n - blk * 64);
      const uint64_t valid = nv >= 64 ? ~0ull : ((1ull << nv) - 1ull);
      uint64_t cand = ~remv[wl] & valid;
      int nk = 0;
      const bool hit0 = ((cand >> lane) & 1ull) && (d0 & cand), hit1 = ((cand >> (lane + 32)) & 1ull) && (d1 & cand);
      if (!__any_sync(0xffffffffu, hit0 || hit1)) {
        // no live box of the block suppresses another live one: all of them are kept, in order, without the serial walk
        const int room = max_out - count;
        nk = min(__popcll(cand), room);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = lane + 32 * h;
          const int rank = __popcll(cand & ((1ull << i) - 1ull));
          if (((cand >> i) & 1ull) && rank < room) skept[rank] = i;
        }
      } else {
        while (cand != 0ull && count + nk < max_out) {
          const int i = __ffsll((long long)cand) - 1;
          if (lane == 0) skept[nk] = i;
          ++nk;
          const uint64_t mine = (i < 32) ? d0 : d1;
          const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)mine, i & 31);
          const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(mine >> 32), i & 31);
          cand &= ~(((uint64_t)hi << 32) | lo);
          cand &= ~(1ull << i);
        }
      }
      if (lane == 0) snk = nk;
    }
    __syncthreads();
    const int nk = snk;
    if (t < nk) {
      const int row = blk * 64 + skept[t];
      p.keep[(long long)b * p.keep_ld + count + t] = p.order[base + row];
      p.kbox[(long long)b * p.kcap + count + t] = p.sboxes[base + row];
    }
    count += nk;
    if (count >= max_out || nk == 0) continue;
    for (int w = wl + 1 + t; w < blk_end - blk0; w += kSweepThreads) {
      uint64_t acc = remv[w];
      for (int q = 0; q < nk; ++q) acc |= mask[(long long)(blk * 64 + skept[q]) * p.mwords + w];
      remv[w] = acc;
    }
  }
  if (t == 0) p.keep_count[b] = count;
}

// ---- keep-driven path: nothing in HBM ---------------------------------------------------------------------
// One CTA per image walks the sorted boxes 64 at a time and keeps the boxes kept so far (<= max_output) in shared memory:
//   (a) pull: the block's 64 boxes are tested against the kept list (thread = column x 1/16 of the list) -> removed word;
//   (c) the 64x64 diagonal words of the still-alive rows are built with ballots (2 rows per warp);
//   (d) warp 0 resolves the block from those words (visiting only un-suppressed boxes), stopping at max_output;
//   (e) the newly kept boxes are appended to the list.
// Pair tests: (boxes examined) x (boxes kept so far) -- a block that is never reached because max_output boxes are
// already kept costs nothing, unlike a push form that suppresses all N columns for every kept row.  Decisions are the
// same overlap predicate as the mask kernel, so the keep list is identical.
constexpr int kFusedThreads = 1024;

__global__ void __launch_bounds__(kFusedThreads) nms_fused_kernel(const NmsArgs p) {
  extern __shared__ __align__(16) unsigned char raw[];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int max_out = min(p.max_out > 0 ? min(p.max_out, p.keep_ld) : p.keep_ld, p.Nmax);  // == host-side kmax
  float4* kbox = reinterpret_cast<float4*>(raw);             // max_out: boxes kept so far
  float* karea = reinterpret_cast<float*>(kbox + max_out);   // max_out
  __shared__ float4 srow[64];
  __shared__ float sarea[64];
  __shared__ uint64_t sdiag[64];
  __shared__ int skept[64];
  __shared__ int snk;
  __shared__ uint32_t srem[2];
  const int n = nms_n(p, b);
  const int nblk = (n + 63) >> 6;
  const long long base = (long long)b * p.Nmax;
  const float4* sb = p.sboxes + base;
  int count = 0;
  for (int blk = 0; blk < nblk && count < max_out; ++blk) {
    __syncthreads();  // previous block's shared rows / kept list updates are complete
    const int r0 = blk * 64;
    const int nv = min(64, n - r0);
    const uint64_t valid = nv >= 64 ? ~0ull : ((1ull << nv) - 1ull);
    if (t < 64) {
      const float4 bx = t < nv ? sb[r0 + t] : make_float4(0.f, 0.f, 0.f, 0.f);
      srow[t] = bx;
      sarea[t] = box_area(bx);
      sdiag[t] = 0ull;
    }
    if (t < 2) srem[t] = 0u;
    __syncthreads();
    {  // (a) pull: which of the block's 64 boxes does a box kept earlier suppress?  thread = (column, 1/16 of the kept list)
      const int col = ((warp & 1) << 5) | lane, slice = warp >> 1;
      bool sup = false;
      if (col < nv) {
        const float4 c = srow[col];
        const float ca = sarea[col];
        for (int q = slice; q < count && !sup; q += kFusedThreads / 64) sup = nms_overlap(kbox[q], karea[q], c, ca, p.thr);
      }
      const uint32_t m = __ballot_sync(0xffffffffu, sup);
      if (lane == 0 && m) atomicOr(&srem[warp & 1], m);
    }
    __syncthreads();
    const uint64_t alive = ~(((uint64_t)srem[1] << 32) | srem[0]) & valid;
    if (alive == 0ull) continue;  // CTA-uniform
    {  // (c) diagonal words: warp w -> rows w and w + 32
      const float4 c0 = srow[lane], c1 = srow[lane + 32];
      const float a0 = sarea[lane], a1 = sarea[lane + 32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = warp + 32 * h;
        if (!((alive >> i) & 1ull)) continue;  // warp-uniform
        const float4 a = srow[i];
        const float sa = sarea[i];
        const bool p0 = (lane > i) && (lane < nv) && nms_overlap(a, sa, c0, a0, p.thr);
        const bool p1 = (lane + 32 > i) && (lane + 32 < nv) && nms_overlap(a, sa, c1, a1, p.thr);
        const uint32_t lo = __ballot_sync(0xffffffffu, p0);
        const uint32_t hi = __ballot_sync(0xffffffffu, p1);
        if (lane == 0) sdiag[i] = ((uint64_t)hi << 32) | lo;
      }
    }
    __syncthreads();
    if (warp == 0) {  // (d) serial resolve, all lanes redundantly (uniform)
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
      p.keep[(long long)b * p.keep_ld + count + t] = p.order[base + r0 + i];
      kbox[count + t] = srow[i];  // (e) the kept boxes join the list the later blocks are tested against
      karea[count + t] = sarea[i];
    }
    count += nk;
  }
  if (t == 0) p.keep_count[b] = count;
}

// 64-box blocks per row chunk of the mask / sweep path
static int chunk_blocks(int nblk) { return nblk <= 512 ? kChunkBlocks : ceil_div(nblk, 32); }

struct NmsWs {
  size_t order, sboxes, maxc, keys, remv, kbox, mask, total;
};
static NmsWs nms_ws(int Nmax, int B) {
  NmsWs w;
  const size_t nwords = (size_t)(Nmax + 63) / 64;
  size_t o = 0;
  w.order = o;
  o += align_up((size_t)B * Nmax * 4, 256);
  w.sboxes = o;
  o += align_up((size_t)B * Nmax * 16, 256);
  w.maxc = o;
  o += align_up((size_t)B * 4, 256);
  w.keys = o;
  if (Nmax > kSmallSortMax) o += align_up((size_t)B * next_pow2(Nmax) * 8, 256);
  w.remv = o;
  o += align_up((size_t)B * nwords * 8, 256);
  w.kbox = o;
  o += align_up((size_t)B * Nmax * 16, 256);
  w.mask = o;
  o += (size_t)B * Nmax * chunk_blocks((int)nwords) * 8;  // only the chunk-local words of a row are ever stored
  w.total = o + 256;
  return w;
}

}  // namespace bdet

using namespace bdet;

extern "C" size_t bdet_nms_workspace(int Nmax, int B) {
  if (Nmax <= 0 || B <= 0) return 16;
  return nms_ws(Nmax, B).total;
}

extern "C" int bdet_nms(const float* boxes, const float* scores, const void* idxs, int idxs_is_float, const int* n_dev,
                        int Nmax, int B, float iou_thresh, int max_output, int* keep, int keep_ld, int* keep_count,
                        void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  return bdet_nms_runs(boxes, scores, idxs, idxs_is_float, n_dev, nullptr, 0, Nmax, B, iou_thresh, max_output, keep, keep_ld,
                       keep_count, workspace, workspace_bytes, stream);
}

extern "C" int bdet_nms_runs(const float* boxes, const float* scores, const void* idxs, int idxs_is_float, const int* n_dev,
                             const int* run_end, int n_runs, int Nmax, int B, float iou_thresh, int max_output, int* keep,
                             int keep_ld, int* keep_count, void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(Nmax >= 0 && B >= 0 && keep_ld >= 0, "negative size");
  BDET_REQUIRE(!run_end || (n_runs >= 1 && n_runs <= kMaxRuns), "n_runs must be in [1, 32]");
  if (B == 0) return BDET_OK;
  BDET_REQUIRE(keep_count, "null keep_count");
  cudaStream_t st = as_stream(stream);
  if (Nmax == 0 || keep_ld == 0) {
    BDET_CUDA(cudaMemsetAsync(keep_count, 0, (size_t)B * 4, st));
    return BDET_OK;
  }
  BDET_REQUIRE(boxes && scores && keep, "null argument");
  BDET_REQUIRE(aligned16(boxes), "boxes must be 16-byte aligned");
  BDET_REQUIRE(B <= 65535, "B > 65535");
  if (Nmax > (1 << 22)) return set_error(BDET_EUNSUPPORTED, "bdet_nms: Nmax > 2^22");
  NmsWs w = nms_ws(Nmax, B);
  if (!workspace || workspace_bytes < w.total)
    return set_error(BDET_EWORKSPACE, "bdet_nms: workspace needs %zu bytes", w.total);
  BDET_REQUIRE(aligned16(workspace), "workspace must be 16-byte aligned");
  char* ws = reinterpret_cast<char*>(workspace);
  NmsArgs a;
  a.boxes = boxes;
  a.scores = scores;
  a.idxs = idxs;
  a.n_dev = n_dev;
  a.run_end = run_end;
  a.n_runs = run_end ? n_runs : 0;
  a.idxs_is_float = idxs_is_float;
  a.Nmax = Nmax;
  a.nwords = (Nmax + 63) / 64;
  a.P = next_pow2(Nmax < 2 ? 2 : Nmax);
  a.thr = iou_thresh;
  a.max_out = max_output;
  a.keep_ld = keep_ld;
  a.order = reinterpret_cast<int*>(ws + w.order);
  a.sboxes = reinterpret_cast<float4*>(ws + w.sboxes);
  a.maxc = reinterpret_cast<uint32_t*>(ws + w.maxc);
  a.keys = reinterpret_cast<uint64_t*>(ws + w.keys);
  a.mask = reinterpret_cast<uint64_t*>(ws + w.mask);
  a.remv_g = reinterpret_cast<uint64_t*>(ws + w.remv);
  a.kbox = reinterpret_cast<float4*>(ws + w.kbox);
  a.kcap = Nmax;
  a.mwords = chunk_blocks(a.nwords);
  a.pull_slices = 1;
  a.keep = keep;
  a.keep_count = keep_count;

  if (Nmax <= kSmallSortMax) {
    size_t smem = (size_t)a.P * 8;
    if (smem > 40 * 1024)
      BDET_CUDA(cudaFuncSetAttribute(nms_sort_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int slices = (a.run_end && Nmax >= 2048) ? 8 : 1;
    BDET_KERNEL("nms_sort_small_kernel", st, nms_sort_small_kernel<<<dim3(B, slices), 1024, smem, st>>>(a));
  } else {
    if (a.P < kSortTile) a.P = kSortTile;
    BDET_CUDA(cudaMemsetAsync(a.maxc, 0, (size_t)B * 4, st));
    if (idxs) BDET_KERNEL("nms_maxcoord_kernel", st, nms_maxcoord_kernel<<<dim3(min(ceil_div(Nmax, 256), 64), B), 256, 0, st>>>(a));
    const int tiles = a.P / kSortTile;
    BDET_KERNEL("nms_tile_sort_kernel", st, nms_tile_sort_kernel<<<dim3(tiles, B), 1024, 0, st>>>(a));
    for (long long size = 2ll * kSortTile; size <= a.P; size <<= 1) {
      for (long long stride = size >> 1; stride >= kSortTile; stride >>= 1)
        BDET_KERNEL("nms_global_step_kernel", st, nms_global_step_kernel<<<dim3(ceil_div(a.P / 2, 256), B), 256, 0, st>>>(a, size, stride));
      BDET_KERNEL("nms_tile_tail_kernel", st, nms_tile_tail_kernel<<<dim3(tiles, B), 1024, 0, st>>>(a, size));
    }
    BDET_KERNEL("nms_gather_kernel", st, nms_gather_kernel<<<dim3(ceil_div(Nmax, 256), B), 256, 0, st>>>(a));
  }
  BDET_LAUNCH_CHECK();
  // keep-driven single-CTA path when (boxes that can be kept) x N stays small; otherwise the full bitmask + sweep
  const long long cap = max_output > 0 ? (long long)min(max_output, Nmax) : (long long)Nmax;
  if (cap * (long long)Nmax <= (2ll << 20)) {  // e.g. 100 detections out of 5 000; RPN (1 000 of ~9 000) takes the mask path
    const int kmax = min(max_output > 0 ? min(max_output, keep_ld) : keep_ld, Nmax);
    const size_t fsmem = (size_t)kmax * 20 + 16;
    if (fsmem > 40 * 1024)
      BDET_CUDA(cudaFuncSetAttribute(nms_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
    BDET_KERNEL("nms_fused_kernel", st, nms_fused_kernel<<<B, kFusedThreads, fsmem, st>>>(a));
    BDET_LAUNCH_CHECK();
    return BDET_OK;
  }
  // Row chunks of kChunkBlocks x 64 sorted boxes.  Per chunk one launch tests the chunk's columns against the boxes
  // kept by the earlier chunks (pull) and builds the chunk's own triangular mask; the sweep then resolves the chunk
  // and appends its kept boxes to the list.  Everything stops once max_output boxes are kept.
  const int nblk = a.nwords;
  if (nblk > 65535) return set_error(BDET_EUNSUPPORTED, "bdet_nms: too many 64-box blocks");
  BDET_CUDA(cudaMemsetAsync(a.remv_g, 0, (size_t)B * a.nwords * 8, st));
  BDET_CUDA(cudaMemsetAsync(keep_count, 0, (size_t)B * 4, st));
  const int chunk = a.mwords;
  const int kmax = min(max_output > 0 ? min(max_output, keep_ld) : keep_ld, Nmax);
  a.pull_slices = max(1, ceil_div(kmax, kPullSlice));
  const int staged = chunk <= kChunkBlocks;  // chunk-local mask fits shared memory
  size_t sweep_smem = (size_t)chunk * 8 + (staged ? (size_t)chunk * 64 * chunk * 8 : 0);
  if (sweep_smem > 40 * 1024)
    BDET_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem));
  for (int blk0 = 0; blk0 < nblk; blk0 += chunk) {
    const int rows = min(chunk, nblk - blk0);
    BDET_KERNEL("nms_chunk_kernel", st, nms_chunk_kernel<<<dim3(rows, a.pull_slices + rows, B), 256, 0, st>>>(a, blk0, rows));
    BDET_KERNEL("nms_sweep_kernel", st, nms_sweep_kernel<<<B, kSweepThreads, sweep_smem, st>>>(a, blk0, rows, staged));
  }
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
