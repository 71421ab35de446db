// a10: score filter + per-level top-k.
// Reference: basedet/models/det/retinanet.py:181-191 (sigmoid -> non_zeros(score > thr) -> F.topk desc),
//            basedet/models/det/fcos.py:194-202 (sqrt(sigmoid(cls)*sigmoid(ctr))), basedet/models/det/rpn.py:155.
//
// Order contract (oracle ASSUMED-3): results sorted by (score desc, flat index asc).  Every element gets
// the UNIQUE 64-bit key  (~ord(score) << 32) | index, so "top-k sorted" == "the k smallest keys, ascending";
// a radix select finds the k-th key, a shared-memory bitonic network sorts the <= k survivors.
//
//   stage 1 (score_filter_kernel, HBM-read bound: 4 B/logit, persistent grid): 128-bit logit loads, a raw-logit
//           pre-filter (monotonicity of sigmoid) rejects ~99 % of the elements with one compare; survivors get the
//           exact fp32 score and the `> thr` test, are staged in shared memory and flushed with one atomic per ~1k keys.
//   stage 2 (select_sort_kernel, one 8-CTA cluster per segment): MSB-first 8-bit radix select over the candidates,
//           histograms merged through distributed shared memory, early exit + small-bucket buffering, then the
//           <= k survivors are gathered into cluster rank 0, bitonic-sorted and written out.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <limits>

#include "common.cuh"
#include "sortnet.cuh"

namespace bdet {

namespace cg = cooperative_groups;

constexpr int kSelThreads = 1024;
constexpr int kBufCap = 4096;  // candidate keys buffered in shared memory once a radix bucket is this small
constexpr int kMaxK = 16384;
constexpr int kSelUnroll = 8;

// One (image, level) segment.  Segments may live in different allocations: `start` is an element offset from the
// base pointer handed to the entry point (any fp32 device address is base + 4*start for some start).
struct SegDesc {
  long long start;      // first element, relative to the scores / logits base pointer
  long long key_off;    // first slot of this segment in the candidate-key buffer
  long long ctr_start;  // FCOS: ctrness index of element 0
  int len;
  int tile_start;       // first filter tile of this segment
  int order;            // select stage: cluster c works on segment seg[c].order (longest segments first)
  int hw;               // > 0: the segment is a head output in NCHW order, (A*C, H, W) with hw = H*W (see key_index)
};

struct RawSrc {
  const float* s;
  __device__ __forceinline__ uint64_t key(int i) const { return make_key(__ldg(s + i), (uint32_t)i); }
};
struct KeySrc {
  const uint64_t* k;
  __device__ __forceinline__ uint64_t key(int i) const { return k[i]; }
};

// Every `stride`-th element of another source (threshold estimation).
template <class Src>
struct SampleSrc {
  Src base;
  int stride;
  __device__ __forceinline__ uint64_t key(int i) const { return base.key(i * stride); }
};

// Shared-memory histogram increment.  Score keys are concentrated (few exponents), so most warps hit ONE bin: that
// case costs one atomic per warp; mixed warps fall back to per-lane shared atomics (__match_any_sync and peeling
// the distinct bins one by one are both slower on a spread-out digit).
__device__ __forceinline__ void hist_add(int* hist, int bin, bool active) {
  const uint32_t act = __ballot_sync(0xffffffffu, active);
  if (act == 0u) return;
  const int leader = __ffs(act) - 1;
  const int b0 = __shfl_sync(0xffffffffu, bin, leader);
  if (__all_sync(0xffffffffu, !active || bin == b0)) {
    if ((threadIdx.x & 31) == leader) atomicAdd(&hist[b0], __popc(act));
  } else if (active) {
    atomicAdd(&hist[bin], 1);
  }
}

// warp-aggregated append; returns the slot for this lane (or -1)
__device__ __forceinline__ int append_slot(int* counter, bool active) {
  uint32_t act = __ballot_sync(0xffffffffu, active);
  if (act == 0) return -1;
  int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(act) - 1) base = atomicAdd(counter, __popc(act));
  base = __shfl_sync(0xffffffffu, base, __ffs(act) - 1);
  return active ? base + __popc(act & ((1u << lane) - 1u)) : -1;
}

// ---- stage 2: cluster radix select ---------------------------------------------------------------------------
// One thread-block CLUSTER of kCS CTAs per segment (a single CTA is instruction-bound on one SM: ~250 thread
// instructions per key over the sweeps).  Each CTA owns 1/kCS of the segment's keys and a private 256-bin histogram;
// after a cluster barrier every CTA sums the kCS histograms through distributed shared memory and redundantly picks
// the same bucket, so nothing is broadcast.  Histograms are double-buffered: one cluster barrier per digit.
// The cluster size is a launch attribute (1, 2, 4 or 8): with many segments one CTA each already fills the GPU.
constexpr int kMaxCS = 8;

struct SelSmem {
  int hist[2][256];
  int scan[256];
  int nbuf;
  int nsel;
  int nsurv;
  int bin, below, cnt;
  uint64_t thr;
};

// k-th smallest key (1 <= k <= n) of the whole segment; this CTA sweeps elements [lo, hi).  Keys unique.
// SOLO: the calling CTA works alone on [lo, hi) (no cluster barrier, own histogram only).
template <bool SOLO, class Src>
__device__ uint64_t radix_select(cg::cluster_group& cluster, const Src& src, int lo, int hi, int k, SelSmem& sm, uint64_t* buf) {
  const int t = threadIdx.x;
  const int CS = SOLO ? 1 : (int)cluster.num_blocks();
  bool buffered = false;
  uint64_t prefix = 0, mask = 0;
  int krem = k;
  for (int d = 7; d >= 0; --d) {
    const int shift = 8 * d;
    int* hist = sm.hist[d & 1];
    if (t < 256) hist[t] = 0;
    __syncthreads();
    const int m = buffered ? sm.nbuf : hi - lo;
    // kSelUnroll independent loads in flight per thread: a one-load-per-iteration sweep is pure L2 latency
    for (int i0 = 0; i0 < m; i0 += kSelThreads * kSelUnroll) {
      uint64_t key[kSelUnroll];
      bool in[kSelUnroll];
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const int i = i0 + u * kSelThreads + t;
        in[u] = i < m;
        key[u] = in[u] ? (buffered ? buf[i] : src.key(lo + i)) : 0;
      }
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const bool act = in[u] && ((key[u] & mask) == prefix);
        hist_add(hist, (int)((key[u] >> shift) & 255), act);
      }
    }
    if (SOLO) __syncthreads();
    else cluster.sync();  // every CTA's histogram of this digit is complete
    int tot = 0;
    if (t < 256) {
      if (SOLO) tot = hist[t];
      else
        for (int r = 0; r < CS; ++r) tot += cluster.map_shared_rank(hist, r)[t];
      int v = tot;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, v, s);
        if ((t & 31) >= s) v += o;
      }
      sm.scan[t] = v;
    }
    __syncthreads();
    if (t < 256) {
      int add = 0;
      for (int w = 0; w < (t >> 5); ++w) add += sm.scan[w * 32 + 31];
      int incl = sm.scan[t] + add;
      int excl = incl - tot;
      if (excl < krem && krem <= incl) {
        sm.bin = t;
        sm.below = excl;
        sm.cnt = tot;
      }
    }
    __syncthreads();
    const int bin = sm.bin, cnt = sm.cnt;
    krem -= sm.below;
    prefix |= (uint64_t)bin << shift;
    mask |= (uint64_t)255 << shift;
    if (krem == cnt) return prefix | (shift ? ((1ull << shift) - 1ull) : 0ull);  // whole bucket selected
    if (!buffered && cnt <= kBufCap) {  // cluster-uniform decision; this CTA keeps ITS keys of the bucket
      if (t == 0) sm.nbuf = 0;
      __syncthreads();
      for (int i0 = 0; i0 < m; i0 += kSelThreads * kSelUnroll) {
        uint64_t key[kSelUnroll];
#pragma unroll
        for (int u = 0; u < kSelUnroll; ++u) {
          const int i = i0 + u * kSelThreads + t;
          key[u] = i < m ? src.key(lo + i) : ~0ull;
        }
#pragma unroll
        for (int u = 0; u < kSelUnroll; ++u) {
          const int i = i0 + u * kSelThreads + t;
          const bool act = i < m && ((key[u] & mask) == prefix);
          const int slot = append_slot(&sm.nbuf, act);
          if (act) buf[slot] = key[u];
        }
      }
      buffered = true;
    }
    __syncthreads();
  }
  return prefix;
}

struct SelArgs {
  const float* scores;        // RAW source (may be nullptr)
  const uint64_t* keys;       // candidate keys (filter path)
  const int* cand_count;      // per segment candidate count (filter path) or nullptr
  const SegDesc* seg;         // device (n_seg + 1)
  float* out_vals;
  int* out_idx;
  int* out_count;
  int k, P;                   // P = pow2 >= k
};

// Sampled threshold: for n >> k the exact select over all n keys is replaced by (1) an exact select of rank r in a
// strided sample of kSample keys (every CTA of the cluster does it redundantly: no barrier), r a few sigma above the
// expected rank k * kSample / n, giving a bound T0 with k <= #{key <= T0} <= kBufCap with overwhelming probability;
// (2) ONE sweep, split over the cluster, that appends the keys <= T0 to rank 0's shared memory through DSMEM;
// (3) the exact select among those in rank 0.  The outcome of (2) is checked, and a miss (adversarial data) falls
// back to the full cluster select, so the result is exact either way.
constexpr int kSample = 16384;      // largest sample; the smallest of {kSample/4, kSample} whose bound fits is used
constexpr int kSampleMinN = 32768;  // shorter segments are selected directly

// Append the keys <= T of [lo, hi) to a (possibly remote) shared-memory list.
template <class Src>
__device__ __forceinline__ void gather_le(const Src& src, int lo, int hi, uint64_t T, int* counter, uint64_t* dst, int cap) {
  const int t = threadIdx.x;
  for (int i0 = lo; i0 < hi; i0 += kSelThreads * kSelUnroll) {
    uint64_t key[kSelUnroll];
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int i = i0 + u * kSelThreads + t;
      key[u] = i < hi ? src.key(i) : ~0ull;
    }
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int i = i0 + u * kSelThreads + t;
      const bool act = i < hi && key[u] <= T;
      const int slot = append_slot(counter, act);
      if (act && slot < cap) dst[slot] = key[u];
    }
  }
}

template <bool RAW, class Src>
__device__ void select_sort_body(cg::cluster_group& cluster, const SelArgs& p, const Src& src, int s, int n, int k, long long off,
                                 SelSmem& sm, uint64_t* sortbuf, uint64_t* buf, uint64_t* surv) {
  const int t = threadIdx.x;
  const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  bool sampled = false;
  int r = 0;
  int msamp = kSample;
  if (n > k && n >= kSampleMinN) {
#pragma unroll
    for (int m = kSample / 4; m <= kSample && !sampled; m *= 4) {
      const float mu = (float)k * (float)m / (float)n;
      r = (int)(mu + 5.f * sqrtf(mu) + 16.f);
      sampled = r < m && 1.1f * (float)r * ((float)n / (float)m) <= (float)kBufCap;  // expected survivors + 10 %
      msamp = m;
    }
  }
  // short segments are handled by rank 0 alone; the peers leave before any cluster barrier
  const bool team = n > k && n >= kSampleMinN && (sampled || CS > 1);
  if (!team && rank != 0) return;
  uint64_t T = ~0ull;
  bool done = false;  // sortbuf of rank 0 holds the k selected keys
  if (team) {
    cluster.sync();  // all CTAs of the cluster are resident, rank 0's counters are initialised
    bool full = !sampled;
    if (sampled) {
      const SampleSrc<Src> ss{src, n / msamp};
      const uint64_t T0 = radix_select<true>(cluster, ss, 0, msamp, r, sm, buf);
      const int lo = (int)((long long)n * rank / CS), hi = (int)((long long)n * (rank + 1) / CS);
      int* nsurv0 = cluster.map_shared_rank(&sm.nsurv, 0);
      gather_le(src, lo, hi, T0, nsurv0, cluster.map_shared_rank(surv, 0), kBufCap);
      cluster.sync();
      const int total = *nsurv0;
      cluster.sync();  // rank 0's counter has been read by every peer
      if (total >= k && total <= kBufCap) {
        if (rank != 0) return;
        if (total <= 2048) {
          // few survivors (~1.2-1.8 k for k = 1000): sorting them all is shorter than an exact select + a sort of k
          int P2 = 2;
          while (P2 < total) P2 <<= 1;
          for (int i = total + t; i < P2; i += kSelThreads) surv[i] = ~0ull;
          __syncthreads();
          bitonic_sort_smem(surv, P2);
          for (int i = t; i < k; i += kSelThreads) {
            const uint64_t key = surv[i];
            const uint32_t idx = (uint32_t)key;
            p.out_idx[(long long)s * p.k + i] = (int)idx;
            p.out_vals[(long long)s * p.k + i] = RAW ? __ldg(p.scores + off + idx) : key_score(key);
          }
          return;
        }
        const KeySrc sv{surv};
        T = total == k ? T0 : radix_select<true>(cluster, sv, 0, total, k, sm, buf);
        __syncthreads();
        gather_le(sv, 0, total, T, &sm.nsel, sortbuf, p.P);
        done = true;
      } else {
        full = true;
      }
    }
    if (full) {
      const int lo = (int)((long long)n * rank / CS), hi = (int)((long long)n * (rank + 1) / CS);
      T = radix_select<false>(cluster, src, lo, hi, k, sm, buf);
      gather_le(src, lo, hi, T, cluster.map_shared_rank(&sm.nsel, 0), cluster.map_shared_rank(sortbuf, 0), p.P);
      cluster.sync();  // also keeps every CTA's shared memory alive until no peer reads it any more
      if (rank != 0) return;
      done = true;
    }
  }
  if (!done) {  // rank 0 alone
    if (n > k) T = radix_select<true>(cluster, src, 0, n, k, sm, buf);
    __syncthreads();
    gather_le(src, 0, n, T, &sm.nsel, sortbuf, p.P);
  }
  __syncthreads();
  for (int i = k + t; i < p.P; i += kSelThreads) sortbuf[i] = ~0ull;
  bitonic_sort_smem(sortbuf, p.P);
  for (int i = t; i < k; i += kSelThreads) {
    uint64_t key = sortbuf[i];
    uint32_t idx = (uint32_t)key;
    p.out_idx[(long long)s * p.k + i] = (int)idx;
    p.out_vals[(long long)s * p.k + i] = RAW ? __ldg(p.scores + off + idx) : key_score(key);
  }
}

template <bool RAW>
__global__ void __launch_bounds__(kSelThreads) select_sort_kernel(const SelArgs p) {
  extern __shared__ __align__(16) unsigned char raw[];
  uint64_t* sortbuf = reinterpret_cast<uint64_t*>(raw);  // P (used in cluster rank 0, written by all ranks)
  uint64_t* buf = sortbuf + p.P;                          // kBufCap: small-bucket buffer of the radix select
  uint64_t* surv = buf + kBufCap;                         // kBufCap: survivors of the sampled threshold (rank 0)
  __shared__ SelSmem sm;
  cg::cluster_group cluster = cg::this_cluster();
  const int s = p.seg[blockIdx.x / cluster.num_blocks()].order, t = threadIdx.x;
  const long long off = RAW ? p.seg[s].start : p.seg[s].key_off;
  const int n = RAW ? p.seg[s].len : p.cand_count[s];
  const int k = min(p.k, n);
  if (t == 0) {
    if (cluster.block_rank() == 0) p.out_count[s] = k;
    sm.nsel = 0;
    sm.nsurv = 0;
  }
  if (k == 0) return;  // cluster-uniform
  __syncthreads();
  if (RAW) select_sort_body<RAW>(cluster, p, RawSrc{p.scores + off}, s, n, k, off, sm, sortbuf, buf, surv);
  else select_sort_body<RAW>(cluster, p, KeySrc{p.keys + off}, s, n, k, off, sm, sortbuf, buf, surv);
}

// ------------------------------------------------------------------------------------------ stage 1
__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

struct FilterArgs {
  const float* logits;
  const float* ctr;           // FCOS: one value per C logits
  const SegDesc* seg;         // device (n_seg + 1); seg[n_seg].tile_start = total tiles
  uint64_t* keys;
  int* cand_count;
  float* scores_out;          // optional dense score output (bdet_scores)
  int n_seg, C, mode;
  int na;                     // anchors per position (NCHW segments only)
  float thr, pre;             // pre: raw-logit pre-filter (sigmoid mode)
  // sampled per-segment thresholds (filter_sample_kernel): seg_thr[s] >= thr is a score bound that still leaves >= k
  // candidates with overwhelming probability; seg_redo[s] = 1 marks a segment whose bound turned out too tight and is
  // swept again with the plain threshold (redo_pass).
  float* seg_thr;
  int* seg_redo;              // [n_seg] flags, [n_seg] = "any segment flagged"
  int k, redo_pass;
};

// Candidate index reported for element e of a segment, and the index of its centerness value.
// Plain segments are already in the reference's flattened (h, w, anchor, class) order (permute_to_N_Any_K,
// layers/common/function.py:26-32).  NCHW segments are the head tensor as the network wrote it, (A*C, H, W): the sweep
// stays linear in memory and only the rare survivors are re-indexed, which removes the transpose pass altogether.
template <bool NCHW>
__device__ __forceinline__ void key_index(const FilterArgs& p, const SegDesc& sd, int e, int& idx, long long& ci) {
  if (NCHW) {
    const int ch = e / sd.hw, pos = e - ch * sd.hw;
    const int a = ch / p.C, c = ch - a * p.C;
    idx = (pos * p.na + a) * p.C + c;
    ci = sd.ctr_start + (long long)a * sd.hw + pos;  // centerness (A, H, W)
  } else {
    idx = e;
    ci = sd.ctr_start + e / p.C;
  }
}

constexpr int kFiltThreads = 256;
constexpr int kFiltVec = 4;                               // float4 per thread per tile
constexpr int kFiltTile = 32 * kFiltVec * 4;              // 512 elements: one WARP iteration

// Exact fp32 score of one candidate (same expressions as scores_kernel, so a dense bdet_scores tensor is bit-identical).
__device__ __forceinline__ bool exact_score(const FilterArgs& p, float x, long long ci, float& s, float thr) {
  if (p.mode == BDET_SCORE_SIGMOID) {
    s = sigmoid_f(x);
  } else if (p.mode == BDET_SCORE_FCOS) {
    s = sqrtf(sigmoid_f(x) * sigmoid_f(__ldg(p.ctr + ci)));  // fcos.py:194
  } else {
    s = x;
  }
  return s > thr;
}
__device__ __forceinline__ bool exact_score(const FilterArgs& p, float x, long long ci, float& s) {
  return exact_score(p, x, ci, s, p.thr);
}

// ---- stage 0: sampled threshold ---------------------------------------------------------------------------------
// With thousands of candidates per anchor level and only k = 1000 wanted (FCOS at the 0.05 threshold: ~28 % of all logits
// pass), exact-scoring every candidate is the cost of the filter.  One CTA per segment scores a strided sample of kFSample
// elements exactly, takes the r-th largest sample score t (r = 2 * k * sample / n + kFMargin, a CTA-wide radix select over
// the fp32 bits) and publishes seg_thr = max(thr, t): the expected number of elements above t is r * n / sample >= 2 k, so
// the sweep exact-scores a few thousand elements per segment instead of hundreds of thousands.  The final top-k is
// unchanged as long as >= k elements pass; filter_check_kernel verifies that and flags the (rare) segment that has to be
// swept again with the plain threshold, so the result is exact either way.
constexpr int kFSample = 8192;
constexpr int kFSampleMin = 16 * kFSample;  // shorter segments are not sampled
constexpr int kFMargin = 16;
constexpr int kFSThreads = 512;             // each thread scores kFSample / kFSThreads elements and keeps their maximum

__global__ void __launch_bounds__(kFSThreads) filter_sample_kernel(const FilterArgs p) {
  __shared__ float gmax[kFSThreads];
  __shared__ int s_above;
  const int s = blockIdx.x, t = threadIdx.x;
  const SegDesc sd = p.seg[s];
  if (t == 0) {
    p.seg_redo[s] = 0;
    if (s == 0) {
      p.seg_redo[p.n_seg] = 0;
      p.seg_redo[p.n_seg + 1] = 0;  // arrival ticket of the sweep (the last warp runs the check)
    }
  }
  const long long want = sd.len >= kFSampleMin ? (2ll * p.k * kFSample + sd.len - 1) / sd.len + kFMargin : 0;
  if (sd.len < kFSampleMin || p.k <= 0 || p.mode == BDET_SCORE_RAW || want > kFSThreads / 2) {
    if (t == 0) p.seg_thr[s] = p.thr;
    return;
  }
  const int stride = sd.len / kFSample;
  if (t == 0) s_above = 0;
  __syncthreads();
  int above = 0;
  float mx = 0.f;  // scores are >= 0; elements at or below the threshold can never be the bound
  constexpr int kPer = kFSample / kFSThreads;
  float xs[kPer], cs[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {  // all the (strided, uncached) loads first: one round trip instead of kPer
    const int i = t + j * kFSThreads;
    const int e = i * stride + (i & 7);  // a little jitter so that a stride that is a multiple of C still visits every class
    int idx;
    long long ci;
    if (sd.hw > 0) key_index<true>(p, sd, e, idx, ci);
    else key_index<false>(p, sd, e, idx, ci);
    xs[j] = __ldg(p.logits + sd.start + e);
    cs[j] = p.mode == BDET_SCORE_FCOS ? __ldg(p.ctr + ci) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    float sc;
    if (p.mode == BDET_SCORE_SIGMOID) sc = sigmoid_f(xs[j]);
    else sc = sqrtf(sigmoid_f(xs[j]) * sigmoid_f(cs[j]));  // fcos.py:194, the same expression as exact_score
    if (sc > p.thr) {
      mx = fmaxf(mx, sc);
      ++above;
    }
  }
  gmax[t] = mx;
  atomicAdd(&s_above, above);
  __syncthreads();
  // Too few sample elements above the threshold for the bound to pay: keep the plain threshold.
  if (want > s_above || (long long)s_above * stride < 8ll * p.k) {
    if (t == 0) p.seg_thr[s] = p.thr;
    return;
  }
  // The `want`-th largest of the per-thread maxima: >= `want` distinct sample elements are at or above it, so it is a
  // (slightly conservative) stand-in for the `want`-th largest sample.  Rank by counting (ties broken by thread index).
  int rank = 0;
  for (int j = 0; j < kFSThreads; ++j) {
    const float o = gmax[j];
    rank += (o > mx) || (o == mx && j < t);
  }
  if (rank == (int)want - 1) p.seg_thr[s] = fmaxf(p.thr, mx);
}

// Run by the last warp of the sweep to finish (arrival ticket): a sampled bound that left fewer than k candidates is
// withdrawn and the segment flagged for the second sweep.
__device__ __forceinline__ void filter_check(const FilterArgs& p, int lane) {
  int any = 0;
  for (int s = lane; s < p.n_seg; s += 32) {
    if (p.seg_thr[s] > p.thr && atomicAdd(p.cand_count + s, 0) < p.k) {
      p.seg_thr[s] = p.thr;
      p.cand_count[s] = 0;
      p.seg_redo[s] = 1;
      any = 1;
    }
  }
  if (__any_sync(0xffffffffu, any) && lane == 0) p.seg_redo[p.n_seg] = 1;
}

// ---- stage 1: warp-autonomous streaming filter --------------------------------------------------------------
// HBM-read bound: every logit is read once (128-bit loads) and rejected by ONE compare against a raw-logit bound
// (monotonicity of sigmoid; for FCOS the bound is per position: sigmoid(x) * sigmoid(ctr) > thr^2  <=>
// x > logit(thr^2 / sigmoid(ctr))).  Each WARP walks its own 512-element tiles (grid stride, 2 KB in flight per warp,
// 32 warps per SM) and keeps two private shared-memory lists: pre-filter survivors (x, index)
// and finished keys.  A rare per-lane event is a frequent per-warp event, so survivors are only RECORDED under the
// ballot (no divergent slow path); the ~150-instruction exact scoring runs 32 survivors at a time with full lanes.
// Keys leave the SM ~100 at a time with one global atomic.  No CTA barrier anywhere.
constexpr int kWarpsPerCta = kFiltThreads / 32;
constexpr int kSurv = 64;       // survivor list per warp (drained whenever >= 32)
constexpr int kKeys = 160;      // key staging per warp (flushed whenever >= 96)
constexpr int kWTab = 64;       // FCOS per-position bounds of one warp tile (needs kFiltTile / C + 2 <= 64)

struct WarpLists {
  float sx[kSurv];
  int se[kSurv];
  uint64_t keys[kKeys];
  float tab[kWTab];
};

template <bool VEC>
__device__ __forceinline__ void load_tile(const float* src, int n, int e0, int lane, float4 (&v)[kFiltVec]) {
#pragma unroll
  for (int j = 0; j < kFiltVec; ++j) {
    const int e = e0 + (j * 32 + lane) * 4;
    const float ninf = -CUDART_INF_F;
    if (VEC && e + 3 < n) {
      v[j] = __ldcs(reinterpret_cast<const float4*>(src + e));
    } else {  // ragged tail / unaligned segment
      v[j].x = e < n ? __ldcs(src + e) : ninf;
      v[j].y = e + 1 < n ? __ldcs(src + e + 1) : ninf;
      v[j].z = e + 2 < n ? __ldcs(src + e + 2) : ninf;
      v[j].w = e + 3 < n ? __ldcs(src + e + 3) : ninf;
    }
  }
}

template <bool VEC, bool NCHW>
__global__ void __launch_bounds__(kFiltThreads, 4) score_filter_kernel(const FilterArgs p, int total_tiles) {
  __shared__ WarpLists lists[kWarpsPerCta];
  const int lane = threadIdx.x & 31;
  WarpLists& L = lists[threadIdx.x >> 5];
  const uint32_t lt = (1u << lane) - 1u;
  const int nwarps = gridDim.x * kWarpsPerCta;
  int tile = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (tile >= total_tiles) return;
  int nsv = 0, nk = 0;  // warp-uniform list lengths
  int s = 0;
  while (s + 1 < p.n_seg && p.seg[s + 1].tile_start <= tile) ++s;
  SegDesc sd = p.seg[s];
  const bool tab_mode = !NCHW && p.mode == BDET_SCORE_FCOS && (kFiltTile / p.C + 2) <= kWTab;
  if (p.redo_pass && p.seg_redo[p.n_seg] == 0) return;  // the usual case: no segment was flagged
  // per-segment score bound and the raw-logit pre-filter that goes with it (sigmoid(x) > t  <=>  x > logit(t))
  float thr_s = p.thr, pre_s = p.pre;
  auto seg_bounds = [&]() {
    thr_s = p.seg_thr ? p.seg_thr[s] : p.thr;
    pre_s = p.pre;
    if (thr_s > p.thr && p.mode != BDET_SCORE_RAW) {
      const float q = p.mode == BDET_SCORE_FCOS ? thr_s * thr_s : thr_s;
      if (q < 1.f) {
        const float l = logf(__fdiv_rn(q, 1.f - q));
        pre_s = fmaxf(p.pre, l - 1e-3f * fmaxf(fabsf(l), 1.f));
      }
    }
  };
  seg_bounds();

  auto flush_keys = [&]() {  // warp-uniform
    if (nk == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(p.cand_count + s, nk);
    base = __shfl_sync(0xffffffffu, base, 0);
    uint64_t* dst = p.keys + sd.key_off + base;
    for (int i = lane; i < nk; i += 32) dst[i] = L.keys[i];
    __syncwarp();
    nk = 0;
  };
  auto drain = [&](int count) {  // exact-score the first `count` survivors (count <= nsv), 32 at a time
    for (int i0 = 0; i0 < count; i0 += 32) {
      const int i = i0 + lane;
      bool ok = false;
      float sc = 0.f;
      int e = 0;
      if (i < count) {
        long long ci;
        key_index<NCHW>(p, sd, L.se[i], e, ci);
        ok = exact_score(p, L.sx[i], ci, sc, thr_s);
      }
      const uint32_t m = __ballot_sync(0xffffffffu, ok);
      if (ok) L.keys[nk + __popc(m & lt)] = make_key(sc, (uint32_t)e);
      nk += __popc(m);
      __syncwarp();
      if (nk >= kKeys - 64) flush_keys();
    }
    const int rem = nsv - count;  // < 32: move the tail to the front
    float tx = 0.f;
    int te = 0;
    if (lane < rem) {
      tx = L.sx[count + lane];
      te = L.se[count + lane];
    }
    __syncwarp();
    if (lane < rem) {
      L.sx[lane] = tx;
      L.se[lane] = te;
    }
    __syncwarp();
    nsv = rem;
  };

  while (true) {
    const int e0 = (tile - sd.tile_start) * kFiltTile;
    const int pos0 = e0 / p.C;
    const bool skip = p.redo_pass && p.seg_redo[s] == 0;  // second pass: only the flagged segments
    // this tile's loads first (latency is hidden by the other ~32 resident warps, each with 2 KB in flight)
    float4 cur[kFiltVec];
    if (!skip) load_tile<VEC>(p.logits + sd.start, sd.len, e0, lane, cur);
    if (tab_mode && !skip) {
      const int npos = min((min(e0 + kFiltTile, sd.len) - 1) / p.C - pos0 + 1, kWTab);
      for (int i = lane; i < npos; i += 32) {
        const float sc = sigmoid_f(__ldg(p.ctr + sd.ctr_start + pos0 + i));
        const float q = __fdiv_rn(thr_s * thr_s, sc);  // need sigmoid(x) > q
        float bound = CUDART_INF_F;
        if (!(thr_s > 0.f)) bound = -CUDART_INF_F;
        else if (q < 1.f) bound = logf(__fdiv_rn(q, 1.f - q)) - 0.01f;  // generous margin: survivors are re-tested exactly
        L.tab[i] = bound;
      }
      __syncwarp();
    }
    const int next = tile + nwarps;
    const bool more = next < total_tiles;
    int ns = s;
    if (more)
      while (ns + 1 < p.n_seg && p.seg[ns + 1].tile_start <= next) ++ns;
#pragma unroll
    for (int j = 0; j < kFiltVec; ++j) {
      if (skip) break;
      const int e = e0 + (j * 32 + lane) * 4;
      const float xs[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
      float bd[4] = {pre_s, pre_s, pre_s, pre_s};
      if (tab_mode) {
        const int q0 = e / p.C - pos0;
        const int r0 = e - (q0 + pos0) * p.C;
#pragma unroll
        for (int q = 0; q < 4; ++q) bd[q] = L.tab[min(q0 + (r0 + q >= p.C), kWTab - 1)];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool pass = xs[q] > bd[q];
        const uint32_t m = __ballot_sync(0xffffffffu, pass);
        if (m) {  // warp-uniform
          if (pass) {
            const int slot = nsv + __popc(m & lt);
            L.sx[slot] = xs[q];
            L.se[slot] = e + q;
          }
          nsv += __popc(m);
          __syncwarp();
          if (nsv >= 32) drain(nsv & ~31);
        }
      }
    }
    if (!more || ns != s) {  // leaving this segment: finish its survivors and keys
      drain(nsv);
      flush_keys();
    }
    if (!more) break;
    tile = next;
    if (ns != s) {
      s = ns;
      sd = p.seg[s];
      seg_bounds();
    }
  }
  if (p.seg_thr && !p.redo_pass) {  // the last warp to finish the first sweep verifies the sampled bounds
    __threadfence();
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(p.seg_redo + p.n_seg + 1, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket == min(nwarps, total_tiles) - 1) {
      __threadfence();
      filter_check(p, lane);
    }
  }
}

// Dense scores (so tests can hand bit-identical scores to the oracle).
__global__ void __launch_bounds__(256) scores_kernel(const float* __restrict__ logits, const float* __restrict__ ctr, int C, long long n,
                                                     int mode, float* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = __ldg(logits + i), s;
  if (mode == BDET_SCORE_SIGMOID) s = sigmoid_f(x);
  else if (mode == BDET_SCORE_FCOS) s = sqrtf(sigmoid_f(x) * sigmoid_f(__ldg(ctr + i / C)));
  else s = x;
  out[i] = s;
}

struct TopkWs {
  SegDesc* seg;
  int* cand_count;
  float* seg_thr;
  int* seg_redo;
  uint64_t* keys;
  size_t bytes;
};
static TopkWs carve_ws(void* base, int64_t total, int n_seg, bool with_keys) {
  TopkWs w;
  char* p = reinterpret_cast<char*>(base);
  size_t o = 0;
  w.seg = reinterpret_cast<SegDesc*>(p + o);
  o += align_up((size_t)(n_seg + 1) * sizeof(SegDesc), 256);
  w.cand_count = reinterpret_cast<int*>(p + o);
  o += align_up((size_t)n_seg * 4, 256);
  w.seg_thr = reinterpret_cast<float*>(p + o);
  o += align_up((size_t)n_seg * 4, 256);
  w.seg_redo = reinterpret_cast<int*>(p + o);
  o += align_up((size_t)(n_seg + 2) * 4, 256);
  w.keys = reinterpret_cast<uint64_t*>(p + o);
  if (with_keys) o += align_up((size_t)total * 8, 256);
  w.bytes = o + 256;
  return w;
}

constexpr int kMaxSeg = 4096;

// Host-side descriptor table; returns total length (or -1 after set_error).
static int64_t build_segments(SegDesc* host, const int64_t* start, const int64_t* len, const int64_t* ctr_start, int C,
                              int n_seg, bool* vec_ok, const char* who, const int* hw = nullptr) {
  int64_t total = 0;
  int tiles = 0;
  for (int s = 0; s < n_seg; ++s) {
    if (len[s] < 0 || len[s] > 0x7fffffffLL) {
      set_error(BDET_EINVAL, "%s: segment length out of range", who);
      return -1;
    }
    host[s].start = start[s];
    host[s].key_off = total;
    host[s].ctr_start = ctr_start ? ctr_start[s] : start[s] / C;
    host[s].len = (int)len[s];
    host[s].hw = hw ? hw[s] : 0;
    host[s].tile_start = tiles;
    if (start[s] % 4 != 0) *vec_ok = false;
    total += len[s];
    tiles += ceil_div(len[s], kFiltTile);
  }
  host[n_seg].start = host[n_seg].key_off = host[n_seg].ctr_start = 0;
  host[n_seg].len = 0;
  host[n_seg].hw = 0;
  host[n_seg].tile_start = tiles;
  // select stage launch order: longest segments first (they finish last otherwise)
  int order[kMaxSeg];
  for (int s = 0; s < n_seg; ++s) order[s] = s;
  std::stable_sort(order, order + n_seg, [&](int a, int b) { return len[a] > len[b]; });
  for (int s = 0; s < n_seg; ++s) host[s].order = order[s];
  host[n_seg].order = 0;
  return total;
}

// The descriptor table travels to the device as KERNEL PARAMETERS (<= 96 descriptors = 3.8 KB per launch) instead of a
// cudaMemcpyAsync from pageable host memory: no staging copy, no hidden host synchronisation, and the whole call can be
// captured into a CUDA graph.
constexpr int kSegPerUpload = 96;
struct SegBatch {
  SegDesc d[kSegPerUpload];
};
__global__ void seg_upload_kernel(SegDesc* dst, const SegBatch b, int n) {
  const int i = threadIdx.x;
  if (i < n) dst[i] = b.d[i];
}
static int upload_segments(SegDesc* dev, const SegDesc* host, int count, cudaStream_t st) {
  for (int o = 0; o < count; o += kSegPerUpload) {
    SegBatch b;
    const int n = std::min(kSegPerUpload, count - o);
    for (int i = 0; i < n; ++i) b.d[i] = host[o + i];
    seg_upload_kernel<<<1, kSegPerUpload, 0, st>>>(dev + o, b, n);
  }
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

template <bool RAW>
static int launch_select(const SelArgs& a, int n_seg, cudaStream_t st) {
  size_t smem = (size_t)(a.P + 2 * kBufCap) * 8;
  if (smem > 40 * 1024)  // static shared memory (SelSmem) counts against the 48 KB default limit too
    BDET_CUDA(cudaFuncSetAttribute(select_sort_kernel<RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // cluster size: the kernel runs one 1024-thread CTA per SM, and a second wave costs more than wider clusters save
  // (measured: 5-40 segments -> 8, 80 -> 4, 320 -> 1)
  int cs = kMaxCS;
  while (cs > 1 && (long long)n_seg * cs * 10 > 22LL * sm_count()) cs >>= 1;
  if (const char* e = getenv("BDET_SELECT_CLUSTER")) {  // tuning / debugging knob
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) cs = v;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_seg * cs));
  cfg.blockDim = dim3(kSelThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = cudaSuccess;
  BDET_KERNEL("select_sort_kernel", st, err = cudaLaunchKernelEx(&cfg, select_sort_kernel<RAW>, a));
  BDET_CUDA(err);
  return BDET_OK;
}

// ---- sample_labels ----------------------------------------------------------------------------
// basedet/layers/common/sampling.py:7-30 with the random draw made explicit (one uniform variate per element, the
// reference's uniform(size=num_valid) being the variates of the selected positions): keep `num_samples` of the
// elements equal to `value`, set the ones with the LARGEST variates to `ignore`.  F.topk(random_tensor, k < 0) there
// selects the |k| largest of a tensor that is 0 outside the mask, ordered (value desc, index asc) -- exactly the
// |k| smallest keys of this file, so the same cluster radix select finds the cut.
struct SampleArgs {
  int* labels;          // (B, A), in place
  const float* noise;   // (B, A)
  const int* ns_dev;    // (B) per-image num_samples or nullptr
  int A, value, ignore, ns_const;
};

struct NoiseSrc {
  const int* lab;
  const float* noise;
  int value;
  __device__ __forceinline__ uint64_t key(int i) const {
    return make_key(__ldg(lab + i) == value ? __ldg(noise + i) : 0.f, (uint32_t)i);  // sampling.py:24-25
  }
};

// Reversed keys of the elements that can be removed without touching the zero-valued remainder of random_tensor:
// ascending order = (variate asc, index desc), everything else is +inf.
struct KeepSrc {
  const int* lab;
  const float* noise;
  int value;
  __device__ __forceinline__ uint64_t key(int i) const {
    const float v = __ldg(noise + i);
    return (__ldg(lab + i) == value && v > 0.f) ? ~make_key(v, (uint32_t)i) : ~0ull;
  }
};

// k-th smallest key of a segment whose `n_finite` keys below ~0ull are known, delivered to EVERY CTA of the cluster.
// Few finite keys: they are gathered into rank 0 and selected there; many: sampled threshold as in select_sort_body
// (checked, falls back); otherwise the full cluster select.  All CTAs must call it (cluster-uniform arguments).
template <class Src>
__device__ uint64_t cluster_select_kth(cg::cluster_group& cluster, const Src& src, int n, int n_finite, int k, SelSmem& sm,
                                       uint64_t* buf, uint64_t* surv) {
  const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  const int lo = (int)((long long)n * rank / CS), hi = (int)((long long)n * (rank + 1) / CS);
  bool gather = n_finite <= kBufCap;
  uint64_t T0 = ~0ull - 1ull;  // every finite key
  if (!gather && n >= kSampleMinN) {
    bool sampled = false;
    int r = 0, msamp = kSample;
#pragma unroll
    for (int m = kSample / 4; m <= kSample && !sampled; m *= 4) {
      const float mu = (float)k * (float)m / (float)n;
      r = (int)(mu + 5.f * sqrtf(mu) + 16.f);
      sampled = r < m && 1.1f * (float)r * ((float)n / (float)m) <= (float)kBufCap;
      msamp = m;
    }
    if (sampled) {
      const SampleSrc<Src> ss{src, n / msamp};
      T0 = radix_select<true>(cluster, ss, 0, msamp, r, sm, buf);
      gather = T0 != ~0ull;  // the sample ran out of finite keys: no usable bound
    }
  }
  if (gather) {
    if (threadIdx.x == 0) sm.nsurv = 0;
    cluster.sync();
    int* nsurv0 = cluster.map_shared_rank(&sm.nsurv, 0);
    gather_le(src, lo, hi, T0, nsurv0, cluster.map_shared_rank(surv, 0), kBufCap);
    cluster.sync();
    const int total = *nsurv0;
    if (total >= k && total <= kBufCap) {
      if (rank == 0) {
        const KeySrc sv{surv};
        const uint64_t T = total == k ? T0 : radix_select<true>(cluster, sv, 0, total, k, sm, buf);
        if (threadIdx.x == 0) sm.thr = T;
      }
      cluster.sync();
      const uint64_t T = *cluster.map_shared_rank(&sm.thr, 0);
      cluster.sync();  // rank 0's shared memory has been read by every peer
      return T;
    }
    cluster.sync();  // peers have read the counter before anything reuses it
  }
  const uint64_t T = radix_select<false>(cluster, src, lo, hi, k, sm, buf);
  cluster.sync();  // a peer may still be summing this CTA's last histogram
  return T;
}

__global__ void __launch_bounds__(kSelThreads) sample_labels_kernel(const SampleArgs p) {
  extern __shared__ __align__(16) unsigned char raw[];
  uint64_t* buf = reinterpret_cast<uint64_t*>(raw);  // kBufCap
  uint64_t* surv = buf + kBufCap;                     // kBufCap
  __shared__ SelSmem sm;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  const int b = blockIdx.x / CS, t = threadIdx.x;
  int* lab = p.labels + (long long)b * p.A;
  const float* nz = p.noise + (long long)b * p.A;
  const int lo = (int)((long long)p.A * rank / CS), hi = (int)((long long)p.A * (rank + 1) / CS);
  if (t == 0) {
    sm.nsel = 0;
    sm.nbuf = 0;
  }
  __syncthreads();
  int mine = 0, pos = 0;
  for (int i = lo + t; i < hi; i += kSelThreads) {
    const bool m = lab[i] == p.value;  // sampling.py:19-20
    mine += m;
    pos += m && nz[i] > 0.f;
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  pos = __reduce_add_sync(0xffffffffu, pos);
  if ((t & 31) == 0 && mine) {
    atomicAdd(&sm.nsel, mine);
    atomicAdd(&sm.nbuf, pos);
  }
  cluster.sync();
  int num_valid = 0, P = 0;
  for (int r = 0; r < CS; ++r) {
    num_valid += *cluster.map_shared_rank(&sm.nsel, r);
    P += *cluster.map_shared_rank(&sm.nbuf, r);
  }
  cluster.sync();  // every peer has read this CTA's counts
  const int ns = max(p.ns_dev ? p.ns_dev[b] : p.ns_const, 0);
  if (num_valid <= ns) return;  // :21-22, cluster-uniform
  const int k = num_valid - ns;  // :27: the k LARGEST entries of random_tensor are dropped
  if (k <= P) {
    // all k of them are selected elements with a positive variate, so the zero-valued rest of random_tensor never
    // enters the order: keep the P - k SMALLEST positive variates instead -- a short select however many are dropped
    const int keep = P - k;
    const KeepSrc ksrc{lab, nz, p.value};
    uint64_t T = 0ull;  // keep == 0: nothing below it
    if (keep > 0) T = cluster_select_kth(cluster, ksrc, p.A, P, keep, sm, buf, surv);
    for (int i = lo + t; i < hi; i += kSelThreads) {
      const uint64_t key = ksrc.key(i);
      if (key != ~0ull && (keep == 0 || key > T)) lab[i] = p.ignore;  // :29
    }
    return;
  }
  // more entries to drop than positive variates (zero / negative variates in play): the literal top-k of random_tensor
  const NoiseSrc src{lab, nz, p.value};
  const uint64_t T = radix_select<false>(cluster, src, lo, hi, k, sm, buf);
  // :29 -- a CTA rewrites only the range it alone reads, so a faster peer cannot disturb a slower one's keys
  for (int i = lo + t; i < hi; i += kSelThreads)
    if (src.key(i) <= T) lab[i] = p.ignore;
  cluster.sync();  // a peer may still be summing this CTA's last histogram: shared memory must outlive that read
}

}  // namespace bdet

using namespace bdet;

extern "C" int bdet_sample_labels(int* labels, const float* noise, int A, int B, int label_value, int ignore_label,
                                  int num_samples, const int* num_samples_dev, bdet_stream_t stream) {
  BDET_REQUIRE(A >= 0 && B >= 0, "negative size");
  if (A == 0 || B == 0) return BDET_OK;
  BDET_REQUIRE(labels && noise, "null argument");
  BDET_REQUIRE(num_samples_dev || num_samples >= 0, "negative num_samples");
  SampleArgs a{labels, noise, num_samples_dev, A, label_value, ignore_label, num_samples};
  cudaStream_t st = as_stream(stream);
  const size_t smem = (size_t)2 * kBufCap * 8;
  BDET_CUDA(cudaFuncSetAttribute(sample_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kSelThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // widest cluster that still runs all B images in ONE wave (a 1024-thread CTA owns an SM, and only ~15 clusters of 8
  // fit the GPCs: a 16th image would wait for a second wave and double the kernel time)
  int cs = kMaxCS;
  for (; cs > 1; cs >>= 1) {
    if (A < cs * 4096) continue;
    attr[0].val.clusterDim.x = (unsigned)cs;
    cfg.gridDim = dim3((unsigned)(B * cs));
    int active = 0;
    if (cudaOccupancyMaxActiveClusters(&active, sample_labels_kernel, &cfg) == cudaSuccess && active >= B) break;
  }
  cudaGetLastError();
  attr[0].val.clusterDim.x = (unsigned)cs;
  cfg.gridDim = dim3((unsigned)(B * cs));
  cudaError_t err = cudaSuccess;
  BDET_KERNEL("sample_labels_kernel", st, err = cudaLaunchKernelEx(&cfg, sample_labels_kernel, a));
  BDET_CUDA(err);
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" size_t bdet_topk_workspace(int64_t total, int n_seg, int k) {
  (void)k;
  return carve_ws(nullptr, total, n_seg < 1 ? 1 : n_seg, false).bytes;
}

extern "C" size_t bdet_score_filter_topk_workspace(int64_t total, int n_seg, int k) {
  (void)k;
  return carve_ws(nullptr, total < 0 ? 0 : total, n_seg < 1 ? 1 : n_seg, true).bytes;
}

extern "C" int bdet_topk(const float* scores, const int64_t* seg_start_host, const int64_t* seg_len_host, int n_seg, int k,
                         float* out_vals, int* out_idx, int* out_count, void* workspace, size_t workspace_bytes,
                         bdet_stream_t stream) {
  BDET_REQUIRE(n_seg >= 0 && k >= 0, "negative size");
  if (n_seg == 0) return BDET_OK;
  BDET_REQUIRE(seg_start_host && seg_len_host && out_count, "null argument");
  BDET_REQUIRE(k == 0 || (out_vals && out_idx), "null output");
  if (k > kMaxK) return set_error(BDET_EUNSUPPORTED, "bdet_topk: k > %d", kMaxK);
  if (n_seg > kMaxSeg) return set_error(BDET_EUNSUPPORTED, "bdet_topk: more than %d segments", kMaxSeg);
  static thread_local SegDesc host[kMaxSeg + 1];
  bool vec = true;
  const int64_t total = build_segments(host, seg_start_host, seg_len_host, nullptr, 1, n_seg, &vec, "bdet_topk");
  if (total < 0) return BDET_EINVAL;
  BDET_REQUIRE(total == 0 || scores, "null scores");
  TopkWs w = carve_ws(workspace, total, n_seg, false);
  if (!workspace || workspace_bytes < w.bytes) return set_error(BDET_EWORKSPACE, "bdet_topk: workspace needs %zu bytes", w.bytes);
  BDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "workspace must be 8-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (int rc = upload_segments(w.seg, host, n_seg + 1, st)) return rc;
  SelArgs a{scores, nullptr, nullptr, w.seg, out_vals, out_idx, out_count, k, next_pow2(k < 2 ? 2 : k)};
  int rc = launch_select<true>(a, n_seg, st);
  if (rc) return rc;
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

static int score_filter_topk_impl(const float* logits, const float* ctrness, int C, const int64_t* seg_start_host,
                                  const int64_t* seg_len_host, const int64_t* ctr_start_host, const int* seg_hw_host, int na,
                                  int n_seg, float threshold, int k, int mode, float* out_scores, int* out_idx,
                                  int* out_count, void* workspace, size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(n_seg >= 0 && k >= 0 && C >= 1, "bad size");
  BDET_REQUIRE(mode >= BDET_SCORE_RAW && mode <= BDET_SCORE_FCOS, "unknown mode");
  BDET_REQUIRE(mode != BDET_SCORE_FCOS || ctrness, "FCOS mode needs ctrness");
  if (n_seg == 0) return BDET_OK;
  BDET_REQUIRE(seg_start_host && seg_len_host && out_count, "null argument");
  BDET_REQUIRE(k == 0 || (out_scores && out_idx), "null output");
  if (k > kMaxK) return set_error(BDET_EUNSUPPORTED, "bdet_score_filter_topk: k > %d", kMaxK);
  if (n_seg > kMaxSeg) return set_error(BDET_EUNSUPPORTED, "bdet_score_filter_topk: more than %d segments", kMaxSeg);
  static thread_local SegDesc host[kMaxSeg + 1];
  bool vec = aligned16(logits);
  const int64_t total = build_segments(host, seg_start_host, seg_len_host, ctr_start_host, C, n_seg, &vec, "bdet_score_filter_topk",
                                       seg_hw_host);
  if (total < 0) return BDET_EINVAL;
  BDET_REQUIRE(total == 0 || logits, "null logits");
  if (seg_hw_host)
    for (int s = 0; s < n_seg; ++s)
      BDET_REQUIRE(seg_hw_host[s] >= 0 && seg_len_host[s] == (int64_t)seg_hw_host[s] * na * C && ctr_start_host,
                   "NCHW segment: len must be H*W*A*C and ctr_start given");
  TopkWs w = carve_ws(workspace, total, n_seg, true);
  if (!workspace || workspace_bytes < w.bytes)
    return set_error(BDET_EWORKSPACE, "bdet_score_filter_topk: workspace needs %zu bytes", w.bytes);
  BDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "workspace must be 8-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int tiles = host[n_seg].tile_start;
  if (int rc = upload_segments(w.seg, host, n_seg + 1, st)) return rc;
  BDET_CUDA(cudaMemsetAsync(w.cand_count, 0, (size_t)n_seg * 4, st));
  FilterArgs f;
  f.logits = logits;
  f.ctr = ctrness;
  f.seg = w.seg;
  f.keys = w.keys;
  f.cand_count = w.cand_count;
  f.scores_out = nullptr;
  f.n_seg = n_seg;
  f.C = C;
  f.mode = mode;
  f.na = seg_hw_host ? na : 0;
  f.thr = threshold;
  f.seg_thr = nullptr;
  f.seg_redo = w.seg_redo;
  f.k = k;
  f.redo_pass = 0;
  bool sampled = false;
  // The sampled bound pays when candidates outnumber k by far: by construction in FCOS mode (sqrt(sigmoid * sigmoid) lifts
  // ~30 % of all logits over 0.05), and for very long segments in any mode.  It costs two short extra launches, so plain
  // sigmoid heads (~1 % candidates, the single-image RetinaNet case) skip it.  The result is exact either way.
  if (mode != BDET_SCORE_RAW && k > 0)
    for (int s = 0; s < n_seg; ++s)
      sampled = sampled || (seg_len_host[s] >= kFSampleMin && (mode == BDET_SCORE_FCOS || seg_len_host[s] >= (1 << 23)));
  // raw-logit pre-filter: sigmoid(x) > q  needs  x > logit(q); keep a safety margin for fp32 rounding.
  f.pre = -std::numeric_limits<float>::infinity();
  if (mode != BDET_SCORE_RAW) {
    double q = mode == BDET_SCORE_FCOS ? (double)threshold * (double)threshold : (double)threshold;
    if (threshold > 0.f && q < 1.0) {
      double l = log(q / (1.0 - q));
      f.pre = (float)(l - 1e-3 * (fabs(l) > 1.0 ? fabs(l) : 1.0));
    } else if (q >= 1.0) {
      f.pre = std::numeric_limits<float>::infinity();  // nothing can pass
    }
  }
  if (tiles > 0) {
    const int grid = min(ceil_div(tiles, kFiltThreads / 32), sm_count() * 4);  // persistent warps, 4 CTAs / SM
    if (sampled) {
      f.seg_thr = w.seg_thr;
      BDET_KERNEL("filter_sample_kernel", st, filter_sample_kernel<<<n_seg, kFSThreads, 0, st>>>(f));
    }
    for (int pass = 0; pass < (sampled ? 2 : 1); ++pass) {
      f.redo_pass = pass;
      if (f.na > 0) {  // NCHW head outputs: survivors are re-indexed (two integer divisions each)
        if (vec)
          BDET_KERNEL("score_filter_kernel", st, score_filter_kernel<true, true><<<grid, kFiltThreads, 0, st>>>(f, tiles));
        else
          BDET_KERNEL("score_filter_kernel", st, score_filter_kernel<false, true><<<grid, kFiltThreads, 0, st>>>(f, tiles));
      } else if (vec) {
        BDET_KERNEL("score_filter_kernel", st, score_filter_kernel<true, false><<<grid, kFiltThreads, 0, st>>>(f, tiles));
      } else {
        BDET_KERNEL("score_filter_kernel", st, score_filter_kernel<false, false><<<grid, kFiltThreads, 0, st>>>(f, tiles));
      }
    }
  }
  SelArgs a{nullptr, w.keys, w.cand_count, w.seg, out_scores, out_idx, out_count, k, next_pow2(k < 2 ? 2 : k)};
  int rc = launch_select<false>(a, n_seg, st);
  if (rc) return rc;
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}

extern "C" int bdet_score_filter_topk(const float* logits, const float* ctrness, int C, const int64_t* seg_start_host,
                                      const int64_t* seg_len_host, const int64_t* ctr_start_host, int n_seg, float threshold,
                                      int k, int mode, float* out_scores, int* out_idx, int* out_count, void* workspace,
                                      size_t workspace_bytes, bdet_stream_t stream) {
  return score_filter_topk_impl(logits, ctrness, C, seg_start_host, seg_len_host, ctr_start_host, nullptr, 0, n_seg, threshold, k,
                                mode, out_scores, out_idx, out_count, workspace, workspace_bytes, stream);
}

extern "C" int bdet_score_filter_topk_nchw(const float* logits, const float* ctrness, int C, int num_anchors,
                                           const int64_t* seg_start_host, const int* seg_hw_host,
                                           const int64_t* ctr_start_host, int n_seg, float threshold, int k, int mode,
                                           float* out_scores, int* out_idx, int* out_count, void* workspace,
                                           size_t workspace_bytes, bdet_stream_t stream) {
  BDET_REQUIRE(num_anchors >= 1 && C >= 1 && n_seg >= 0 && n_seg <= kMaxSeg, "bad size");
  BDET_REQUIRE(n_seg == 0 || (seg_hw_host && seg_start_host), "null argument");
  static thread_local int64_t len[kMaxSeg], cst[kMaxSeg];
  for (int s = 0; s < n_seg; ++s) {
    BDET_REQUIRE(seg_hw_host[s] >= 0, "negative H*W");
    len[s] = (int64_t)seg_hw_host[s] * num_anchors * C;
    cst[s] = ctr_start_host ? ctr_start_host[s] : 0;
  }
  return score_filter_topk_impl(logits, ctrness, C, seg_start_host, len, cst, seg_hw_host, num_anchors, n_seg, threshold, k, mode,
                                out_scores, out_idx, out_count, workspace, workspace_bytes, stream);
}

extern "C" int bdet_scores(const float* logits, const float* ctrness, int C, int64_t n, int mode, float* out,
                           bdet_stream_t stream) {
  BDET_REQUIRE(n >= 0 && C >= 1, "bad size");
  BDET_REQUIRE(mode >= BDET_SCORE_RAW && mode <= BDET_SCORE_FCOS, "unknown mode");
  BDET_REQUIRE(mode != BDET_SCORE_FCOS || ctrness, "FCOS mode needs ctrness");
  if (n == 0) return BDET_OK;
  BDET_REQUIRE(logits && out, "null argument");
  BDET_KERNEL("scores_kernel", as_stream(stream), scores_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(logits, ctrness, C, n, mode, out));
  BDET_LAUNCH_CHECK();
  return BDET_OK;
}
