/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of BaseDet's box-op hot path (same op order as
 * oracle/ref_ops.py, which is pinned bit-exact to the reference's own source through tests/golden).
 * Compiled with -ffp-contract=off: one fp32 rounding per operation, no FMA.
 * Used (a) to check the CUDA path at sizes where numpy is too slow, (b) as the CPU baseline of bench.py
 * (MegEngine CPU cannot be installed offline).  Citations are into megvii-research/basedet.
 *
 * The matrix is MATERIALISED and every reference op is its own pass, as the un-fused MegEngine CPU graph does. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

static inline float emax(float x, float y) { return x > y ? x : y; } /* MegDNN MAX (ASSUMED-1) */
static inline float emin(float x, float y) { return x < y ? x : y; }

/* ---- minimal pthread parallel-for (libgomp is not in the image) ---------------------------------------- */
static int g_threads = 0;
int oracle_num_threads(void) {
  if (g_threads <= 0) {
    const char* e = getenv("ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    g_threads = (int)(n < 1 ? 1 : (n > 256 ? 256 : n));
  }
  return g_threads;
}
void oracle_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

typedef void (*range_fn)(int lo, int hi, void* ctx);
typedef struct { range_fn fn; void* ctx; int lo, hi; } job_t;
static void* job_main(void* p) {
  job_t* j = (job_t*)p;
  j->fn(j->lo, j->hi, j->ctx);
  return NULL;
}
static void parallel_for(int n, range_fn fn, void* ctx) {
  int nt = oracle_num_threads();
  if (nt > n) nt = n;
  if (nt <= 1) { if (n > 0) fn(0, n, ctx); return; }
  pthread_t th[256];
  job_t jobs[256];
  int chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].lo = t * chunk; jobs[t].hi = (t + 1) * chunk > n ? n : (t + 1) * chunk;
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
  }
  for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}

/* structures/op_patch.py:33-78: b1 (N, ld1), b2 (M, ld2) -> out (N, M) */
typedef struct { const float* b1; int ld1; const float* b2; int ld2, M; float* out; } iou_ctx;
static void iou_rows(int lo, int hi, void* p) {
  iou_ctx* c = (iou_ctx*)p;
  const float* b1 = c->b1; const float* b2 = c->b2; float* out = c->out;
  const int ld1 = c->ld1, ld2 = c->ld2, M = c->M;
  for (int i = lo; i < hi; ++i) {
    const float* a = b1 + (size_t)i * ld1;
    float area1 = (a[2] - a[0]) * (a[3] - a[1]);
    float* o = out + (size_t)i * M;
    for (int j = 0; j < M; ++j) {
      const float* b = b2 + (size_t)j * ld2;
      float iw = emin(a[2], b[2]) - emax(a[0], b[0]);
      float ih = emin(a[3], b[3]) - emax(a[1], b[1]);
      iw = emax(iw, 0.f);
      ih = emax(ih, 0.f);
      float inter = iw * ih;
      float area2 = (b[2] - b[0]) * (b[3] - b[1]);
      float uni = area1 + area2;
      uni = uni - inter;
      float iou = inter / uni;
      o[j] = emax(iou, 0.f);
    }
  }
}
void oracle_box_iou(const float* b1, int ld1, int N, const float* b2, int ld2, int M, float* out) {
  iou_ctx c = {b1, ld1, b2, ld2, M, out};
  parallel_for(N, iou_rows, &c);
}

/* layers/common/matcher.py:31-51; thr has n_labels+1 entries (-inf ... +inf) */
typedef struct { const float* m; int G, A; const float* thr; const int* labels; int n_labels; int* match_idx; int* out_labels; float* rowmax; } match_ctx;
static void match_cols(int lo, int hi, void* p) {
  match_ctx* c = (match_ctx*)p;
  const float* m = c->m; const int G = c->G, A = c->A, n_labels = c->n_labels;
  const float* thr = c->thr; const int* labels = c->labels; int* match_idx = c->match_idx; int* out_labels = c->out_labels;
  for (int a = lo; a < hi; ++a) {
    float best = m[a];
    int bi = 0;
    for (int g = 1; g < G; ++g) {
      float v = m[(size_t)g * A + a];
      if (v > best) { best = v; bi = g; } /* first index on ties (ASSUMED-2) */
    }
    int lab = -1;
    for (int k = 0; k < n_labels; ++k)
      if (best >= thr[k] && best < thr[k + 1]) lab = labels[k];
    match_idx[a] = bi;
    out_labels[a] = lab;
  }
}
static void match_rowmax(int lo, int hi, void* p) {
  match_ctx* c = (match_ctx*)p;
  for (int g = lo; g < hi; ++g) {
    const float* row = c->m + (size_t)g * c->A;
    float rm = row[0];
    for (int a = 1; a < c->A; ++a) rm = row[a] > rm ? row[a] : rm;
    c->rowmax[g] = rm;
  }
}
static void match_lq_cols(int lo, int hi, void* p) {
  match_ctx* c = (match_ctx*)p;
  for (int g = 0; g < c->G; ++g) {
    const float* row = c->m + (size_t)g * c->A;
    const float rm = c->rowmax[g];
    for (int a = lo; a < hi; ++a)
      if (row[a] == rm) c->out_labels[a] = 1;
  }
}
void oracle_matcher(const float* m, int G, int A, const float* thr, const int* labels, int n_labels, int allow_lq,
                    int* match_idx, int* out_labels) {
  match_ctx c = {m, G, A, thr, labels, n_labels, match_idx, out_labels, NULL};
  parallel_for(A, match_cols, &c);
  if (allow_lq) {
    c.rowmax = (float*)malloc(sizeof(float) * (size_t)(G > 0 ? G : 1));
    parallel_for(G, match_rowmax, &c);
    parallel_for(A, match_lq_cols, &c);
    free(c.rowmax);
  }
}

/* structures/boxcoder.py:44-73; gather != NULL: gt row = gt[gather[i]] (ld floats per row) */
typedef struct { const float* bbox; const float* gt; int gt_ld; const int* gather; const float* mean; const float* std; float* out; } enc_ctx;
static void enc_rows(int lo, int hi, void* p) {
  enc_ctx* c = (enc_ctx*)p;
  const float* bbox = c->bbox; const float* gt = c->gt; const int gt_ld = c->gt_ld; const int* gather = c->gather;
  const float* mean = c->mean; const float* std = c->std; float* out = c->out;
  for (int i = lo; i < hi; ++i) {
    const float* b = bbox + (size_t)i * 4;
    const float* g = gt + (size_t)(gather ? gather[i] : i) * gt_ld;
    float bw = b[2] - b[0], bh = b[3] - b[1];
    float bcx = b[0] + 0.5f * bw, bcy = b[1] + 0.5f * bh;
    float gw = g[2] - g[0], gh = g[3] - g[1];
    float gcx = g[0] + 0.5f * gw, gcy = g[1] + 0.5f * gh;
    float t[4];
    t[0] = (gcx - bcx) / bw;
    t[1] = (gcy - bcy) / bh;
    t[2] = logf(gw / bw);
    t[3] = logf(gh / bh);
    for (int k = 0; k < 4; ++k) out[(size_t)i * 4 + k] = (t[k] - mean[k]) / std[k];
  }
}
void oracle_box_encode(const float* bbox, const float* gt, int gt_ld, const int* gather, int N, const float* mean,
                       const float* std, float* out) {
  enc_ctx c = {bbox, gt, gt_ld, gather, mean, std, out};
  parallel_for(N, enc_rows, &c);
}

/* models/det/retinanet.py:211-232 for one image */
void oracle_retinanet_targets(const float* anchors, int A, const float* gt5, int G, const float* thr, const int* labels,
                              int n_labels, int allow_lq, const float* mean, const float* std, float* iou_scratch,
                              int* match_idx, int* out_labels, float* offsets) {
  oracle_box_iou(gt5, 5, G, anchors, 4, A, iou_scratch);
  oracle_matcher(iou_scratch, G, A, thr, labels, n_labels, allow_lq, match_idx, out_labels);
  for (int a = 0; a < A; ++a)
    if (out_labels[a] == 1) out_labels[a] = (int)gt5[(size_t)match_idx[a] * 5 + 4];
  oracle_box_encode(anchors, gt5, 5, match_idx, A, mean, std, offsets);
}

/* F.vision.nms on boxes already sorted by score (ASSUMED-5): returns #kept, kept[] = sorted positions */
typedef struct { const float* b; const float* area; uint8_t* removed; int i; float thr; } nms_ctx;
static void nms_range(int lo, int hi, void* p) {
  nms_ctx* c = (nms_ctx*)p;
  const float* a = c->b + 4 * (size_t)c->i;
  const int base = c->i + 1;
  for (int jj = lo; jj < hi; ++jj) {
    const int j = base + jj;
    if (c->removed[j]) continue;
    const float* q = c->b + 4 * (size_t)j;
    float w = fmaxf(fminf(a[2], q[2]) - fmaxf(a[0], q[0]), 0.f);
    float h = fmaxf(fminf(a[3], q[3]) - fmaxf(a[1], q[1]), 0.f);
    float inter = w * h;
    if (!(inter > 0.f) && c->thr >= 0.f) continue; /* 0 / x can only exceed a negative threshold */
    float iou = inter / ((c->area[c->i] + c->area[j]) - inter);
    if (iou > c->thr) c->removed[j] = 1;
  }
}
static int nms_sorted_simple(const float* b, int n, float thr, int max_output, int* kept) {
  uint8_t* removed = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);
  float* area = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) area[i] = (b[4 * i + 2] - b[4 * i]) * (b[4 * i + 3] - b[4 * i + 1]);
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    kept[cnt++] = i;
    if (max_output > 0 && cnt >= max_output) break;
    nms_ctx c = {b, area, removed, i, thr};
    const int rest = n - i - 1;
    if (rest > 0) nms_range(0, rest, &c);
  }
  free(removed);
  free(area);
  return cnt;
}

/* Same greedy result for large n, blocked so that the bulk of the pair tests runs on all cores: box j is removed iff a
 * KEPT box i < j has IoU(i, j) > thr.  Per block of kNmsBlock sorted boxes: (1) in parallel, every box of the block is
 * tested against the boxes kept in EARLIER blocks (that list is final); (2) the survivors are resolved sequentially
 * inside the block.  Identical decisions, identical fp32 expression. */
enum { kNmsBlock = 4096 };
typedef struct { const float* b; const float* area; const int* kept; int n_kept; uint8_t* removed; int base; float thr; } nmsb_ctx;
static inline int nms_hit(const float* a, float aa, const float* q, float qa, float thr) {
  /* emax / emin: the same values as fmaxf / fminf on NaN-free boxes, without the libm call */
  float w = emax(emin(a[2], q[2]) - emax(a[0], q[0]), 0.f);
  float h = emax(emin(a[3], q[3]) - emax(a[1], q[1]), 0.f);
  float inter = w * h;
  if (!(inter > 0.f) && thr >= 0.f) return 0;
  return inter / ((aa + qa) - inter) > thr;
}
static void nms_block_pull(int lo, int hi, void* p) {
  nmsb_ctx* c = (nmsb_ctx*)p;
  for (int jj = lo; jj < hi; ++jj) {
    const int j = c->base + jj;
    const float* q = c->b + 4 * (size_t)j;
    const float qa = c->area[j];
    for (int t = 0; t < c->n_kept; ++t) {
      const int i = c->kept[t];
      if (nms_hit(c->b + 4 * (size_t)i, c->area[i], q, qa, c->thr)) { c->removed[j] = 1; break; }
    }
  }
}
int oracle_nms_sorted(const float* b, int n, float thr, int max_output, int* kept) {
  if (n <= 2 * kNmsBlock) return nms_sorted_simple(b, n, thr, max_output, kept);
  uint8_t* removed = (uint8_t*)calloc((size_t)n, 1);
  float* area = (float*)malloc(sizeof(float) * (size_t)n);
  for (int i = 0; i < n; ++i) area[i] = (b[4 * i + 2] - b[4 * i]) * (b[4 * i + 3] - b[4 * i + 1]);
  int cnt = 0, done = 0;
  for (int base = 0; base < n && !done; base += kNmsBlock) {
    const int m = n - base < kNmsBlock ? n - base : kNmsBlock;
    nmsb_ctx c = {b, area, kept, cnt, removed, base, thr};
    if (cnt > 0) parallel_for(m, nms_block_pull, &c);
    const int first = cnt;
    for (int j = base; j < base + m; ++j) {
      if (removed[j]) continue;
      int hit = 0;
      for (int t = first; t < cnt && !hit; ++t) hit = nms_hit(b + 4 * (size_t)kept[t], area[kept[t]], b + 4 * (size_t)j, area[j], thr);
      if (hit) continue;
      kept[cnt++] = j;
      if (max_output > 0 && cnt >= max_output) { done = 1; break; }
    }
  }
  free(removed);
  free(area);
  return cnt;
}

/* F.nn.roi_align(mode=average, aligned) (ASSUMED-6): feat (B,C,H,W), rois (K,5) -> out (K,C,PH,PW).
 * ROIs are independent: parallel over k. */
typedef struct {
  const float* feat; const float* rois; const float* dout; float* out; double* grad;
  int C, H, W, K, PH, PW, SH, SW; float scale, offset;
} roi_ctx;
static void roi_fwd_range(int lo, int hi, void* p) {
  roi_ctx* q = (roi_ctx*)p;
  const int C = q->C, H = q->H, W = q->W, PH = q->PH, PW = q->PW, SH = q->SH, SW = q->SW;
  const float scale = q->scale, offset = q->offset;
  for (int k = lo; k < hi; ++k) {
    const float* r = q->rois + 5 * (size_t)k;
    const float* fm = q->feat + (size_t)((int)r[0]) * C * H * W;
    float sw_ = r[1] * scale - offset, sh_ = r[2] * scale - offset;
    float ew_ = r[3] * scale - offset, eh_ = r[4] * scale - offset;
    float rw = emax(ew_ - sw_, 0.f), rh = emax(eh_ - sh_, 0.f);
    float bh = rh / (float)PH, bw = rw / (float)PW;
    for (int c = 0; c < C; ++c) {
      const float* f = fm + (size_t)c * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.f;
          for (int iy = 0; iy < SH; ++iy)
            for (int ix = 0; ix < SW; ++ix) {
              float hc = sh_ + bh * ((float)ph + ((float)iy + 0.5f) / (float)SH);
              float wc = sw_ + bw * ((float)pw + ((float)ix + 0.5f) / (float)SW);
              float fh = floorf(hc), fw = floorf(wc);
              int h0 = (int)fh, w0 = (int)fw, h1 = h0 + 1, w1 = w0 + 1;
              float lh = hc - fh, lw = wc - fw;
              float tl = (h0 >= 0 && h0 < H && w0 >= 0 && w0 < W) ? f[h0 * W + w0] : 0.f;
              float tr = (h0 >= 0 && h0 < H && w1 >= 0 && w1 < W) ? f[h0 * W + w1] : 0.f;
              float bl = (h1 >= 0 && h1 < H && w0 >= 0 && w0 < W) ? f[h1 * W + w0] : 0.f;
              float br = (h1 >= 0 && h1 < H && w1 >= 0 && w1 < W) ? f[h1 * W + w1] : 0.f;
              float top = tl + (tr - tl) * lw;
              float bot = bl + (br - bl) * lw;
              acc = acc + (top + (bot - top) * lh);
            }
          q->out[(((size_t)k * C + c) * PH + ph) * PW + pw] = acc / (float)(SH * SW);
        }
    }
  }
}
void oracle_roi_align_fwd(const float* feat, int B, int C, int H, int W, const float* rois, int K, int PH, int PW,
                          int SH, int SW, float scale, float offset, float* out) {
  (void)B;
  roi_ctx q = {feat, rois, NULL, out, NULL, C, H, W, K, PH, PW, SH, SW, scale, offset};
  parallel_for(K, roi_fwd_range, &q);
}

/* backward w.r.t. feat, accumulated in double (order-independent reference); grad (B,C,H,W) must be zeroed.
 * Parallel over channels: every thread walks all ROIs for its own channel range, so no two threads share an element. */
static void roi_bwd_range(int lo, int hi, void* p) {
  roi_ctx* q = (roi_ctx*)p;
  const int C = q->C, H = q->H, W = q->W, PH = q->PH, PW = q->PW, SH = q->SH, SW = q->SW;
  const float scale = q->scale, offset = q->offset;
  for (int k = 0; k < q->K; ++k) {
    const float* r = q->rois + 5 * (size_t)k;
    double* gm = q->grad + (size_t)((int)r[0]) * C * H * W;
    float sw_ = r[1] * scale - offset, sh_ = r[2] * scale - offset;
    float ew_ = r[3] * scale - offset, eh_ = r[4] * scale - offset;
    float rw = emax(ew_ - sw_, 0.f), rh = emax(eh_ - sh_, 0.f);
    float bh = rh / (float)PH, bw = rw / (float)PW;
    for (int c = lo; c < hi; ++c) {
      double* g = gm + (size_t)c * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float gv = q->dout[(((size_t)k * C + c) * PH + ph) * PW + pw] / (float)(SH * SW);
          for (int iy = 0; iy < SH; ++iy)
            for (int ix = 0; ix < SW; ++ix) {
              float hc = sh_ + bh * ((float)ph + ((float)iy + 0.5f) / (float)SH);
              float wc = sw_ + bw * ((float)pw + ((float)ix + 0.5f) / (float)SW);
              float fh = floorf(hc), fw = floorf(wc);
              int h0 = (int)fh, w0 = (int)fw, h1 = h0 + 1, w1 = w0 + 1;
              float lh = hc - fh, lw = wc - fw;
              if (h0 >= 0 && h0 < H && w0 >= 0 && w0 < W) g[h0 * W + w0] += (double)(gv * ((1.f - lh) * (1.f - lw)));
              if (h0 >= 0 && h0 < H && w1 >= 0 && w1 < W) g[h0 * W + w1] += (double)(gv * ((1.f - lh) * lw));
              if (h1 >= 0 && h1 < H && w0 >= 0 && w0 < W) g[h1 * W + w0] += (double)(gv * (lh * (1.f - lw)));
              if (h1 >= 0 && h1 < H && w1 >= 0 && w1 < W) g[h1 * W + w1] += (double)(gv * (lh * lw));
            }
        }
    }
  }
}
void oracle_roi_align_bwd(const float* dout, int B, int C, int H, int W, const float* rois, int K, int PH, int PW, int SH,
                          int SW, float scale, float offset, double* grad) {
  (void)B;
  roi_ctx q = {NULL, rois, dout, NULL, grad, C, H, W, K, PH, PW, SH, SW, scale, offset};
  parallel_for(C, roi_bwd_range, &q);
}
