// Error reporting, device queries and small host-side helpers shared by all entry points.
#include <stdarg.h>
#include <string.h>

#include <limits>
#include <vector>

#include "common.cuh"

namespace bdet {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

int make_match_cfg(MatchCfg* cfg, const float* thresholds_host, const int* labels_host, int n_labels) {
  if (n_labels < 1 || n_labels > BDET_MAX_MATCH_LABELS)
    return set_error(BDET_EINVAL, "matcher: n_labels must be in [1, %d]", BDET_MAX_MATCH_LABELS);
  if (!labels_host || (n_labels > 1 && !thresholds_host))
    return set_error(BDET_EINVAL, "matcher: thresholds and labels are not matched");
  // layers/common/matcher.py:23: thresholds must be non-decreasing
  for (int k = 0; k + 2 < n_labels; ++k)
    if (!(thresholds_host[k] <= thresholds_host[k + 1]))
      return set_error(BDET_EINVAL, "matcher: thresholds must be sorted (low <= high)");
  memset(cfg, 0, sizeof(*cfg));
  cfg->n = n_labels;
  cfg->thr[0] = -std::numeric_limits<float>::infinity();
  for (int k = 0; k + 1 < n_labels; ++k) cfg->thr[k + 1] = thresholds_host[k];
  cfg->thr[n_labels] = std::numeric_limits<float>::infinity();
  for (int k = 0; k < n_labels; ++k) cfg->lab[k] = labels_host[k];
  return BDET_OK;
}

struct ProfPair {
  cudaEvent_t a, b;
  const char* name;
};
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfPair>* g_prof = nullptr;
static thread_local std::vector<cudaEvent_t>* g_pool = nullptr;  // events are recycled: creating one costs microseconds
static thread_local char g_only[64] = "";                        // bracket only this kernel (others are just counted)
static thread_local int g_untimed = 0;
static thread_local bool g_open = false;

static bool pool_get(cudaEvent_t* e) {
  if (!g_pool->empty()) {
    *e = g_pool->back();
    g_pool->pop_back();
    return true;
  }
  return cudaEventCreate(e) == cudaSuccess;
}

void prof_pre(cudaStream_t st, const char* name) {
  if (!g_prof_on) return;
  g_open = false;
  if (g_only[0] && strcmp(g_only, name) != 0) {
    ++g_untimed;
    return;
  }
  ProfPair p;
  p.name = name;
  if (!pool_get(&p.a)) return;
  if (!pool_get(&p.b)) {
    g_pool->push_back(p.a);
    return;
  }
  cudaEventRecord(p.a, st);
  g_prof->push_back(p);
  g_open = true;
}

void prof_post(cudaStream_t st) {
  if (!g_prof_on || !g_open) return;
  cudaEventRecord(g_prof->back().b, st);
  g_open = false;
}

}  // namespace bdet

extern "C" {

int bdet_profile_begin(void) {
  if (!bdet::g_prof) bdet::g_prof = new std::vector<bdet::ProfPair>();
  if (!bdet::g_pool) bdet::g_pool = new std::vector<cudaEvent_t>();
  bdet_profile_end();
  bdet::g_prof_on = true;
  return BDET_OK;
}

int bdet_profile_select(const char* name) {
  if (name && strlen(name) >= sizeof(bdet::g_only)) return bdet::set_error(BDET_EINVAL, "bdet_profile_select: name too long");
  strcpy(bdet::g_only, name ? name : "");
  return BDET_OK;
}

int bdet_profile_collect(const char* name, float* total_ms_host, int* launches_host) {
  float total = 0.f;
  int n = 0;
  if (bdet::g_prof) {
    for (auto& p : *bdet::g_prof) {
      if (name && strcmp(name, p.name) != 0) continue;
      BDET_CUDA(cudaEventSynchronize(p.b));
      float ms = 0.f;
      BDET_CUDA(cudaEventElapsedTime(&ms, p.a, p.b));
      total += ms;
      ++n;
    }
  }
  if (!name) n += bdet::g_untimed;  // launches that were counted but not bracketed (bdet_profile_select)
  if (total_ms_host) *total_ms_host = total;
  if (launches_host) *launches_host = n;
  return BDET_OK;
}

int bdet_profile_report(char* buf_host, size_t buf_bytes) {
  if (!buf_host || buf_bytes < 2) return bdet::set_error(BDET_EINVAL, "bdet_profile_report: buffer too small");
  buf_host[0] = 0;
  if (!bdet::g_prof) return BDET_OK;
  std::vector<const char*> names;
  for (auto& p : *bdet::g_prof) {
    bool seen = false;
    for (auto n : names) seen = seen || strcmp(n, p.name) == 0;
    if (!seen) names.push_back(p.name);
  }
  size_t used = 0;
  for (auto n : names) {
    float ms = 0.f;
    int cnt = 0;
    int rc = bdet_profile_collect(n, &ms, &cnt);
    if (rc) return rc;
    int w = snprintf(buf_host + used, buf_bytes - used, "%s %.6f %d\n", n, ms, cnt);
    if (w < 0 || (size_t)w >= buf_bytes - used) return bdet::set_error(BDET_EINVAL, "bdet_profile_report: buffer too small");
    used += (size_t)w;
  }
  return BDET_OK;
}

int bdet_profile_end(void) {
  bdet::g_prof_on = false;
  bdet::g_untimed = 0;
  bdet::g_open = false;
  if (bdet::g_prof) {
    for (auto& p : *bdet::g_prof) {
      bdet::g_pool->push_back(p.a);
      bdet::g_pool->push_back(p.b);
    }
    bdet::g_prof->clear();
  }
  return BDET_OK;
}

int bdet_abi_version(void) { return BDET_ABI_VERSION; }

const char* bdet_last_error(void) { return bdet::g_err; }

int bdet_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
  int dev = 0;
  BDET_CUDA(cudaGetDevice(&dev));
  int n = 0, maj = 0, min = 0;
  BDET_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  BDET_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  BDET_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count_host) *sm_count_host = n;
  if (cc_major_host) *cc_major_host = maj;
  if (cc_minor_host) *cc_minor_host = min;
  return BDET_OK;
}

}  // extern "C"
