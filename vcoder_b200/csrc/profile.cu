// Launch accounting and optional per-kernel-category timing (CUDA events on the launching stream).
// bench.py uses it for `gpu_launches` and for the live roofline numbers of the dominant kernel.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.h"

namespace vcb {
namespace {
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_cat_launches[kNumCat];
bool g_prof = false;
struct Rec { int cat; cudaEvent_t a, b; };
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
std::mutex g_mu;

cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VCB_PDL"); v = e ? atoi(e) : 1; }
  return v != 0;
}

uint64_t launch_count() { return g_launches.load(); }
void launch_count_add(int64_t n) { g_launches.fetch_add(static_cast<uint64_t>(n)); }
bool prof_enabled() { return g_prof; }

void prof_enable(bool on) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
  g_recs.clear();
  g_prof = on;
}

KernelScope::KernelScope(int cat, cudaStream_t st, int n_kernels) : cat_(cat), st_(st), rec_(-1) {
  g_launches.fetch_add(n_kernels);
  g_cat_launches[cat].fetch_add(n_kernels);
  if (g_prof) {
    std::lock_guard<std::mutex> lk(g_mu);
    Rec r{cat, get_event(), get_event()};
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
    rec_ = static_cast<int>(g_recs.size()) - 1;
  }
}
KernelScope::~KernelScope() {
  if (rec_ >= 0) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (rec_ < static_cast<int>(g_recs.size())) cudaEventRecord(g_recs[rec_].b, st_);
  }
}

// ---- device-side timeline -----------------------------------------------------------------------------------
namespace {
struct TraceLaunch { int tag, n, base, aux; };
unsigned long long* g_trace_dev = nullptr;
int g_trace_cap = 0, g_trace_used = 0, g_trace_epoch = 0;
std::vector<TraceLaunch> g_trace_launches;
}  // namespace

int trace_epoch() { return g_trace_epoch; }

int trace_begin(int max_ctas) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_trace_dev) { cudaFree(g_trace_dev); g_trace_dev = nullptr; }
  VCB_CHECK(max_ctas > 0, kErrInvalid, "trace_begin: max_ctas must be positive");
  VCB_CUDA(cudaMalloc(reinterpret_cast<void**>(&g_trace_dev), static_cast<size_t>(max_ctas) * 24));
  VCB_CUDA(cudaMemset(g_trace_dev, 0, static_cast<size_t>(max_ctas) * 24));
  g_trace_cap = max_ctas;
  g_trace_used = 0;
  g_trace_launches.clear();
  ++g_trace_epoch;
  return kOk;
}

unsigned long long* trace_slots(int tag, int n_ctas, int aux) {
  if (g_trace_dev == nullptr) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_trace_dev == nullptr || g_trace_used + n_ctas > g_trace_cap) return nullptr;
  g_trace_launches.push_back({tag, n_ctas, g_trace_used, aux});
  unsigned long long* p = g_trace_dev + static_cast<size_t>(g_trace_used) * 3;
  g_trace_used += n_ctas;
  return p;
}

int trace_read(int32_t* launches_out, int max_launches, int32_t* n_launches, unsigned long long* stamps_out, long cap) {
  VCB_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_mu);
  VCB_CHECK(g_trace_dev != nullptr, kErrState, "trace_read: no trace is open");
  const int n = std::min<int>(static_cast<int>(g_trace_launches.size()), max_launches);
  for (int i = 0; i < n; ++i) {
    launches_out[4 * i] = g_trace_launches[i].tag; launches_out[4 * i + 1] = g_trace_launches[i].n;
    launches_out[4 * i + 2] = g_trace_launches[i].base; launches_out[4 * i + 3] = g_trace_launches[i].aux;
  }
  *n_launches = n;
  const long words = std::min<long>(cap, static_cast<long>(g_trace_used) * 3);
  VCB_CUDA(cudaMemcpy(stamps_out, g_trace_dev, static_cast<size_t>(words) * 8, cudaMemcpyDeviceToHost));
  cudaFree(g_trace_dev);
  g_trace_dev = nullptr;
  ++g_trace_epoch;
  return kOk;
}

// ms_by_cat / n_by_cat: arrays of kNumCat. Synchronises the device.
int prof_read(float* ms_by_cat, uint64_t* n_by_cat) {
  VCB_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_mu);
  for (int c = 0; c < kNumCat; ++c) { ms_by_cat[c] = 0.f; n_by_cat[c] = 0; }
  for (auto& r : g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { ms_by_cat[r.cat] += ms; n_by_cat[r.cat] += 1; }
  }
  return kOk;
}

}  // namespace vcb
