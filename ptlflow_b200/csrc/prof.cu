// Launch accounting and optional per-kernel-class CUDA-event timing.
// bench.py uses this to report (a) how many of this library's kernels ran in the timed region and
// (b) the live average duration of each kernel class for the roofline numbers -- events are recorded
// on the launching stream around each launch only while profiling is enabled (never during the
// timed headline run, never under graph capture).
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

#include <stdlib.h>

namespace pfb {

int pdl_enabled() {
  static const int v = getenv("PFB_PDL") ? atoi(getenv("PFB_PDL")) : 1;
  return v;
}


static std::atomic<unsigned long long> g_launches[KC_COUNT];
static std::atomic<int> g_prof_on{0};
struct Span { int kc; cudaEvent_t a, b; };
static std::mutex g_mu;
static std::vector<Span> g_spans;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(int kc, cudaStream_t s) : kc_(kc), s_(s), a_(nullptr) {
  g_launches[kc].fetch_add(1, std::memory_order_relaxed);
  if (g_prof_on.load(std::memory_order_relaxed)) {
    std::lock_guard<std::mutex> lk(g_mu);
    a_ = get_event();
    cudaEventRecord(reinterpret_cast<cudaEvent_t>(a_), s_);
  }
}

ProfScope::~ProfScope() {
  if (a_) {
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEvent_t b = get_event();
    cudaEventRecord(b, s_);
    g_spans.push_back({kc_, reinterpret_cast<cudaEvent_t>(a_), b});
  }
}

}  // namespace pfb

using namespace pfb;

extern "C" PFB_API unsigned long long pfb_launch_count(int kernel_class) {
  if (kernel_class >= 0 && kernel_class < KC_COUNT) return g_launches[kernel_class].load();
  unsigned long long t = 0;
  for (int i = 0; i < KC_COUNT; ++i) t += g_launches[i].load();
  return t;
}

extern "C" PFB_API int pfb_profile_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return PFB_OK;
}

// Synchronises the device, sums the recorded spans per class into ms[KC_COUNT] / n[KC_COUNT] and clears them.
extern "C" PFB_API int pfb_profile_collect(double* ms, unsigned long long* n, int len) {
  PFB_CHECK_ARG(ms && n && len >= KC_COUNT, "profile_collect: need arrays of %d entries", (int)KC_COUNT);
  PFB_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < len; ++i) { ms[i] = 0.0; n[i] = 0; }
  for (const Span& sp : g_spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) { ms[sp.kc] += t; n[sp.kc] += 1; }
    g_pool.push_back(sp.a);
    g_pool.push_back(sp.b);
  }
  g_spans.clear();
  return PFB_OK;
}
