// Library identity, the thread-local error string, the launch counter and the optional per-kernel
// CUDA-event timers of the C-ABI.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace desire {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- visibility of the non-tensor-core / non-fused paths.  Shapes the tcgen05 kernels do not cover run on FP32
// CUDA cores or on the materialising social path: correct, much slower, and easy not to notice.  Every such launch is
// counted per kind (desire_fallback_count) and, with DESIRE_LOG_FALLBACK=1, the first one of each kind is reported.
static std::atomic<long> g_fallbacks[DESIRE_FALLBACK_KINDS];
void note_fallback(int kind, const char* what, int a, int b, int c) {
  if (kind < 0 || kind >= DESIRE_FALLBACK_KINDS) return;
  const long n = g_fallbacks[kind].fetch_add(1, std::memory_order_relaxed);
  static const bool log = [] {
    const char* e = getenv("DESIRE_LOG_FALLBACK");
    return e && e[0] == '1';
  }();
  if (log && n == 0) fprintf(stderr, "libdesire_b200: fallback kind %d: %s (%d, %d, %d)\n", kind, what, a, b, c);
}

// ---- per-kernel timing: ProfScope names the slot, DESIRE_LAUNCH brackets the launch itself with two
// events taken from a pool (no event creation on the hot path), so host-side preparation between
// launches never lands inside a bracket.
static bool g_prof = false;
static std::mutex g_prof_mu;
static std::vector<cudaEvent_t> g_pool;
static size_t g_pool_next = 0;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_slots[DESIRE_PROF_SLOTS];
static thread_local int t_slot = -1;
static thread_local cudaEvent_t t_open = nullptr;

int prof_set_slot(int slot) {
  int old = t_slot;
  t_slot = slot;
  return old;
}
static cudaEvent_t pool_take() {
  if (g_pool_next == g_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    g_pool.push_back(e);
  }
  return g_pool[g_pool_next++];
}
void prof_kernel_begin(cudaStream_t st) {
  if (!g_prof || t_slot < 0 || t_slot >= DESIRE_PROF_SLOTS) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  t_open = pool_take();
  if (t_open) cudaEventRecord(t_open, st);
}
void prof_kernel_end(cudaStream_t st) {
  if (!g_prof || !t_open) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t e = pool_take();
  if (e) {
    cudaEventRecord(e, st);
    g_slots[t_slot].emplace_back(t_open, e);
  }
  t_open = nullptr;
}
}  // namespace desire

extern "C" int desire_version(void) { return DESIRE_ABI_VERSION; }
extern "C" const char* desire_last_error(void) { return desire::g_err; }
extern "C" long desire_launch_count(void) { return desire::g_launches.load(); }
extern "C" long desire_fallback_count(int kind) {
  return (kind >= 0 && kind < DESIRE_FALLBACK_KINDS) ? desire::g_fallbacks[kind].load() : -1;
}

extern "C" int desire_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(desire::g_prof_mu);
  for (auto& s : desire::g_slots) s.clear();
  desire::g_pool_next = 0;           // events are reused; earlier readings are discarded
  desire::g_prof = on != 0;
  return DESIRE_OK;
}

extern "C" int desire_prof_read(int slot, long* launches, double* total_ms) {
  DESIRE_CHECK_ARG(slot >= 0 && slot < DESIRE_PROF_SLOTS && launches && total_ms, "desire_prof_read: bad arguments");
  std::lock_guard<std::mutex> lk(desire::g_prof_mu);
  double tot = 0;
  for (auto& p : desire::g_slots[slot]) {
    DESIRE_CUDA(cudaEventSynchronize(p.second));
    float ms = 0;
    DESIRE_CUDA(cudaEventElapsedTime(&ms, p.first, p.second));
    tot += ms;
  }
  *launches = (long)desire::g_slots[slot].size();
  *total_ms = tot;
  return DESIRE_OK;
}
