// Library identity and the thread-local error string of the C-ABI.
#include <stdarg.h>

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

struct ProfSlot {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  cudaEvent_t open = nullptr;
};
static bool g_prof = false;
static ProfSlot g_slots[DESIRE_PROF_SLOTS];
static std::mutex g_prof_mu;

void prof_begin(int slot, cudaStream_t st) {
  if (!g_prof || slot < 0 || slot >= DESIRE_PROF_SLOTS) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_slots[slot].open = e;
}
void prof_end(int slot, cudaStream_t st) {
  if (!g_prof || slot < 0 || slot >= DESIRE_PROF_SLOTS) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfSlot& s = g_slots[slot];
  if (!s.open) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  s.ev.emplace_back(s.open, e);
  s.open = nullptr;
}
static void prof_reset() {
  for (auto& s : g_slots) {
    for (auto& p : s.ev) {
      cudaEventDestroy(p.first);
      cudaEventDestroy(p.second);
    }
    s.ev.clear();
    if (s.open) cudaEventDestroy(s.open);
    s.open = nullptr;
  }
}
}  // namespace desire

extern "C" long desire_launch_count(void) { return desire::g_launches.load(); }
extern "C" int desire_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(desire::g_prof_mu);
  desire::prof_reset();
  desire::g_prof = on != 0;
  return DESIRE_OK;
}
extern "C" int desire_prof_read(int slot, long* launches, double* total_ms) {
  DESIRE_CHECK_ARG(slot >= 0 && slot < DESIRE_PROF_SLOTS && launches && total_ms, "desire_prof_read: bad arguments");
  std::lock_guard<std::mutex> lk(desire::g_prof_mu);
  double tot = 0;
  for (auto& p : desire::g_slots[slot].ev) {
    DESIRE_CUDA(cudaEventSynchronize(p.second));
    float ms = 0;
    DESIRE_CUDA(cudaEventElapsedTime(&ms, p.first, p.second));
    tot += ms;
  }
  *launches = (long)desire::g_slots[slot].ev.size();
  *total_ms = tot;
  return DESIRE_OK;
}

extern "C" int desire_version(void) { return DESIRE_ABI_VERSION; }
extern "C" const char* desire_last_error(void) { return desire::g_err; }
