// Library-level entry points of libkpreg_b200.so (see include/kpreg_b200.h).
#include <atomic>
#include <mutex>
#include <vector>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace kpreg {
namespace {
thread_local char g_last_error[512] = "";
std::atomic<unsigned long long> g_launches{0};
}  // namespace

void set_last_error(const char* what, cudaError_t err) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", what, cudaGetErrorName(err), cudaGetErrorString(err));
  cudaGetLastError();  // clear the sticky-less error state so later calls report their own failures
}
void count_launches(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {
struct ProfRecord { cudaEvent_t begin, end; int family; };
std::mutex g_prof_mutex;
std::vector<ProfRecord> g_prof_pool;  // events are created once and reused
size_t g_prof_used = 0;
bool g_prof_on = false;
}  // namespace

ProfScope::ProfScope(int family, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (g_prof_used == g_prof_pool.size()) {
    ProfRecord r;
    if (cudaEventCreate(&r.begin) != cudaSuccess || cudaEventCreate(&r.end) != cudaSuccess) return;
    g_prof_pool.push_back(r);
  }
  slot = (int)g_prof_used++;
  g_prof_pool[slot].family = family;
  cudaEventRecord(g_prof_pool[slot].begin, stream);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof_pool[slot].end, stream);
}
}  // namespace kpreg

extern "C" int kpreg_profile(int enable) {
  std::lock_guard<std::mutex> lock(kpreg::g_prof_mutex);
  if (enable) kpreg::g_prof_used = 0;
  kpreg::g_prof_on = enable != 0;
  return KPREG_OK;
}

// Pre-create the CUDA events of `n_records` timed scopes, so that a measured region creates none.
extern "C" int kpreg_profile_reserve(int n_records) {
  std::lock_guard<std::mutex> lock(kpreg::g_prof_mutex);
  while ((int)kpreg::g_prof_pool.size() < n_records) {
    kpreg::ProfRecord r;
    KP_CUDA_TRY(cudaEventCreate(&r.begin));
    KP_CUDA_TRY(cudaEventCreate(&r.end));
    r.family = -1;
    kpreg::g_prof_pool.push_back(r);
  }
  return KPREG_OK;
}

extern "C" int kpreg_profile_read(double* ms, unsigned long long* launches) {
  if (!ms || !launches) return KPREG_E_INVALID;
  std::lock_guard<std::mutex> lock(kpreg::g_prof_mutex);
  for (int f = 0; f < KPREG_N_FAMILIES; ++f) { ms[f] = 0.0; launches[f] = 0; }
  for (size_t i = 0; i < kpreg::g_prof_used; ++i) {
    const kpreg::ProfRecord& r = kpreg::g_prof_pool[i];
    float t = 0.f;
    KP_CUDA_TRY(cudaEventSynchronize(r.end));
    KP_CUDA_TRY(cudaEventElapsedTime(&t, r.begin, r.end));
    if (r.family >= 0 && r.family < KPREG_N_FAMILIES) { ms[r.family] += (double)t; launches[r.family] += 1; }
  }
  return KPREG_OK;
}

extern "C" int kpreg_version(void) { return 100; }
extern "C" const char* kpreg_last_error(void) { return kpreg::g_last_error; }
extern "C" unsigned long long kpreg_launch_count(void) { return kpreg::g_launches.load(std::memory_order_relaxed); }
