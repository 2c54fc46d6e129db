// Library-level entry points of libkpreg_b200.so (see include/kpreg_b200.h).
#include <atomic>
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
}  // namespace kpreg

extern "C" int kpreg_version(void) { return 100; }
extern "C" const char* kpreg_last_error(void) { return kpreg::g_last_error; }
extern "C" unsigned long long kpreg_launch_count(void) { return kpreg::g_launches.load(std::memory_order_relaxed); }
