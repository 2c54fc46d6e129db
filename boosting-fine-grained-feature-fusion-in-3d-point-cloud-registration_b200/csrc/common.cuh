// Shared helpers for the kpreg_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kpreg_b200.h"

namespace kpreg {

// ---- host-side bookkeeping ---------------------------------------------------------------------
void set_last_error(const char* what, cudaError_t err);
void count_launches(unsigned long long n);

#define KP_CUDA_TRY(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) {                                \
      ::kpreg::set_last_error(#expr, _e);                   \
      return KPREG_E_CUDA;                                  \
    }                                                       \
  } while (0)

// Check the launch that was just issued and count it.
#define KP_LAUNCH_CHECK()                                   \
  do {                                                      \
    ::kpreg::count_launches(1);                             \
    cudaError_t _e = cudaPeekAtLastError();                 \
    if (_e != cudaSuccess) {                                \
      ::kpreg::set_last_error("kernel launch", _e);         \
      return KPREG_E_CUDA;                                  \
    }                                                       \
  } while (0)

// Times everything enqueued on `stream` during its lifetime when profiling is on (capi.cu).
struct ProfScope {
  int slot;
  cudaStream_t stream;
  ProfScope(int family, cudaStream_t s);
  ~ProfScope();
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline int bits_for(uint64_t n) {  // bits needed to represent values < n
  int b = 0;
  while (b < 63 && (1ull << b) < n) ++b;
  return b;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: call `configure` once per device ordinal and
// kernel (thread-safe; one atomic load on the hot path).  A process that drives several GPUs therefore configures each.
struct PerDeviceOnce {
  unsigned long long done[2] = {0ull, 0ull};  // one bit per device ordinal (0..127)
  template <typename F>
  int run(F&& configure) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return configure();
    unsigned long long* word = &done[dev >> 6];
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(word, __ATOMIC_ACQUIRE) & bit) return KPREG_OK;
    const int rc = configure();
    if (rc == KPREG_OK) __atomic_fetch_or(word, bit, __ATOMIC_RELEASE);
    return rc;
  }
};

// Bump allocator over a caller-provided workspace (256-byte aligned carve-outs).
struct Carver {
  char* base;
  size_t used;
  explicit Carver(void* p) : base(static_cast<char*>(p)), used(0) {}
  template <typename T>
  T* take(size_t count) {
    used = align_up(used, 256);
    T* p = base ? reinterpret_cast<T*>(base + used) : nullptr;
    used += count * sizeof(T);
    return p;
  }
};

constexpr int kNumSMs = 148;  // B200

// ---- device helpers ----------------------------------------------------------------------------
#ifdef __CUDACC__
// Monotone float <-> uint mapping so atomicMin/atomicMax order floats.
__device__ __forceinline__ unsigned int float_to_ordered(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Index of the cloud that owns stacked row i: largest c with off[c] <= i (off has n_clouds+1 entries).
__device__ __forceinline__ int cloud_of(const int64_t* __restrict__ off, int n_clouds, int64_t i) {
  int lo = 0, hi = n_clouds;  // invariant: off[lo] <= i < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ unsigned int lane_id() { return threadIdx.x & 31u; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Inclusive warp prefix sum.
__device__ __forceinline__ int warp_scan_inclusive(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if ((int)lane_id() >= o) v += t;
  }
  return v;
}

template <typename IdxT>
__device__ __forceinline__ int64_t load_index(const void* p, int64_t i) {
  return (int64_t) static_cast<const IdxT*>(p)[i];
}
#endif

// Exclusive prefix sum of per-cloud lengths on the device: off[0..n_clouds] (int64).
int launch_cloud_offsets(const int32_t* lens, int n_clouds, int64_t* off, cudaStream_t stream);

}  // namespace kpreg
