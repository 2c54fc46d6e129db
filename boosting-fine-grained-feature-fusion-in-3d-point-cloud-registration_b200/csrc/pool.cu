// Strided-block shortcut pooling on sm_100a.
//
// Replaces max_pool() (reference models/backbone_kpconv/finegrained_kpconv_blocks.py:125-141):
// out[n,c] = max over h of [x; 0][idx[n,h], c] — a shadow index selects the appended zero row, so it
// takes part in the max.  One warp per query row, lanes over channels (coalesced row reads).
#include "common.cuh"

namespace kpreg {
namespace {

// ARGMAX = false (inference): the winning row is not tracked — a third fewer instructions in the inner loop
template <typename IdxT, bool ARGMAX>
__global__ void __launch_bounds__(256) k_max_pool(const float* __restrict__ x, const IdxT* __restrict__ idx, int64_t n_q, int64_t n_s,
                                                  int n_nbrs, int channels, float* __restrict__ out, int32_t* __restrict__ argmax,
                                                  const int32_t* __restrict__ order) {
  const int lane = threadIdx.x & 31;
  const int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (it >= n_q) return;
  const int64_t n = order ? (int64_t)order[it] : it;  // processing order only
  const bool vec = (channels & 3) == 0;  // float4 per lane: 128 channels per pass
  const int step = vec ? 128 : 32;
  for (int c0 = 0; c0 < channels; c0 += step) {
    const int c = c0 + (vec ? 4 * lane : lane);
    float best[4] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
    int32_t best_j[4] = {(int32_t)n_s, (int32_t)n_s, (int32_t)n_s, (int32_t)n_s};
    for (int h0 = 0; h0 < n_nbrs; h0 += 32) {
      int64_t j = n_s;
      if (h0 + lane < n_nbrs) j = (int64_t)idx[n * n_nbrs + h0 + lane];
      const int lim = min(32, n_nbrs - h0);
#pragma unroll 4
      for (int hh = 0; hh < lim; ++hh) {
        int64_t jj = __shfl_sync(0xffffffffu, j, hh);
        const bool real = jj >= 0 && jj < n_s;
        if (!real) jj = n_s;  // shadow row: zeros
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (real && c < channels) {
          if (vec) v = *reinterpret_cast<const float4*>(x + jj * channels + c);
          else v.x = x[jj * channels + c];
        }
        // strict > keeps the first maximum, like a sequential scan over the row
        if constexpr (ARGMAX) {
          if (v.x > best[0]) { best[0] = v.x; best_j[0] = (int32_t)jj; }
          if (v.y > best[1]) { best[1] = v.y; best_j[1] = (int32_t)jj; }
          if (v.z > best[2]) { best[2] = v.z; best_j[2] = (int32_t)jj; }
          if (v.w > best[3]) { best[3] = v.w; best_j[3] = (int32_t)jj; }
        } else {
          best[0] = fmaxf(best[0], v.x); best[1] = fmaxf(best[1], v.y);
          best[2] = fmaxf(best[2], v.z); best[3] = fmaxf(best[3], v.w);
        }
      }
    }
    if (c < channels) {
      if (n_nbrs == 0) best[0] = best[1] = best[2] = best[3] = 0.f;
      if (vec) {
        *reinterpret_cast<float4*>(out + n * channels + c) = make_float4(best[0], best[1], best[2], best[3]);
        if (argmax) *reinterpret_cast<int4*>(argmax + n * channels + c) = make_int4(best_j[0], best_j[1], best_j[2], best_j[3]);
      } else {
        out[n * channels + c] = best[0];
        if (argmax) argmax[n * channels + c] = best_j[0];
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_max_pool_bwd(const float* __restrict__ g, const int32_t* __restrict__ argmax, int64_t n_q,
                                                      int64_t n_s, int channels, float* __restrict__ d_x) {
  const int64_t total = n_q * (int64_t)channels;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = argmax[i];
    if (j >= 0 && j < n_s) atomicAdd(d_x + (int64_t)j * channels + (i % channels), g[i]);
  }
}

}  // namespace
}  // namespace kpreg

using namespace kpreg;

extern "C" int kpreg_max_pool_forward(const float* x, const void* idx, int idx64, int64_t n_q, int64_t n_s, int n_nbrs,
                                      int channels, const int32_t* order, float* out, int32_t* argmax, void* stream_) {
  if (n_q < 0 || n_s < 0 || n_nbrs < 0 || channels < 1) return KPREG_E_INVALID;
  if (n_q == 0) return KPREG_OK;
  if (!out || (n_nbrs > 0 && !idx) || (n_s > 0 && !x)) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int blocks = ceil_div(n_q * 32, 256);
  ProfScope prof(KPREG_FAM_POOL, stream);
  if (argmax) {
    if (idx64) k_max_pool<int64_t, true><<<blocks, 256, 0, stream>>>(x, static_cast<const int64_t*>(idx), n_q, n_s, n_nbrs, channels, out, argmax, order);
    else k_max_pool<int32_t, true><<<blocks, 256, 0, stream>>>(x, static_cast<const int32_t*>(idx), n_q, n_s, n_nbrs, channels, out, argmax, order);
  } else {
    if (idx64) k_max_pool<int64_t, false><<<blocks, 256, 0, stream>>>(x, static_cast<const int64_t*>(idx), n_q, n_s, n_nbrs, channels, out, argmax, order);
    else k_max_pool<int32_t, false><<<blocks, 256, 0, stream>>>(x, static_cast<const int32_t*>(idx), n_q, n_s, n_nbrs, channels, out, argmax, order);
  }
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

extern "C" int kpreg_max_pool_backward(const float* grad_out, const int32_t* argmax, int64_t n_q, int64_t n_s, int channels,
                                       float* d_x, void* stream_) {
  if (n_q < 0 || n_s < 0 || channels < 1) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_s > 0) {
    if (!d_x) return KPREG_E_INVALID;
    KP_CUDA_TRY(cudaMemsetAsync(d_x, 0, sizeof(float) * (size_t)n_s * (size_t)channels, stream));
  }
  if (n_q == 0 || n_s == 0) return KPREG_OK;
  if (!grad_out || !argmax) return KPREG_E_INVALID;
  int blocks = ceil_div(n_q * (int64_t)channels, 256);
  if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
  k_max_pool_bwd<<<blocks, 256, 0, stream>>>(grad_out, argmax, n_q, n_s, channels, d_x);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}
