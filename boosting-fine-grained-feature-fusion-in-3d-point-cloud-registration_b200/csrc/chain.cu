// The hierarchical chain of the fork's "fine-grained feature fusion" unit in ONE kernel.
//
// my_Bottle2neck.forward (reference models/backbone_kpconv/res2net.py:137-150) splits conv1's output t into `scale`
// groups of `width` channels and runs
//     sp = t_0;  for i in 0..scale-2:  sp = relu(bn_i(conv_i(sp)));  out_i = sp;  sp = sp + t_{i+1}
// then concatenates out_0 .. out_{scale-2} and the untouched last group.  As separate layers that is seven dependent
// [N, w] x [w, w] GEMM launches which each read two [N, w] tensors and write two — latency-bound, 2 TB/s on B200.
// Here a warp keeps 32 rows of the running activation in REGISTERS across all layers: t is read once, the concatenated
// tensor written once (2 * scale * w * 4 bytes per row instead of ~4x that), nothing else touches HBM.
//
// The products run on mma.sync.m16n8k8 (TF32, fp32 accumulate) with the 3xTF32 operand split of kpconv_gemm.cu
// (fp32-grade accuracy).  The MMA's K and N indices are only labels, so they are bound to channels such that
//   * the D fragment of layer i IS the A fragment of layer i + 1 (k-slot t <- column 2t, k-slot t+4 <- column 2t+1):
//     no shuffle, no shared-memory round trip between layers;
//   * thread t of a quad owns the 2 * NT contiguous channels [2 NT t, 2 NT (t + 1)) of rows g and g + 8, so global
//     loads / stores are contiguous float2 runs.
// The (BatchNorm-folded) weights of all layers are pre-arranged by k_chain_pack in exactly the order the lanes read
// their B fragments (hi and lo parts, zero-padded to 8 NT channels) and stay resident in shared memory.
#include "common.cuh"

namespace kpreg {
namespace {

constexpr int kChainWarps = 8;
constexpr int kChainRows = 32 * kChainWarps;  // rows per CTA pass (two m16 tiles per warp)

__device__ __forceinline__ void chain_mma(float (&d)[4], const float (&a)[4], float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                 "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
// x = hi + lo: hi = x rounded to the nearest TF32 by integer arithmetic on the bit pattern, lo = x - hi exact in fp32 and
// left unrounded (the tensor core reads only the upper 19 bits of a TF32 operand: an error of 2^-21 |x|, below the lo*lo
// term 3xTF32 drops anyway).
__device__ __forceinline__ void chain_split(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// pack layout (floats): [layer][jk][jn][hi|lo][lane][2] fragments, then [layer][8 NT] shifts (natural channel order)
__host__ __device__ inline size_t chain_frag_floats(int nt, int n_layers) { return (size_t)n_layers * nt * nt * 128; }
__host__ __device__ inline size_t chain_pack_floats(int nt, int n_layers) { return chain_frag_floats(nt, n_layers) + (size_t)n_layers * 8 * nt; }

__global__ void __launch_bounds__(256) k_chain_pack(const float* __restrict__ weights, const float* __restrict__ shifts, int w,
                                                    int nt, int n_layers, float* __restrict__ pack) {
  const size_t n_frag = chain_frag_floats(nt, n_layers), total = chain_pack_floats(nt, n_layers);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n_frag) {
      const int e = (int)(i & 1), lane = (int)((i >> 1) & 31), hl = (int)((i >> 6) & 1);
      const size_t tile = i >> 7;
      const int jn = (int)(tile % nt), jk = (int)((tile / nt) % nt), layer = (int)(tile / ((size_t)nt * nt));
      const int g = lane >> 2, t = lane & 3;
      const int n_act = (g >> 1) * 2 * nt + 2 * jn + (g & 1);  // channel of D column g of n-tile jn
      const int k_act = 2 * nt * t + 2 * jk + e;                // channel of A k-slot t (e = 0) / t + 4 (e = 1) of k-tile jk
      float v = 0.f;
      if (n_act < w && k_act < w) v = weights[((size_t)layer * w + n_act) * w + k_act];
      const float hi = tf32_rna(v);
      pack[i] = hl ? tf32_rna(v - hi) : hi;
    } else {
      const size_t j = i - n_frag;
      const int c = (int)(j % (8 * nt)), layer = (int)(j / (8 * nt));
      pack[i] = c < w ? shifts[(size_t)layer * w + c] : 0.f;
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(kChainWarps * 32, NT <= 4 ? 2 : 1) k_chain(
    const float* __restrict__ t_in, int ld_t, const float* __restrict__ pack, int w, int n_layers, int64_t m_rows,
    float* __restrict__ z, int ld_z, const float* __restrict__ x_copy, int ld_x, int c_x) {
  extern __shared__ __align__(16) float s_pack[];
  {
    const int total4 = (int)(chain_pack_floats(NT, n_layers) / 4);  // 128 NT^2 + 8 NT floats per layer: a multiple of 4
    const float4* __restrict__ src = reinterpret_cast<const float4*>(pack);
    float4* dst = reinterpret_cast<float4*>(s_pack);
    for (int i = threadIdx.x; i < total4; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const float* __restrict__ s_shift = s_pack + chain_frag_floats(NT, n_layers);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ch0 = 2 * NT * t;  // first of this thread's 2 NT channels

  for (int64_t blk = blockIdx.x; blk * kChainRows < m_rows; blk += gridDim.x) {
    const int64_t r0 = blk * kChainRows + warp * 32;
    if (r0 >= m_rows) continue;
    // element [m][j][e]: row r0 + 16 m + g + 8 (e >> 1), channel ch0 + 2 j + (e & 1)  (the mma D-fragment order)
    auto load_group = [&](int grp, float (&dst)[2][NT][4]) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int64_t row = r0 + 16 * m + g + 8 * half;
          const float* __restrict__ p = t_in + row * ld_t + (int64_t)grp * w + ch0;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            float2 v = make_float2(0.f, 0.f);
            if (row < m_rows && ch0 + 2 * j < w) v = __ldg(reinterpret_cast<const float2*>(p + 2 * j));
            dst[m][j][2 * half] = v.x;
            dst[m][j][2 * half + 1] = v.y;
          }
        }
    };
    auto store_group = [&](int grp, const float (&src)[2][NT][4]) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int64_t row = r0 + 16 * m + g + 8 * half;
          float* __restrict__ p = z + row * ld_z + (int64_t)grp * w + ch0;
#pragma unroll
          for (int j = 0; j < NT; ++j)
            if (row < m_rows && ch0 + 2 * j < w) *reinterpret_cast<float2*>(p + 2 * j) = make_float2(src[m][j][2 * half], src[m][j][2 * half + 1]);
        }
    };

    float cur[2][NT][4];
    load_group(0, cur);
    for (int layer = 0; layer < n_layers; ++layer) {
      float nxt[2][NT][4];
      load_group(layer + 1, nxt);  // consumed after this layer's MMAs: the loads fly under them
      float acc[2][NT][4];
      {
        const float* __restrict__ sh = s_shift + layer * 8 * NT + ch0;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float2 b = *reinterpret_cast<const float2*>(sh + 2 * j);
#pragma unroll
          for (int m = 0; m < 2; ++m) { acc[m][j][0] = b.x; acc[m][j][1] = b.y; acc[m][j][2] = b.x; acc[m][j][3] = b.y; }
        }
      }
      const float* __restrict__ fw = s_pack + (size_t)layer * NT * NT * 128 + lane * 2;
#pragma unroll
      for (int jk = 0; jk < NT; ++jk) {
        // A fragment = the previous layer's D fragment: a0 (g, slot t) = d0, a1 (g+8, t) = d2, a2 (g, t+4) = d1, a3 (g+8, t+4) = d3
        float a_hi[2][4], a_lo[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          chain_split(cur[m][jk][0], a_hi[m][0], a_lo[m][0]);
          chain_split(cur[m][jk][2], a_hi[m][1], a_lo[m][1]);
          chain_split(cur[m][jk][1], a_hi[m][2], a_lo[m][2]);
          chain_split(cur[m][jk][3], a_hi[m][3], a_lo[m][3]);
        }
#pragma unroll
        for (int jn = 0; jn < NT; ++jn) {
          const float2 bh = *reinterpret_cast<const float2*>(fw + (jk * NT + jn) * 128);
          const float2 bl = *reinterpret_cast<const float2*>(fw + (jk * NT + jn) * 128 + 64);
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            chain_mma(acc[m][jn], a_lo[m], bh.x, bh.y);
            chain_mma(acc[m][jn], a_hi[m], bl.x, bl.y);
            chain_mma(acc[m][jn], a_hi[m], bh.x, bh.y);
          }
        }
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[m][j][e] = fmaxf(acc[m][j][e], 0.f);
      store_group(layer, acc);
      if (layer + 1 < n_layers) {
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) cur[m][j][e] = acc[m][j][e] + nxt[m][j][e];
      } else {
        store_group(n_layers, nxt);  // the last group passes through unchanged
      }
    }
    if (x_copy != nullptr) {
      // the block input rides along behind the concatenation (the K-concatenated residual projection of conv3)
      const int64_t rows = min((int64_t)32, m_rows - r0);
      float* __restrict__ zc = z + (int64_t)(n_layers + 1) * w;
      if ((c_x & 3) == 0 && (ld_x & 3) == 0 && (ld_z & 3) == 0 && (((n_layers + 1) * w) & 3) == 0 &&
          ((reinterpret_cast<uintptr_t>(x_copy) | reinterpret_cast<uintptr_t>(z)) & 15) == 0) {
        const int c4 = c_x >> 2;
        for (int i = lane; i < (int)rows * c4; i += 32) {
          const int r = i / c4, c = (i - r * c4) * 4;
          *reinterpret_cast<float4*>(zc + (r0 + r) * ld_z + c) = __ldg(reinterpret_cast<const float4*>(x_copy + (r0 + r) * ld_x + c));
        }
      } else {
        for (int i = lane; i < (int)rows * c_x; i += 32) {
          const int r = i / c_x, c = i - r * c_x;
          zc[(r0 + r) * ld_z + c] = x_copy[(r0 + r) * ld_x + c];
        }
      }
    }
  }
}

int chain_nt(int width) { return (width + 7) / 8; }
size_t chain_smem_bytes(int width, int n_layers) { return chain_pack_floats(chain_nt(width), n_layers) * sizeof(float); }

bool chain_supported(int width, int n_layers) {
  const int nt = chain_nt(width);
  if (width < 2 || (width & 1) || n_layers < 1 || n_layers > 64) return false;
  if (!(nt == 2 || nt == 4 || nt == 7 || nt == 8)) return false;  // instantiated tile counts (width 28 -> 4, 56 -> 7)
  return chain_smem_bytes(width, n_layers) <= 200 * 1024;
}

template <int NT>
int launch_chain(const float* t, int ld_t, const float* pack, int w, int n_layers, int64_t m_rows, float* z, int ld_z,
                 const float* x_copy, int ld_x, int c_x, cudaStream_t stream) {
  const size_t smem = chain_smem_bytes(w, n_layers);
  static PerDeviceOnce once;  // per device: the largest footprint kpreg_chain_supported admits
  const int rc_cfg = once.run([]() -> int {
    KP_CUDA_TRY(cudaFuncSetAttribute(k_chain<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return KPREG_OK;
  });
  if (rc_cfg) return rc_cfg;
  const int per_sm = (NT <= 4 && 2 * smem <= 200 * 1024) ? 2 : 1;
  int blocks = ceil_div(m_rows, kChainRows);
  if (blocks > per_sm * kNumSMs) blocks = per_sm * kNumSMs;
  k_chain<NT><<<blocks, kChainWarps * 32, smem, stream>>>(t, ld_t, pack, w, n_layers, m_rows, z, ld_z, x_copy, ld_x, c_x);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

}  // namespace
}  // namespace kpreg

using namespace kpreg;

extern "C" int kpreg_chain_supported(int width, int n_layers) { return chain_supported(width, n_layers) ? 1 : 0; }

extern "C" int kpreg_chain_pack_bytes(int width, int n_layers, size_t* bytes) {
  if (!bytes || !chain_supported(width, n_layers)) return KPREG_E_INVALID;
  *bytes = align_up(chain_smem_bytes(width, n_layers), 256);
  return KPREG_OK;
}

extern "C" int kpreg_chain_pack(const float* weights, const float* shifts, int width, int n_layers, void* pack, size_t pack_bytes,
                                void* stream_) {
  if (!weights || !shifts || !pack || !chain_supported(width, n_layers)) return KPREG_E_INVALID;
  if (pack_bytes < chain_smem_bytes(width, n_layers)) return KPREG_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(pack) & 15) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nt = chain_nt(width);
  const int blocks = ceil_div((int64_t)chain_pack_floats(nt, n_layers), 256);
  k_chain_pack<<<blocks, 256, 0, stream>>>(weights, shifts, width, nt, n_layers, static_cast<float*>(pack));
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

extern "C" int kpreg_chain_forward(const float* t, int ld_t, const void* pack, int width, int n_layers, int64_t m_rows, float* z,
                                   int ld_z, const float* x_copy, int ld_x, int c_x, void* stream_) {
  if (!chain_supported(width, n_layers) || m_rows < 0) return KPREG_E_INVALID;
  if (m_rows == 0) return KPREG_OK;
  const int groups = n_layers + 1;
  if (!t || !pack || !z || ld_t < groups * width || (ld_t & 1) || (ld_z & 1)) return KPREG_E_INVALID;
  if (ld_z < groups * width + (x_copy ? c_x : 0) || (x_copy && (c_x < 1 || ld_x < c_x))) return KPREG_E_INVALID;
  if (((reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(z)) & 7) || (reinterpret_cast<uintptr_t>(pack) & 15)) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  ProfScope prof(KPREG_FAM_LINEAR, stream);
  const float* pk = static_cast<const float*>(pack);
  switch (chain_nt(width)) {
    case 2: return launch_chain<2>(t, ld_t, pk, width, n_layers, m_rows, z, ld_z, x_copy, ld_x, c_x, stream);
    case 4: return launch_chain<4>(t, ld_t, pk, width, n_layers, m_rows, z, ld_z, x_copy, ld_x, c_x, stream);
    case 7: return launch_chain<7>(t, ld_t, pk, width, n_layers, m_rows, z, ld_z, x_copy, ld_x, c_x, stream);
    case 8: return launch_chain<8>(t, ld_t, pk, width, n_layers, m_rows, z, ld_z, x_copy, ld_x, c_x, stream);
  }
  return KPREG_E_INVALID;
}
