// Weight-gradient GEMM on tcgen05 (sm_100a):   C[Ma, Nb] += sum_k A[k, Ma] * B[k, Nb]      ("TN", split over k)
//
// The backward pass of every dense contraction of the path reduces over the ROWS of two row-major activations:
//   KPConv   d_weights[K*c_in, c_out] = agg^T g'          (reference finegrained_kpconv_blocks.py:388-393, autograd)
//   Linear   d_weight[n_out, n_in]    = d_out^T x         (UnaryBlock.mlp / my_Bottle2neck's layers)
// i.e. both operands are "MN-major": the output dimension is the contiguous one and the reduction index strides.
// tcgen05 takes such operands directly (instruction-descriptor bits 15 / 16 = MN-major A / B).  For 32-bit elements the
// MN-major shared-memory layout is the "128-byte swizzle with 32-byte atoms" (matrix-descriptor layout type 1,
// TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: the four 32-byte units of a 128-byte row are permuted by row mod 4): a
// [32 k x 32 floats] TMA box is one column of the canonical layout ((32 floats, m),(4, k)) — 4-row k groups 512 B apart,
// 32-float chunks 4096 B apart along M / N — so nothing is transposed in memory.  (With the plain 128-byte swizzle,
// layout type 2, the tensor core returns zeros for MN-major TF32 operands.)
//
// Precision: 3xTF32 as in kpconv_gemm.cu (A split hi/lo on the fly in shared memory by four splitter warps; B arrives
// pre-split: it is the small operand and is produced by an element-wise kernel anyway).  The reduction is long (k = number
// of points), so hi*hi rotates over three TMEM accumulators and the cross terms have their own.
//
// Work = (128 x BLOCK_N output tile) x (k range); the k dimension is cut so that ~2 work items per SM exist, every item
// adds its partial tile into C with red.global.add.f32 (C is zeroed by the caller).  The order of those additions is not
// fixed, so d_weights is reproducible to fp32 rounding only — as with the atomics of the CUDA-core kernel it replaces.
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace kpreg {
namespace {

using namespace tc;

constexpr int TN_BLOCK_M = 128;
constexpr int TN_BLOCK_K = 32;   // k rows per stage
constexpr int TN_CHUNK = 32;     // floats per 128-byte swizzle row (the contiguous M / N extent of one TMA box)
constexpr int TN_STAGES = 3;
constexpr uint32_t kChunkBytes = TN_BLOCK_K * TN_CHUNK * 4;  // 4 KiB: one [32 k x 32 floats] box
constexpr int TN_THREADS = 64 + 128 + 128;                   // TMA, MMA, 4 splitter warps, 4 epilogue warps

// MN-major, SWIZZLE_128B_BASE32B shared-memory matrix descriptor: leading byte offset = distance between 32-float chunks
// (4096 B), stride byte offset = distance between 4-row k groups (512 B), version 1 (sm_100), layout type 1.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(kChunkBytes >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) |
         (1ull << 61);
}
// kind::tf32, fp32 accumulate, A and B MN-major, shape M x N x 8.
__device__ __forceinline__ uint32_t make_instr_desc_mn(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BLOCK_N>
struct TnLayout {
  static constexpr uint32_t kABytes = (TN_BLOCK_M / TN_CHUNK) * kChunkBytes;  // 16 KiB
  static constexpr uint32_t kBBytes = (BLOCK_N / TN_CHUNK) * kChunkBytes;
  static constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;         // A hi | A lo | B hi | B lo
  static constexpr uint32_t kTileBytes = TN_STAGES * kStageBytes;
  static constexpr uint32_t kTotal = kTileBytes + 256 + 1024;
  static constexpr uint32_t kTmemCols = 4 * BLOCK_N;                         // 3 hi*hi accumulators + the cross terms
  static_assert(kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM allocation must be a power of two <= 512");
};

template <int BLOCK_N>
__global__ void __launch_bounds__(TN_THREADS, 1) k_gemm_tn(const __grid_constant__ CUtensorMap map_a,
                                                           const __grid_constant__ CUtensorMap map_b_hi,
                                                           const __grid_constant__ CUtensorMap map_b_lo, float* __restrict__ C,
                                                           int ldc, int Ma, int Nb, int num_kb, int kb_per_item, int n_splits, int c_transposed) {
  using L = TnLayout<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + L::kTileBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto split_bar = [&](int s) { return bar_base + 8u * (TN_STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * TN_STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (3 * TN_STAGES);
  const uint32_t tmem_empty_bar = bar_base + 8u * (3 * TN_STAGES + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + L::kTileBytes + 8u * (3 * TN_STAGES + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n = (Nb + BLOCK_N - 1) / BLOCK_N;
  const int num_m = (Ma + TN_BLOCK_M - 1) / TN_BLOCK_M;
  const uint32_t num_items = (uint32_t)(num_m * num_n * n_splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b_hi);
    tma_prefetch_desc(&map_b_lo);
    for (int s = 0; s < TN_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), 4);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(L::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (m block, n block, k-block range); every role walks the same static sequence
  auto decode = [&](uint32_t item, int& m0, int& n0, int& kb0, int& kb1) {
    const uint32_t tile = item / (uint32_t)n_splits;
    const int split = (int)(item - tile * (uint32_t)n_splits);
    const uint32_t mb = tile / (uint32_t)num_n;
    m0 = (int)mb * TN_BLOCK_M;
    n0 = (int)(tile - mb * (uint32_t)num_n) * BLOCK_N;
    kb0 = split * kb_per_item;
    kb1 = min(num_kb, kb0 + kb_per_item);
  };

  if (warp == 0) {
    // ---------------- TMA producer: per stage 4 boxes of A (one per 32-float chunk of M) and BLOCK_N/32 of B hi / lo
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t item = blockIdx.x; item < num_items; item += gridDim.x) {
        int m0, n0, kb0, kb1;
        decode(item, m0, n0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = (int)(it % TN_STAGES);
          mbar_wait(empty_bar(s), ((it / TN_STAGES) & 1u) ^ 1u);
          const uint32_t st = base + (uint32_t)s * L::kStageBytes;
          mbar_expect_tx(full_bar(s), L::kABytes + 2 * L::kBBytes);
#pragma unroll
          for (int c = 0; c < TN_BLOCK_M / TN_CHUNK; ++c) tma_load_2d(st + c * kChunkBytes, &map_a, full_bar(s), m0 + c * TN_CHUNK, kb * TN_BLOCK_K);
#pragma unroll
          for (int c = 0; c < BLOCK_N / TN_CHUNK; ++c) {
            tma_load_2d(st + 2 * L::kABytes + c * kChunkBytes, &map_b_hi, full_bar(s), n0 + c * TN_CHUNK, kb * TN_BLOCK_K);
            tma_load_2d(st + 2 * L::kABytes + L::kBBytes + c * kChunkBytes, &map_b_lo, full_bar(s), n0 + c * TN_CHUNK, kb * TN_BLOCK_K);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_instr_desc_mn(TN_BLOCK_M, BLOCK_N);
      uint32_t it = 0, li = 0;
      for (uint32_t item = blockIdx.x; item < num_items; item += gridDim.x, ++li) {
        int m0, n0, kb0, kb1;
        decode(item, m0, n0, kb0, kb1);
        mbar_wait(tmem_empty_bar, (li & 1u) ^ 1u);  // the epilogue has drained the accumulators of the previous item
        tcgen05_fence_after();
        const uint32_t acc_x = tmem_base + 3u * BLOCK_N;
        int ks = 0;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = (int)(it % TN_STAGES);
          mbar_wait(split_bar(s), (it / TN_STAGES) & 1u);
          tcgen05_fence_after();
          const uint32_t st = base + (uint32_t)s * L::kStageBytes;
          const uint64_t a_hi = make_smem_desc_mn(st), a_lo = make_smem_desc_mn(st + L::kABytes);
          const uint64_t b_hi = make_smem_desc_mn(st + 2 * L::kABytes), b_lo = make_smem_desc_mn(st + 2 * L::kABytes + L::kBBytes);
#pragma unroll
          for (int k = 0; k < TN_BLOCK_K / 8; ++k, ++ks) {
            const uint64_t adv = (uint64_t)((k * 1024) >> 4);  // next 8-row k group
            umma_tf32(acc_x, a_lo + adv, b_hi + adv, idesc, ks != 0 ? 1u : 0u);
            umma_tf32(acc_x, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_tf32(tmem_base + (uint32_t)(ks % 3) * BLOCK_N, a_hi + adv, b_hi + adv, idesc, ks >= 3 ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tmem_full_bar);
      }
    }
  } else if (warp < 6) {
    // ---------------- splitters (warps 2..5): A box -> hi (in place) and lo, element-wise (swizzle-agnostic)
    const int t = threadIdx.x - 64;  // 0..127
    uint32_t it = 0;
    for (uint32_t item = blockIdx.x; item < num_items; item += gridDim.x) {
      int m0, n0, kb0, kb1;
      decode(item, m0, n0, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = (int)(it % TN_STAGES);
        mbar_wait(full_bar(s), (it / TN_STAGES) & 1u);
        float4* hi = reinterpret_cast<float4*>(base_ptr + (size_t)s * L::kStageBytes);
        float4* lo = reinterpret_cast<float4*>(base_ptr + (size_t)s * L::kStageBytes + L::kABytes);
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = hi[t + 128 * j];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 h, l;
          split_tf32_fast(v[j].x, h.x, l.x);
          split_tf32_fast(v[j].y, h.y, l.y);
          split_tf32_fast(v[j].z, h.z, l.z);
          split_tf32_fast(v[j].w, h.w, l.w);
          hi[t + 128 * j] = h;
          lo[t + 128 * j] = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(s));
      }
    }
  } else {
    // ---------------- epilogue (warps 6..9): warp w owns TMEM lanes 32*(w%4) .. +31; thread = one row of C
    const int q = warp & 3;
    uint32_t li = 0;
    for (uint32_t item = blockIdx.x; item < num_items; item += gridDim.x, ++li) {
      int m0, n0, kb0, kb1;
      decode(item, m0, n0, kb0, kb1);
      mbar_wait(tmem_full_bar, li & 1u);
      tcgen05_fence_after();
      const int m = m0 + 32 * q + lane;
      const uint32_t acc0 = tmem_base + ((uint32_t)(32 * q) << 16);
      const int n_ks = (kb1 - kb0) * (TN_BLOCK_K / 8);  // accumulators never written (fewer than 3 k-steps) hold stale data
#pragma unroll
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32], r2[32];
        float sum[32];
        tmem_ld_32x32b_x32(acc0 + (uint32_t)c0, r);
        tmem_ld_32x32b_x32(acc0 + (uint32_t)(3 * BLOCK_N + c0), r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
        tmem_ld_32x32b_x32(acc0 + (uint32_t)(BLOCK_N + c0), r);
        tmem_ld_32x32b_x32(acc0 + (uint32_t)(2 * BLOCK_N + c0), r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[j] += (n_ks > 1 ? __uint_as_float(r[j]) : 0.f) + (n_ks > 2 ? __uint_as_float(r2[j]) : 0.f);
        if (m < Ma) {
          if (c_transposed) {  // C holds the transposed product: element (m, n) lives at C[n * ldc + m] (coalesced across lanes)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < Nb) atomicAdd(C + (int64_t)(n0 + c0 + j) * ldc + m, sum[j]);
          } else {
            float* __restrict__ crow = C + (int64_t)m * ldc + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < Nb) atomicAdd(crow + j, sum[j]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(L::kTmemCols));
  }
}

// B -> hi / lo TF32 operand pair (optionally scaled per row: g' = grad_out * inv_num of KPConv's normalisation)
__global__ void __launch_bounds__(256) k_split_rows(const float* __restrict__ in, int ld_in, const float* __restrict__ row_scale,
                                                    int64_t rows, int cols, float* __restrict__ hi, float* __restrict__ lo, int ld_out) {
  const int64_t total = rows * (int64_t)ld_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_out;
    const int c = (int)(i - r * ld_out);
    float v = 0.f;
    if (c < cols) v = in[r * ld_in + c] * (row_scale ? row_scale[r] : 1.0f);
    float h, l;
    split_tf32(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tn_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows (k), cols (contiguous M / N)] fp32, row pitch ld floats; box = [32 k rows, 32 floats], SWIZZLE_128B_ATOM_32B
int make_map_mn(CUtensorMap* map, const float* ptr, int64_t rows, int cols, int64_t ld) {
  EncodeTiledFn fn = tn_encode_fn();
  if (!fn) return KPREG_E_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TN_CHUNK, (cuuint32_t)TN_BLOCK_K};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled", cudaErrorInvalidValue);
    return KPREG_E_CUDA;
  }
  return KPREG_OK;
}

template <int BLOCK_N>
int launch_tn(const CUtensorMap& ma, const CUtensorMap& mbh, const CUtensorMap& mbl, float* c, int ldc, int ma_dim, int nb_dim,
              int64_t k_dim, int c_transposed, cudaStream_t stream) {
  using L = TnLayout<BLOCK_N>;
  static PerDeviceOnce once;
  const int rc_cfg = once.run([]() -> int {
    KP_CUDA_TRY(cudaFuncSetAttribute(k_gemm_tn<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kTotal));
    return KPREG_OK;
  });
  if (rc_cfg) return rc_cfg;
  const int num_kb = ceil_div(k_dim, TN_BLOCK_K);
  const int tiles = ceil_div(ma_dim, TN_BLOCK_M) * ceil_div(nb_dim, BLOCK_N);
  int splits = (2 * kNumSMs + tiles - 1) / tiles;       // ~2 work items per SM
  if (splits > ceil_div(num_kb, 4)) splits = ceil_div(num_kb, 4);  // at least 4 k-blocks per item
  if (splits < 1) splits = 1;
  const int kb_per_item = ceil_div(num_kb, splits);
  splits = ceil_div(num_kb, kb_per_item);
  const int64_t items = (int64_t)tiles * splits;
  const unsigned grid = (unsigned)(items < kNumSMs ? items : kNumSMs);
  k_gemm_tn<BLOCK_N><<<grid, TN_THREADS, L::kTotal, stream>>>(ma, mbh, mbl, c, ldc, ma_dim, nb_dim, num_kb, kb_per_item, splits, c_transposed);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

}  // namespace

size_t gemm_tn_workspace_bytes(int64_t k_rows, int nb_dim) {
  const size_t ld = (size_t)((nb_dim + 3) / 4 * 4);
  return align_up((size_t)2 * (size_t)(k_rows > 0 ? k_rows : 1) * ld * sizeof(float) + 512, 256);
}

bool gemm_tn_supported(int64_t k_rows, int ma_dim, int nb_dim, int lda, const void* a) {
  return k_rows > 0 && k_rows < ((int64_t)1 << 31) && ma_dim >= 8 && nb_dim >= 8 && (lda % 4) == 0 &&
         (reinterpret_cast<uintptr_t>(a) % 16) == 0;
}

// C[ma_dim, nb_dim] (row pitch ldc, ZEROED by the caller) += A^T (B * row_scale):  A [k_rows, lda] (ma_dim columns used),
// B [k_rows, ldb] (nb_dim columns used), row_scale [k_rows] or null; c_transposed: C is [nb_dim, ma_dim] (holds the transpose).  `split_ws` holds gemm_tn_workspace_bytes(k_rows, nb_dim).
int launch_gemm_tn(const float* a, int lda, const float* b, int ldb, const float* row_scale, float* c, int ldc, int64_t k_rows,
                   int ma_dim, int nb_dim, int c_transposed, void* split_ws, cudaStream_t stream) {
  if (!gemm_tn_supported(k_rows, ma_dim, nb_dim, lda, a)) return KPREG_E_INVALID;
  const int ld_s = (nb_dim + 3) / 4 * 4;
  float* hi = static_cast<float*>(split_ws);
  float* lo = hi + (size_t)k_rows * ld_s;
  {
    int blocks = ceil_div(k_rows * (int64_t)ld_s, 256);
    if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
    k_split_rows<<<blocks, 256, 0, stream>>>(b, ldb, row_scale, k_rows, nb_dim, hi, lo, ld_s);
    KP_LAUNCH_CHECK();
  }
  CUtensorMap ma, mbh, mbl;
  int rc = make_map_mn(&ma, a, k_rows, ma_dim, lda);
  if (rc) return rc;
  rc = make_map_mn(&mbh, hi, k_rows, nb_dim, ld_s);
  if (rc) return rc;
  rc = make_map_mn(&mbl, lo, k_rows, nb_dim, ld_s);
  if (rc) return rc;
  if (nb_dim <= 32) return launch_tn<32>(ma, mbh, mbl, c, ldc, ma_dim, nb_dim, k_rows, c_transposed, stream);
  if (nb_dim <= 64) return launch_tn<64>(ma, mbh, mbl, c, ldc, ma_dim, nb_dim, k_rows, c_transposed, stream);
  return launch_tn<128>(ma, mbh, mbl, c, ldc, ma_dim, nb_dim, k_rows, c_transposed, stream);
}

}  // namespace kpreg
