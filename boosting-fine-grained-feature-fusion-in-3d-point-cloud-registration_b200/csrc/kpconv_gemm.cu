// fp32-accurate tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32) with TMEM accumulators, fed by TMA.
//
//   C[M,N] = epilogue( A[M,K] * B^T ),   A row-major [M,K] fp32,  B given K-major as Bt[N,K]
//
// This is the one dense contraction of the KPConv path — the [n_q, K*c_in] x [K*c_in, c_out] product
// of KPConv.forward (reference models/backbone_kpconv/finegrained_kpconv_blocks.py:388-393) — and the
// same kernel serves the Linear layers of the encoder blocks.
//
// Precision.  The reference computes in fp32 and parity is 1e-4 relative, which a single TF32 pass
// (10-bit mantissa) does not meet.  Every fp32 operand is therefore split x = hi + lo with
// hi = rn_tf32(x) and lo = x - hi (exact in fp32), and the product is accumulated in fp32 TMEM as
// A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (3xTF32; what is dropped — lo*lo and the tensor core's truncation of lo to
// 11 bits — is ~2^-21 relative).  B (the weights) is split once (k_split_weights, cvt.rna for hi and lo); A is split on
// the fly in shared memory by the CTA's four "splitter" warps (integer round-to-nearest, 3 instructions per element),
// so A crosses HBM once, as plain fp32.
//
// Persistent CTA = 320 threads, 128 x BLOCK_N output tiles, K in blocks of 32 floats (one 128-byte swizzle row):
//   warp 0    TMA producer: per stage one box of A (128 x 32) and two of Bt (BLOCK_N x 32: hi, lo),
//             SWIZZLE_128B, completion on full[stage]; runs ahead across tile boundaries
//   warp 1    TMEM allocation; one elected lane issues 12 tcgen05.mma per stage (4 k-steps x 3 products),
//             tcgen05.commit frees the stage (empty[stage]) and signals tmem_full[buffer] per tile
//   warps 2-5 splitters: wait full[stage], rewrite the A box in place as hi and write lo to a second
//             box (element-wise, so swizzle-agnostic), fence.proxy.async, arrive on split_done[stage]
//   warps 6-9 epilogue (6-13 for tiles >= 64 columns): tcgen05.ld their 32 TMEM lanes from the finished accumulator
//             buffer, release it (tmem_empty[buffer]), apply row scale / column scale+shift / residual / activation /
//             post-activation shortcut and leave through a swizzled staging tile + TMA store — while the next
//             tile's MMAs run.
//
// Measured lessons built into the epilogue (in-kernel clock64 traces of every role, B200):
//   * on short-K layers (K <= 128: one to four k-blocks per tile) the epilogue warps, not HBM, set the tile rate.  A
//     per-column body that re-tested every run-time option (residual? activation? second output? edge?) cost ~4 300
//     cycles per 32-column chunk; testing each option once per chunk and issuing a side input's loads together brought
//     conv1 of the finest level from 1 270 to 600 us (4.5 TB/s) and the unary layers to 5.3 TB/s;
//   * column scale / shift arrive by one coalesced load per lane issued before the accumulator wait + shuffles (with
//     227 KB of the SM given to shared memory, eight broadcast float4 loads per chunk missed L1);
//   * tile coordinates use 32-bit division (a 64-bit division by a run-time num_n is a ~100-instruction routine, and
//     every role ran it per tile);
//   * a staged tile flushed by coalesced STG.128 instead of the TMA store was 20-40 % slower.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace kpreg {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;          // floats per 128-byte swizzle row
constexpr int UMMA_K = 8;            // tf32
// warps: TMA, MMA, 4 splitters, 4 | 8 epilogue warps.  (Measured for the A-through-TMEM variant: 8 splitters — two per TMEM
// lane quarter, 16 of the k-block's 32 columns each — with 4 epilogue warps were 2-10 % slower than 4 + 8: the splitters do not
// pace the long-K tiles.)
constexpr int gemm_split_warps(bool at) { return 4; }
constexpr int gemm_epi_warps(int block_n, bool at) { return block_n >= 64 ? 8 : 4; }
constexpr int gemm_threads(int block_n, bool at) { return 64 + 32 * gemm_split_warps(at) + 32 * gemm_epi_warps(block_n, at); }
constexpr uint32_t kABoxBytes = BLOCK_M * BLOCK_K * 4;  // 16 KiB

using namespace tc;  // PTX wrappers (tc_ptx.cuh)

// (the fp16 operand split helpers — split_h2, SWIZZLE_64B descriptors, kind::f16 MMA — live in tc_ptx.cuh)

struct Epilogue {
  const float* row_scale;  // [M] or null
  const float* col_scale;  // [N] or null
  const float* col_shift;  // [N] or null
  const float* residual;   // [M, ld_res] or null, added before the activation
  int ld_res;
  int act;                 // 0 none, 1 relu, 2 leaky relu (slope below)
  float slope;
  float* out2;             // optional second output [M, ld2] = C + addend (the next chained layer's input)
  const float* addend;     // [M, ld_add]
  int ld2, ld_add;
  int vec_ok;              // every row pointer (C, residual, out2, addend) is 16-byte aligned: float4 accesses allowed
  int tma_store;           // C (and out2) leave through TMA stores from a swizzled shared-memory tile (needs vec_ok)
  const float* post_res;   // [M, ld_post] or null: added AFTER the activation, followed by post_act (a block's shortcut)
  int ld_post;
  int post_act;            // 0 none, 1 relu, 2 leaky relu (same slope)
};

// TMEM accumulators per tile: NUM_HI for the hi*hi products (round-robin over k-steps) + 1 for the cross terms.
// H2 = fp16 operand split: stage = [A raw fp32 (TMA) | A hi f16 | A lo f16 | B hi f16 | B lo f16], rows of 32 halves (SWIZZLE_64B)
// AT (with H2) = the split A operand lives in tensor memory: the splitter warps write hi / lo with tcgen05.st (32 columns per
// stage) and the MMAs take A from TMEM — the stage holds only the raw A box and B, and the MMAs read only B from shared memory.
template <int BLOCK_N, int NUM_HI, int STAGES, int ACC_BUFS, bool H2 = false, bool AT = false>
struct SmemLayout {
  static constexpr int kNumAcc = NUM_HI + 1;
  static constexpr uint32_t kAHalfBytes = BLOCK_M * BLOCK_K * 2;  // 8 KiB
  static constexpr uint32_t kBBoxBytes = H2 ? BLOCK_N * BLOCK_K * 2 : BLOCK_N * BLOCK_K * 4;
  static constexpr uint32_t kBOffset = AT ? kABoxBytes : (H2 ? kABoxBytes + 2 * kAHalfBytes : 2 * kABoxBytes);  // B hi inside a stage
  static constexpr uint32_t kStageBytes = kBOffset + 2 * kBBoxBytes;
  static constexpr uint32_t kTileBytes = STAGES * kStageBytes;
  static constexpr int kSplitWarps = gemm_split_warps(AT);
  static constexpr int kEpiWarps = gemm_epi_warps(BLOCK_N, AT);  // eight: two warps per TMEM lane quarter share a tile's 32-column chunks
  static constexpr uint32_t kStoreBytes = kEpiWarps * 4096;   // per epilogue warp: one 32 x 32 fp32 tile for TMA stores
  static constexpr uint32_t kEpiBytes = kStoreBytes;
  static constexpr uint32_t kBarrierBytes = 256;
  static constexpr uint32_t kTotal = kTileBytes + kEpiBytes + kBarrierBytes + 1024;  // + slack for the 1024-byte alignment
  static constexpr uint32_t kAccCols = ACC_BUFS * kNumAcc * BLOCK_N;  // ACC_BUFS = 2: double-buffered accumulators
  static constexpr uint32_t kATmemCols = AT ? STAGES * BLOCK_K : 0;    // per stage: 16 columns of hi pairs, 16 of lo pairs
  static constexpr uint32_t kTmemUsed = kAccCols + kATmemCols;
  static constexpr uint32_t kTmemCols = kTmemUsed <= 128 ? 128 : (kTmemUsed <= 256 ? 256 : 512);
  static_assert(kTmemUsed <= 512, "TMEM holds 512 columns");
  static_assert(!AT || H2, "the TMEM A operand is the fp16 split");
  static_assert(kTotal <= 232448, "shared memory budget of one CTA");
  static_assert(kStageBytes % 1024 == 0 && kBOffset % 1024 == 0 && kBBoxBytes % 1024 == 0, "swizzled boxes stay 1024-byte aligned");
};

// Persistent kernel: grid = min(#tiles, #SMs); every role walks the same static tile sequence
// tile = blockIdx.x + i * gridDim.x, (m_blk, n_blk) = (tile / num_n, tile % num_n).  The shared-memory ring and
// its phases run on across tiles, and the accumulators are double-buffered in TMEM, so TMA prefetch, operand
// splitting, MMA issue and the epilogue of consecutive tiles overlap.
//
// Accuracy note: tcgen05 adds each 8-deep product into the fp32 accumulator with truncation, a bias that grows
// with the number of accumulation steps (measured 3e-5 relative at K = 3840 with a single accumulator).  The
// (2^-11 smaller) cross terms therefore get their own accumulator, and for long K the hi*hi products of
// consecutive k-steps rotate over three accumulators; the epilogue adds them in fp32 round-to-nearest.
// ACC_BUFS = 1 (long reductions with wide tiles: 4 accumulators x 128 columns fill TMEM) trades the epilogue / mainloop
// overlap — a few per cent of a K > 1024 mainloop — for 128-column MMAs, which halve the shared-memory reads of A per flop.
template <int BLOCK_N, int NUM_HI, int STAGES, int ACC_BUFS, bool H2, bool AT>
__global__ void __launch_bounds__(gemm_threads(BLOCK_N, AT), 1) k_gemm_tc(const __grid_constant__ CUtensorMap map_a,
                                                          const __grid_constant__ CUtensorMap map_a2, int kb_split,
                                                          const __grid_constant__ CUtensorMap map_b_hi,
                                                          const __grid_constant__ CUtensorMap map_b_lo,
                                                          const __grid_constant__ CUtensorMap map_c,
                                                          const __grid_constant__ CUtensorMap map_o2, float* __restrict__ C,
                                                          int64_t M, int N, int K, int ldc, Epilogue ep) {
  using L = SmemLayout<BLOCK_N, NUM_HI, STAGES, ACC_BUFS, H2, AT>;
  constexpr int kNumAcc = L::kNumAcc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  // stage s: [A hi (TMA lands raw A here) | A lo | B hi | B lo]
  const uint32_t bar_base = base + L::kTileBytes + L::kEpiBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto split_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tmem_full_bar = [&](int b) { return bar_base + 8u * (3 * STAGES + b); };
  auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (3 * STAGES + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + L::kTileBytes + L::kEpiBytes + 8u * (3 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  // tile indices fit 32 bits (the launcher checks): 64-bit divisions by num_n cost several hundred cycles per tile and role
  const uint32_t num_tiles = (uint32_t)(((M + BLOCK_M - 1) / BLOCK_M) * num_n);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_a2);
    tma_prefetch_desc(&map_b_hi);
    tma_prefetch_desc(&map_b_lo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), L::kSplitWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < ACC_BUFS; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), L::kEpiWarps);
    }
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
  // Programmatic dependent launch: everything above (descriptor prefetch, barrier init, TMEM allocation) touches no data of the
  // kernel in front of this one on the stream and may run while that kernel's last CTAs drain; every global access is below.
  // (A no-op when the launch does not carry the attribute.)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ---------------- TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const uint32_t m_blk = tile / (uint32_t)num_n;
        const int m0 = (int)m_blk * BLOCK_M, n0 = (int)(tile - m_blk * (uint32_t)num_n) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = (int)(it % STAGES);
          mbar_wait(empty_bar(s), ((it / STAGES) & 1u) ^ 1u);
          const uint32_t st = base + (uint32_t)s * L::kStageBytes;
          mbar_expect_tx(full_bar(s), kABoxBytes + 2 * L::kBBoxBytes);
          // k-blocks [0, kb_split) come from A, the rest from a second row-major operand (K-concatenated inputs that live
          // in different tensors: [A | A2] * [W1 | W2]^T without materialising the concatenation)
          if (kb < kb_split) tma_load_2d(st, &map_a, full_bar(s), kb * BLOCK_K, m0);
          else tma_load_2d(st, &map_a2, full_bar(s), (kb - kb_split) * BLOCK_K, m0);
          tma_load_2d(st + L::kBOffset, &map_b_hi, full_bar(s), kb * BLOCK_K, n0);
          tma_load_2d(st + L::kBOffset + L::kBBoxBytes, &map_b_lo, full_bar(s), kb * BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = H2 ? make_instr_desc_f16(BLOCK_M, BLOCK_N) : make_instr_desc(BLOCK_M, BLOCK_N);
      uint32_t it = 0, lt = 0;
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const uint32_t buf = lt % ACC_BUFS;
        mbar_wait(tmem_empty_bar(buf), ((lt / ACC_BUFS) & 1u) ^ 1u);  // epilogue has drained this accumulator set
        tcgen05_fence_after();
        const uint32_t acc0 = tmem_base + buf * (kNumAcc * BLOCK_N);
        const uint32_t acc_x = acc0 + (uint32_t)NUM_HI * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = (int)(it % STAGES);
          mbar_wait(split_bar(s), (it / STAGES) & 1u);
          tcgen05_fence_after();
          const uint32_t st = base + (uint32_t)s * L::kStageBytes;
          if constexpr (AT) {
            const uint32_t ta = tmem_base + L::kAccCols + (uint32_t)s * BLOCK_K;  // hi pairs at +0, lo pairs at +16
            const uint64_t b_hi = make_smem_desc_sw64(st + L::kBOffset), b_lo = make_smem_desc_sw64(st + L::kBOffset + L::kBBoxBytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
              const int ks = kb * (BLOCK_K / 16) + k;
              umma_f16_ta(acc_x, ta + 16u + 8u * k, b_hi + adv, idesc, ks != 0 ? 1u : 0u);
              umma_f16_ta(acc_x, ta + 8u * k, b_lo + adv, idesc, 1u);
              umma_f16_ta(acc0 + (uint32_t)(ks % NUM_HI) * BLOCK_N, ta + 8u * k, b_hi + adv, idesc, ks >= NUM_HI ? 1u : 0u);
            }
          } else if constexpr (H2) {
            const uint64_t a_hi = make_smem_desc_sw64(st + kABoxBytes), a_lo = make_smem_desc_sw64(st + kABoxBytes + L::kAHalfBytes);
            const uint64_t b_hi = make_smem_desc_sw64(st + L::kBOffset), b_lo = make_smem_desc_sw64(st + L::kBOffset + L::kBBoxBytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);  // 32 bytes per k-step inside the 64-byte swizzle row
              const int ks = kb * (BLOCK_K / 16) + k;
              umma_f16(acc_x, a_lo + adv, b_hi + adv, idesc, ks != 0 ? 1u : 0u);
              umma_f16(acc_x, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_f16(acc0 + (uint32_t)(ks % NUM_HI) * BLOCK_N, a_hi + adv, b_hi + adv, idesc, ks >= NUM_HI ? 1u : 0u);
            }
          } else {
          const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + kABoxBytes);
          const uint64_t b_hi = make_smem_desc(st + 2 * kABoxBytes), b_lo = make_smem_desc(st + 2 * kABoxBytes + L::kBBoxBytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);  // 32 bytes per k-step inside the swizzle row
            const int ks = kb * (BLOCK_K / UMMA_K) + k;
            umma_tf32(acc_x, a_lo + adv, b_hi + adv, idesc, ks != 0 ? 1u : 0u);
            umma_tf32(acc_x, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_tf32(acc0 + (uint32_t)(ks % NUM_HI) * BLOCK_N, a_hi + adv, b_hi + adv, idesc, ks >= NUM_HI ? 1u : 0u);
          }
          }
          umma_commit(empty_bar(s));  // stage reusable once these MMAs have read it
        }
        umma_commit(tmem_full_bar(buf));  // this tile's accumulators are complete
      }
    }
  } else if (warp < 2 + L::kSplitWarps) {
    // ---------------- splitters (warps 2..5)
    const int t = threadIdx.x - 64;  // 0..127
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = (int)(it % STAGES);
        mbar_wait(full_bar(s), (it / STAGES) & 1u);
        if constexpr (AT) {
          // thread = one row of the raw box (TMEM lane 32 (warp % 4) + lane): its 32 floats -> 16 columns of hi pairs and 16
          // of 2^11-scaled lo pairs in this stage's TMEM slot.  (The MMAs of the k-block that last used the slot are complete:
          // the producer waited for their commit before it refilled the stage.)
          constexpr int kCols = 32 / (L::kSplitWarps / 4);  // columns of the k-block per splitter thread
          const int q = warp & 3, r = 32 * q + lane, kh = (warp - 2) >> 2;
          const uint8_t* rawrow = base_ptr + (size_t)s * L::kStageBytes + (size_t)r * 128;
          uint32_t hw[kCols / 2], lw[kCols / 2];
#pragma unroll
          for (int j = 0; j < kCols / 4; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(rawrow + ((((kCols / 4) * kh + j) ^ (r & 7)) << 4));
            const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn((v.x - f01.x) * kLoScale, (v.y - f01.y) * kLoScale);
            const __half2 l23 = __floats2half2_rn((v.z - f23.x) * kLoScale, (v.w - f23.y) * kLoScale);
            hw[2 * j] = *reinterpret_cast<const uint32_t*>(&h01);
            hw[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h23);
            lw[2 * j] = *reinterpret_cast<const uint32_t*>(&l01);
            lw[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&l23);
          }
          const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + L::kAccCols + (uint32_t)s * BLOCK_K + (uint32_t)(kCols / 2) * kh;
          if constexpr (kCols == 32) {
            tmem_st_32x32b_x16(ta, hw);
            tmem_st_32x32b_x16(ta + 16u, lw);
          } else {
            tmem_st_32x32b_x8(ta, hw);
            tmem_st_32x32b_x8(ta + 16u, lw);
          }
          tmem_st_wait();
          tcgen05_fence_before();
        } else if constexpr (H2) {
          // raw fp32 box (128 rows x 128 B, SWIZZLE_128B as TMA wrote it) -> fp16 hi / lo boxes (128 rows x 64 B, SWIZZLE_64B).
          // float4 f of the raw box = row f / 8, physical 16-byte chunk f % 8 = logical chunk (f % 8) ^ (row % 8), i.e.
          // k = 4 c .. 4 c + 3; its four halves land in logical 16-byte chunk c / 2 (physical (c / 2) ^ ((row / 2) % 4)) at
          // byte (c % 2) * 8.  Consecutive threads read consecutive float4 and write into one 64-byte row: conflict-free.
          const float4* raw = reinterpret_cast<const float4*>(base_ptr + (size_t)s * L::kStageBytes);
          uint8_t* hi8 = base_ptr + (size_t)s * L::kStageBytes + kABoxBytes;
          uint8_t* lo8 = hi8 + L::kAHalfBytes;
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = raw[t + 128 * j];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int f = t + 128 * j;
            const int r = f >> 3, c = (f & 7) ^ (r & 7);
            const uint32_t o = (uint32_t)r * 64u + (uint32_t)(((c >> 1) ^ ((r >> 1) & 3)) << 4) + (uint32_t)((c & 1) << 3);
            __half h0, h1, h2, h3, l0, l1, l2, l3;
            split_h2(v[j].x, h0, l0);
            split_h2(v[j].y, h1, l1);
            split_h2(v[j].z, h2, l2);
            split_h2(v[j].w, h3, l3);
            const __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3), la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
            *reinterpret_cast<uint2*>(hi8 + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&ha), *reinterpret_cast<const uint32_t*>(&hb));
            *reinterpret_cast<uint2*>(lo8 + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&la), *reinterpret_cast<const uint32_t*>(&lb));
          }
        } else {
        float4* hi = reinterpret_cast<float4*>(base_ptr + (size_t)s * L::kStageBytes);
        float4* lo = reinterpret_cast<float4*>(base_ptr + (size_t)s * L::kStageBytes + kABoxBytes);
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
        }
        if constexpr (!AT) fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(s));
      }
    }
  } else {
    // ---------------- epilogue (warps 6..): warp w owns TMEM lanes 32*(w%4) .. +31; thread = one output row.  With
    // 8 epilogue warps, the two warps of a lane quarter take alternate 32-column chunks of the tile.
    const int q = warp & 3;
    const int ew = warp - (2 + L::kSplitWarps);    // 0 .. kEpiWarps-1
    const int half = ew >> 2;                      // which chunk parity this warp handles
    constexpr int kChunkStep = 32 * (L::kEpiWarps / 4);
    uint8_t* stg_ptr = base_ptr + L::kTileBytes + ew * 4096;  // 1024-byte aligned (SWIZZLE_128B)
    const uint32_t stg = base + L::kTileBytes + (uint32_t)ew * 4096u;
    const bool vec_ok = ep.vec_ok != 0;
    const bool tma_out = ep.tma_store != 0;
    if (tma_out && lane == 0) {
      tma_prefetch_desc(&map_c);
      if (ep.out2) tma_prefetch_desc(&map_o2);
    }
    // The side inputs of the epilogue (residual, addend, post-residual rows) are requested into L2 one tile ahead:
    // their DRAM latency would otherwise sit in the epilogue's critical path, once per 32-column chunk.
    const bool side_inputs = ep.residual || ep.out2 || ep.post_res;
    auto prefetch_side = [&](uint32_t tile_p) {
      const uint32_t mb = tile_p / (uint32_t)num_n;
      const int64_t mp = (int64_t)mb * BLOCK_M + 32 * q + lane;
      const int np0 = (int)(tile_p - mb * (uint32_t)num_n) * BLOCK_N;
      if (mp >= M) return;
      for (int c0 = 32 * half; c0 < BLOCK_N && np0 + c0 < N; c0 += kChunkStep) {
        if (ep.residual) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.residual + mp * (int64_t)ep.ld_res + np0 + c0));
        if (ep.out2) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.addend + mp * (int64_t)ep.ld_add + np0 + c0));
        if (ep.post_res) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.post_res + mp * (int64_t)ep.ld_post + np0 + c0));
      }
    };
    if (side_inputs && blockIdx.x < num_tiles) prefetch_side(blockIdx.x);
    uint32_t lt = 0;
    constexpr int kChunksPerWarp = BLOCK_N / kChunkStep;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const uint32_t m_blk = tile / (uint32_t)num_n;
      const int row0 = (int)m_blk * BLOCK_M + 32 * q;
      const int64_t m = (int64_t)row0 + lane;
      const int n0 = (int)(tile - m_blk * (uint32_t)num_n) * BLOCK_N;
      const uint32_t buf = lt % ACC_BUFS;
      const float rs = (ep.row_scale && m < M) ? ep.row_scale[m] : 1.0f;
      // column scale / shift of this warp's chunks: lane j fetches column j's pair now (one coalesced load, in flight while
      // the accumulators are awaited); the values reach the other lanes by shuffle.  (Eight broadcast float4 loads per
      // chunk, issued after the TMEM read, were the longest stall of the epilogue: with 227 KB of the SM given to shared
      // memory they miss L1.)
      float cs_pre[kChunksPerWarp], cb_pre[kChunksPerWarp];
#pragma unroll
      for (int ci = 0; ci < kChunksPerWarp; ++ci) {
        const int n = n0 + 32 * half + kChunkStep * ci + lane;
        cs_pre[ci] = (ep.col_scale && n < N) ? __ldg(ep.col_scale + n) : 1.0f;
        cb_pre[ci] = (ep.col_shift && n < N) ? __ldg(ep.col_shift + n) : 0.0f;
      }
      if (side_inputs && tile + gridDim.x < num_tiles) prefetch_side(tile + gridDim.x);
      mbar_wait(tmem_full_bar(buf), (lt / ACC_BUFS) & 1u);
      tcgen05_fence_after();
      const uint32_t acc0 = tmem_base + buf * (kNumAcc * BLOCK_N) + ((uint32_t)(32 * q) << 16);
      float* __restrict__ crow = C + m * (int64_t)ldc;
      const float* __restrict__ rrow = ep.residual ? ep.residual + m * (int64_t)ep.ld_res : nullptr;
      float* __restrict__ orow = ep.out2 ? ep.out2 + m * (int64_t)ep.ld2 : nullptr;
      const float* __restrict__ arow = ep.out2 ? ep.addend + m * (int64_t)ep.ld_add : nullptr;
      const float* __restrict__ prow = ep.post_res ? ep.post_res + m * (int64_t)ep.ld_post : nullptr;
#pragma unroll
      for (int ci = 0; ci < kChunksPerWarp; ++ci) {
        const int c0 = 32 * half + kChunkStep * ci;
        float sum[32];
        {
          // the accumulators of this chunk: issue the TMEM loads pairwise, one wait per pair
          uint32_t r[32], r2[32];
          tmem_ld_32x32b_x32(acc0 + (uint32_t)c0, r);
          tmem_ld_32x32b_x32(acc0 + (uint32_t)(BLOCK_N + c0), r2);
          tmem_ld_wait();
          // the cross-term accumulator is the LAST one; the fp16 split keeps it scaled by 2^11
          constexpr float kX = H2 ? kLoUnscale : 1.0f;
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]) * (kNumAcc == 2 ? kX : 1.0f);
          if constexpr (kNumAcc == 3) {
            // two hi*hi accumulators (summed above, unscaled) + the cross terms
            tmem_ld_32x32b_x32(acc0 + (uint32_t)(2 * BLOCK_N + c0), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r[j]) * kX;
          }
          if constexpr (kNumAcc == 4) {
            tmem_ld_32x32b_x32(acc0 + (uint32_t)(2 * BLOCK_N + c0), r);
            tmem_ld_32x32b_x32(acc0 + (uint32_t)(3 * BLOCK_N + c0), r2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r[j]) + __uint_as_float(r2[j]) * kX;
          }
        }
        // row scale, column scale / shift (same addresses in every lane: broadcast loads, L1 resident)
        if (ep.row_scale) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] *= rs;
        }
        if (ep.col_scale) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] *= __shfl_sync(0xffffffffu, cs_pre[ci], j);
        }
        if (ep.col_shift) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] += __shfl_sync(0xffffffffu, cb_pre[ci], j);
        }
        if (c0 + kChunkStep >= BLOCK_N) {
          // this warp's TMEM reads of the tile are done: hand the accumulator set back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar(buf));
        }
        if (tma_out) {
          // values -> swizzled 32 x 32 tile in shared memory -> one TMA store per warp (clipped at M and N).
          // Every run-time condition is tested once per chunk, outside the loops over the 32 columns, and the loads of a
          // side input are issued together: a branchy per-column body made the epilogue the slowest role of short-K layers.
          const bool row_ok = m < M;
          const bool full_chunk = n0 + c0 + 32 <= N;
          const int nc = n0 + c0;
          // sum[] += side[row, nc .. nc + 31]
          auto add_row = [&](const float* __restrict__ side_row) {
            if (row_ok && full_chunk) {
              float4 t4[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) t4[j] = *reinterpret_cast<const float4*>(side_row + nc + 4 * j);
#pragma unroll
              for (int j = 0; j < 8; ++j) { sum[4 * j] += t4[j].x; sum[4 * j + 1] += t4[j].y; sum[4 * j + 2] += t4[j].z; sum[4 * j + 3] += t4[j].w; }
            } else if (row_ok) {
#pragma unroll
              for (int e = 0; e < 32; ++e) if (nc + e < N) sum[e] += side_row[nc + e];
            }
          };
          auto activate_all = [&](int act) {
            if (act == 1) {
#pragma unroll
              for (int e = 0; e < 32; ++e) sum[e] = fmaxf(sum[e], 0.f);
            } else if (act == 2) {
#pragma unroll
              for (int e = 0; e < 32; ++e) sum[e] = sum[e] > 0.f ? sum[e] : sum[e] * ep.slope;
            }
          };
          auto stage_and_store = [&](const CUtensorMap* map) {
            if (lane == 0) tma_store_wait_read();  // the previous store has finished reading this warp's tile
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(stg_ptr + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(sum[4 * j], sum[4 * j + 1], sum[4 * j + 2], sum[4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) tma_store_2d(map, stg, nc, row0);
          };
          if (rrow) add_row(rrow);
          activate_all(ep.act);
          if (prow) {
            add_row(prow);
            activate_all(ep.post_act);
          }
          stage_and_store(&map_c);
          if (orow) {
            add_row(arow);
            stage_and_store(&map_o2);
          }
        } else if (m < M) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = n0 + c0 + j;
            if (n >= N) break;
            const bool full4 = vec_ok && (n + 3 < N);
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = sum[j + e];
            if (rrow) {
              if (full4) {
                const float4 rr = *reinterpret_cast<const float4*>(rrow + n);
                v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (n + e < N) v[e] += rrow[n + e];
              }
            }
            if (ep.act == 1) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
            } else if (ep.act == 2) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * ep.slope;
            }
            if (prow) {
              if (full4) {
                const float4 pp = *reinterpret_cast<const float4*>(prow + n);
                v[0] += pp.x; v[1] += pp.y; v[2] += pp.z; v[3] += pp.w;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (n + e < N) v[e] += prow[n + e];
              }
              if (ep.post_act == 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
              } else if (ep.post_act == 2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * ep.slope;
              }
            }
            if (full4) {
              *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) if (n + e < N) crow[n + e] = v[e];
            }
            if (orow) {
              if (full4) {
                const float4 aa = *reinterpret_cast<const float4*>(arow + n);
                *reinterpret_cast<float4*>(orow + n) = make_float4(v[0] + aa.x, v[1] + aa.y, v[2] + aa.z, v[3] + aa.w);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (n + e < N) orow[n + e] = v[e] + arow[n + e];
              }
            }
          }
        }
      }
      __syncwarp();
    }
    if (tma_out && lane == 0) tma_store_wait_all();  // every store of this warp has landed before the CTA exits
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(L::kTmemCols));
  }
}

// Bt_hi / Bt_lo [N, ldb] (K-major) from W: either [K, N] row-major (transpose = 1, the KPConv weights
// flattened to [K*c_in, c_out]) or [N, K] row-major (transpose = 0, an nn.Linear weight).
__global__ void __launch_bounds__(256) k_split_weights(const float* __restrict__ w, int k_dim, int n_dim, int ldb, int transpose,
                                                       float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t total = (int64_t)n_dim * ldb;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / ldb), k = (int)(i - (int64_t)n * ldb);
    float v = 0.f;
    if (k < k_dim) v = transpose ? w[(int64_t)k * n_dim + n] : w[(int64_t)n * k_dim + k];
    float h, l;
    split_tf32(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// The same for the fp16 split: hi / lo as __half [N, ldb] (ldb a multiple of 8 halves).
__global__ void __launch_bounds__(256) k_split_weights_h2(const float* __restrict__ w, int k_dim, int n_dim, int ldb, int transpose,
                                                          __half* __restrict__ hi, __half* __restrict__ lo) {
  const int64_t total = (int64_t)n_dim * ldb;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / ldb), k = (int)(i - (int64_t)n * ldb);
    float v = 0.f;
    if (k < k_dim) v = transpose ? w[(int64_t)k * n_dim + n] : w[(int64_t)n * k_dim + k];
    __half h, l;
    split_h2(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch ld (floats), box = [box_rows, 32 floats], SWIZZLE_128B.
int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int cols, int64_t ld, int box_rows) {
  if (ptr == nullptr) { memset(map, 0, sizeof(*map)); return KPREG_OK; }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return KPREG_E_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled", cudaErrorInvalidValue);
    return KPREG_E_CUDA;
  }
  return KPREG_OK;
}

// 2-D fp16 tensor [rows, cols] with row pitch ld (halves), box = [box_rows, 32 halves], SWIZZLE_64B (the fp16-split weights)
int make_map_h(CUtensorMap* map, const __half* ptr, int64_t rows, int cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return KPREG_E_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled", cudaErrorInvalidValue);
    return KPREG_E_CUDA;
  }
  return KPREG_OK;
}

// Operand format of the tensor-core GEMM for this process: fp16 split (default) or 3xTF32 (KPREG_GEMM_TF32=1).  The
// pre-split weight buffers are in the matching format, so the choice is made once.
bool gemm_h2() {
  static const bool h2 = [] { const char* e = getenv("KPREG_GEMM_TF32"); return !(e && e[0] == '1'); }();
  return h2;
}

template <int BLOCK_N, int NUM_HI, int STAGES, int ACC_BUFS, bool H2, bool AT = false>
int launch_tile_config(const CUtensorMap& ma, const CUtensorMap& ma2, int kb_split, const CUtensorMap& mbh, const CUtensorMap& mbl, const CUtensorMap& mc,
                       const CUtensorMap& mo2, float* C, int64_t M, int N, int K, int ldc, const Epilogue& ep, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, NUM_HI, STAGES, ACC_BUFS, H2, AT>;
  static PerDeviceOnce once;  // (the attribute is per device: a second GPU in the same process needs its own call)
  const int rc_cfg = once.run([]() -> int {
    KP_CUDA_TRY(cudaFuncSetAttribute(k_gemm_tc<BLOCK_N, NUM_HI, STAGES, ACC_BUFS, H2, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kTotal));
    return KPREG_OK;
  });
  if (rc_cfg) return rc_cfg;
  const int64_t tiles = (int64_t)ceil_div(M, BLOCK_M) * ceil_div(N, BLOCK_N);
  if (tiles >= ((int64_t)1 << 31) || M >= ((int64_t)1 << 31)) return KPREG_E_RANGE;  // 32-bit tile / row arithmetic in the kernel
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  static const bool pdl = [] { const char* e = getenv("KPREG_GEMM_PDL"); return e && e[0] == '1'; }();
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3((unsigned)gemm_threads(BLOCK_N, AT));
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_i = N, k_i = K, ldc_i = ldc;
    KP_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_gemm_tc<BLOCK_N, NUM_HI, STAGES, ACC_BUFS, H2, AT>, ma, ma2, kb_split, mbh, mbl, mc, mo2, C, M, n_i, k_i,
                                   ldc_i, ep));
    return KPREG_OK;
  }
  k_gemm_tc<BLOCK_N, NUM_HI, STAGES, ACC_BUFS, H2, AT><<<grid, gemm_threads(BLOCK_N, AT), L::kTotal, stream>>>(ma, ma2, kb_split, mbh, mbl, mc, mo2, C, M, N, K, ldc, ep);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

int ldb_for(int k_dim) { return (k_dim + 3) / 4 * 4; }
int ldb_h_for(int k_dim) { return (k_dim + 7) / 8 * 8; }  // halves: rows stay 16-byte aligned
int npad_for(int n_dim) { return (n_dim + 15) / 16 * 16; }

}  // namespace

size_t kpconv_gemm_tc_weight_bytes(int kd, int n) {
  // hi and lo copies of Bt [npad, ldb]
  return align_up((size_t)2 * (size_t)npad_for(n) * (size_t)ldb_for(kd) * sizeof(float) + 512, 256);
}

// Tile width (output columns per CTA tile) the tensor-core GEMM uses for a [*, kd] x [kd, n] product.
int gemm_tc_block_n(int kd, int n) {
  const bool h2 = gemm_h2();
  const int num_hi = kd > 1024 ? 3 : 1;
  // wide outputs of long reductions: 128-column tiles with single-buffered accumulators (4 x 128 TMEM columns)
  static const bool no_wide3 = [] { const char* e = getenv("KPREG_GEMM_NO_WIDE3"); return e && e[0] == '1'; }();
  const bool wide3 = num_hi == 3 && n > 64 && !no_wide3;
  // fp16 split, K <= 1024, wide outputs: 128 x 256 tiles (two single-buffered 256-column accumulators fill TMEM) — the
  // mainloop is bound by shared-memory bandwidth and a 256-column MMA reads A once for twice the output.  Outputs of
  // 129 .. 255 columns (res2net's w = 224 chain layers) take ONE 256-column tile rather than a 128 + remainder pair: A is
  // read once, and a single column tile is what lets the chained layers work in place (res2net.py).
  static const bool no_n256 = [] { const char* e = getenv("KPREG_GEMM_NO_N256"); return e && e[0] == '1'; }();
  const bool n256 = h2 && num_hi == 1 && n > 128 && kd >= 192 && !no_n256;
  return n256 ? 256 : (n <= 32 ? 32 : ((n <= 64 || (num_hi == 3 && !wide3)) ? 64 : 128));
}

bool gemm_tc_supported(int64_t m, int kd, int n, int lda, const void* a) {
  return m > 0 && kd >= 4 && (lda % 4) == 0 && n >= 8 && (reinterpret_cast<uintptr_t>(a) % 16) == 0;
}

// Split the weights: w is [kd, n] (transpose = 1) or [n, kd] (transpose = 0).  w_split holds hi then lo.
int kpconv_gemm_tc_prepare_weights(const float* weights, int kd, int n, int transpose, float* w_split, cudaStream_t stream) {
  if (gemm_h2()) {
    const int ldb = ldb_h_for(kd);
    __half* hi = reinterpret_cast<__half*>(w_split);
    __half* lo = hi + (size_t)npad_for(n) * ldb;
    const int64_t total = (int64_t)n * ldb;
    int blocks = ceil_div(total, 256);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    k_split_weights_h2<<<blocks, 256, 0, stream>>>(weights, kd, n, ldb, transpose, hi, lo);
    KP_LAUNCH_CHECK();
    return KPREG_OK;
  }
  const int ldb = ldb_for(kd);
  float* hi = w_split;
  float* lo = w_split + (size_t)npad_for(n) * ldb;
  const int64_t total = (int64_t)n * ldb;
  int blocks = ceil_div(total, 256);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  k_split_weights<<<blocks, 256, 0, stream>>>(weights, kd, n, ldb, transpose, hi, lo);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

int launch_gemm_tc_pair(const float* a, int lda, int k1, const float* a2, int lda2, const float* w_split, float* c, int ldc, int64_t m,
                        int kd, int n, const float* row_scale, const float* col_scale, const float* col_shift, const float* residual,
                        int ld_res, int act, float slope, float* out2, int ld2, const float* addend, int ld_add, const float* post_res,
                        int ld_post, int post_act, cudaStream_t stream);

// C[m, n] (row pitch ldc) = epilogue(A[m, kd] (row pitch lda) * Bt^T) with pre-split weights.
int launch_gemm_tc(const float* a, int lda, const float* w_split, float* c, int ldc, int64_t m, int kd, int n,
                   const float* row_scale, const float* col_scale, const float* col_shift, const float* residual, int ld_res,
                   int act, float slope, float* out2, int ld2, const float* addend, int ld_add, const float* post_res, int ld_post,
                   int post_act, cudaStream_t stream) {
  return launch_gemm_tc_pair(a, lda, kd, nullptr, 0, w_split, c, ldc, m, kd, n, row_scale, col_scale, col_shift, residual, ld_res, act, slope,
                             out2, ld2, addend, ld_add, post_res, ld_post, post_act, stream);
}

// The same with the reduction index split over two row-major operands: columns [0, k1) of the product's K come from a
// (k1 valid columns, zero-filled up to the next multiple of 32), columns [pad32(k1), kd) from a2.  a2 == nullptr: plain GEMM.
int launch_gemm_tc_pair(const float* a, int lda, int k1, const float* a2, int lda2, const float* w_split, float* c, int ldc, int64_t m,
                        int kd, int n, const float* row_scale, const float* col_scale, const float* col_shift, const float* residual,
                        int ld_res, int act, float slope, float* out2, int ld2, const float* addend, int ld_add, const float* post_res,
                        int ld_post, int post_act, cudaStream_t stream) {
  if (!gemm_tc_supported(m, kd, n, lda, a)) return KPREG_E_INVALID;
  const int k1_pad = (k1 + BLOCK_K - 1) / BLOCK_K * BLOCK_K;
  if (a2 && (k1 < 4 || k1_pad >= kd || (lda2 % 4) != 0 || (reinterpret_cast<uintptr_t>(a2) % 16) != 0)) return KPREG_E_INVALID;
  const bool h2 = gemm_h2();
  const int ldb = ldb_for(kd);
  const float* hi = w_split;
  const float* lo = w_split + (size_t)npad_for(n) * ldb;
  // long reductions rotate the hi*hi products over three accumulators (see the accuracy note above)
  const int num_hi = kd > 1024 ? 3 : 1;
  const int block_n = gemm_tc_block_n(kd, n);
  CUtensorMap ma, ma2, mbh, mbl;
  int rc = make_map(&ma, a, m, a2 ? k1 : kd, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_map(&ma2, a2 ? a2 : a, m, a2 ? kd - k1_pad : kd, a2 ? lda2 : lda, BLOCK_M);
  if (rc) return rc;
  const int kb_split = a2 ? k1_pad / BLOCK_K : (1 << 30);
  if (h2) {
    const int ldh = ldb_h_for(kd);
    const __half* hh = reinterpret_cast<const __half*>(w_split);
    rc = make_map_h(&mbh, hh, n, kd, ldh, block_n);
    if (rc) return rc;
    rc = make_map_h(&mbl, hh + (size_t)npad_for(n) * ldh, n, kd, ldh, block_n);
    if (rc) return rc;
  } else {
    rc = make_map(&mbh, hi, n, kd, ldb, block_n);
    if (rc) return rc;
    rc = make_map(&mbl, lo, n, kd, ldb, block_n);
    if (rc) return rc;
  }
  auto aligned = [](const void* p, int ld) { return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0); };
  const int vec_ok = aligned(c, ldc) && aligned(residual, ld_res) && aligned(out2, ld2) && aligned(addend, ld_add) && aligned(post_res, ld_post);
  // outputs leave through TMA when every pointer / pitch is 16-byte aligned (the store clips at M and N itself)
  // (a staged tile flushed with coalesced STG.128 instead of the TMA store was measured 20-40 % slower on short-K layers)
  const int tma_store = vec_ok;
  CUtensorMap mc, mo2;
  rc = make_map(&mc, tma_store ? c : nullptr, m, n, ldc, 32);
  if (rc) return rc;
  rc = make_map(&mo2, (tma_store && out2) ? out2 : nullptr, m, n, ld2, 32);
  if (rc) return rc;
  Epilogue ep{row_scale, col_scale, col_shift, residual, ld_res, act, slope, out2, addend, ld2, ld_add, vec_ok, tma_store,
              post_res, ld_post, post_act};
  // Long reductions with 64 / 128-column tiles: the split A operand goes through tensor memory (the mainloop is bound by
  // shared-memory bandwidth: A's hi / lo boxes cost 16 KB of writes and 24 KB of MMA reads per k-block).  TMEM then holds the
  // accumulators + 4 x 32 columns of A, so long K rotates hi*hi over two accumulators instead of three.
  static const bool no_at = [] { const char* e = getenv("KPREG_GEMM_NO_AT"); return e && e[0] == '1'; }();
  if (h2 && !no_at && kd >= 512 && block_n >= 64 && block_n <= 128) {
    if (block_n == 64) {
      if (num_hi == 3) return launch_tile_config<64, 3, 4, 1, true, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
      return launch_tile_config<64, 1, 4, 2, true, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
    }
    if (num_hi == 3) return launch_tile_config<128, 2, 4, 1, true, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
    return launch_tile_config<128, 1, 4, 1, true, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
  }
  if (h2) {
    if (num_hi == 3) {
      if (block_n == 32) return launch_tile_config<32, 3, 4, 2, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
      if (block_n == 64) return launch_tile_config<64, 3, 4, 2, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
      return launch_tile_config<128, 3, 4, 1, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
    }
    if (block_n == 256) return launch_tile_config<256, 1, 3, 1, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
    if (block_n == 32) return launch_tile_config<32, 1, 4, 2, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
    if (block_n == 64) return launch_tile_config<64, 1, 4, 2, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
    return launch_tile_config<128, 1, 4, 2, true>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
  }
  if (num_hi == 3) {
    if (block_n == 32) return launch_tile_config<32, 3, 4, 2, false>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
    if (block_n == 64) return launch_tile_config<64, 3, 4, 2, false>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
    return launch_tile_config<128, 3, 3, 1, false>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
  }
  if (block_n == 32) return launch_tile_config<32, 1, 4, 2, false>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
  if (block_n == 64) return launch_tile_config<64, 1, 4, 2, false>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
  return launch_tile_config<128, 1, 3, 2, false>(ma, ma2, kb_split, mbh, mbl, mc, mo2, c, m, n, kd, ldc, ep, stream);
}

}  // namespace kpreg
