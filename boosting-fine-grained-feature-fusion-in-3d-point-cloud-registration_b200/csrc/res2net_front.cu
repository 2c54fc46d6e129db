// The front half of the fork's "fine-grained feature fusion" unit — conv1 + bn1 + relu, the split into groups and the
// hierarchical chain of per-group Linear + BN + ReLU layers (reference models/backbone_kpconv/res2net.py:125-150) — as ONE
// tcgen05 kernel per 128-row tile:
//
//     t_g = relu(x W1_g^T + b1_g)                     g = 0 .. G-1      (conv1, one group of `width` channels at a time)
//     y_0 = relu(t_0 Wc_0^T + bc_0),   y_g = relu((y_{g-1} + t_g) Wc_g^T + bc_g)   g = 1 .. G-2
//     z   = [y_0 | ... | y_{G-2} | t_{G-1} | x]       (x: the optional copy behind the concatenation for conv3's K-concatenated
//                                                     residual projection)
//
// Layer by layer this is a GEMM that writes t (G w floats per row), a chain kernel that reads t and writes z, i.e. 3 G w + c
// floats of HBM traffic per row; here x is read once and z written once (G w + 2 c floats per row): t and the running
// activation never leave the SM.
//
//   last warp (a lane)  producer: TMA load of the x tile (SWIZZLE_128B boxes), the optional x -> z copy as a TMA store straight
//                       from that tile, and — when the weights do not fit shared memory (width > 32) — the per-group weight
//                       blobs through a two-stage ring of 1-D bulk copies
//   the warp before it  MMA issuer: conv1 of group g+1 is issued as soon as the workers have drained group g's accumulator, so it
//                       runs under the workers' conversion of group g; the chain MMA of group g follows when its operand is ready
//   warps 0-3 | 0-7     workers: thread = one row (TMEM lane) x 32 columns of a group.  Split the x tile into fp16 hi / lo
//                       operand boxes once per tile; per group: read conv1's accumulator (t_g), read the chain accumulator
//                       (y_{g-1}), store y_{g-1} through a swizzled staging tile + TMA store, form y_{g-1} + t_g, split it into
//                       the chain MMA's operand boxes.
// Precision: the fp16 hi / lo operand split of the GEMM kernel (tc_ptx.cuh; three kind::f16 products, cross terms in their own
// accumulator) — the same arithmetic the layer-by-layer path runs.
//
// Shared memory (width 28, c_in 32): weights of all groups resident (60 KB), two CTAs per SM; (width 56, c_in 64): weights
// streamed per group from L2 (32 KB per group and tile against 256 KB of HBM traffic per tile), one CTA per SM.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace kpreg {
namespace {

using namespace tc;

constexpr int kFrontRows = 128;
constexpr int kFrontMaxGroups = 8;

template <int WP, int K1>
struct FrontLayout {
  static constexpr int NKX = K1 / 32, NKA = WP / 32;
  // Worker warps: one per TMEM lane quarter and 32-column slice of a group.  (A 16-column split — eight warps at WP = 32 —
  // made every output store a box of 64-byte rows, and the TMA store path, not the chain, set the tile rate: +47 %.)
  static constexpr int CW = 32;                                      // columns of a group per worker thread
  static constexpr int kWorkers = 4 * (WP / CW);
  static constexpr int kThreads = 32 * (kWorkers + 2);               // + MMA warp + producer warp
  static constexpr uint32_t kBox = kFrontRows * 64;                  // one [128 x 32 halves] operand box (SWIZZLE_64B)
  static constexpr uint32_t kXRaw = NKX * 16384;                     // fp32 boxes [128 x 32 floats] as TMA delivers them
  static constexpr uint32_t kXSplit = 2 * NKX * kBox;                // hi boxes, then lo boxes
  static constexpr uint32_t kA = 2 * NKA * kBox;
  static constexpr uint32_t kW1 = 2 * NKX * WP * 64;                 // one group's conv1 weights: hi boxes [WP x 32 halves], then lo
  static constexpr uint32_t kWc = 2 * NKA * WP * 64;                 // one group's chain weights
  static constexpr uint32_t kBlob = kW1 + kWc;
  static constexpr int kWStages = WP == 32 ? 3 : 2;                  // ring of per-group weight blobs (streamed from L2)
  static constexpr uint32_t kW = kWStages * kBlob;
  static constexpr int kStgBufs = WP == 32 ? 2 : 1;                  // staging tiles per worker warp
  static constexpr uint32_t kStgTile = 32 * CW * 4;                  // one [32 rows x CW floats] tile
  static constexpr uint32_t kStg = kWorkers * kStgBufs * kStgTile;
  static constexpr uint32_t kShift = 2 * kFrontMaxGroups * WP * 4;   // b1 [8][WP], bc [8][WP]
  static constexpr uint32_t oXRaw = 0;
  static constexpr uint32_t oXSplit = oXRaw + kXRaw;
  static constexpr uint32_t oA = oXSplit + kXSplit;
  static constexpr uint32_t oW = oA + kA;
  static constexpr uint32_t oStg = oW + kW;
  static constexpr uint32_t oShift = oStg + kStg;
  static constexpr uint32_t oBar = oShift + kShift;
  static constexpr uint32_t kTotal = oBar + 256 + 1024;              // + slack for the 1024-byte alignment
  static constexpr int kCtasPerSm = kTotal <= 115712u ? 2 : 1;
  static constexpr uint32_t kTmemCols = 4 * WP;                      // conv1 hi*hi, conv1 cross, chain hi*hi, chain cross
  static constexpr size_t kPackBytes = (size_t)kFrontMaxGroups * kBlob + kShift;
  static_assert(WP == 32 || WP == 64, "group width padded to 32 or 64 columns");
  static_assert(kBlob % 1024 == 0 && kW1 % 1024 == 0 && oW % 1024 == 0 && oStg % 1024 == 0 && kStgTile % 1024 == 0,
                "swizzled boxes stay 1024-byte aligned");
  static_assert(kTotal <= 232448u, "shared memory budget of one CTA");
};

enum FrontBar { kXFull = 0, kXEmpty, kXsFull, kTFull, kTEmpty, kAFull, kCFull, kWFull0, kWEmpty0 = kWFull0 + 4, kNumBars = kWEmpty0 + 4 };

template <int WP, int K1>
__global__ void __launch_bounds__(FrontLayout<WP, K1>::kThreads, FrontLayout<WP, K1>::kCtasPerSm) k_res2net_front(const __grid_constant__ CUtensorMap map_x,
                                                                                  const __grid_constant__ CUtensorMap map_z,
                                                                                  const __grid_constant__ CUtensorMap map_xc,
                                                                                  const uint8_t* __restrict__ pack, int n_groups,
                                                                                  int64_t m_rows, int copy_x) {
  using L = FrontLayout<WP, K1>;
  constexpr int NKX = L::NKX, NKA = L::NKA, CW = L::CW, WS = L::kWStages, NW = L::kWorkers;
  constexpr int kMmaWarp = NW, kProdWarp = NW + 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  auto bar = [&](int b) { return base + L::oBar + 8u * (uint32_t)b; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + L::oBar + 8u * kNumBars);
  float* s_shift = reinterpret_cast<float*>(base_ptr + L::oShift);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_layers = n_groups - 1;
  const uint32_t num_tiles = (uint32_t)((m_rows + kFrontRows - 1) / kFrontRows);

  if (warp == kProdWarp && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_z);
    if (copy_x) tma_prefetch_desc(&map_xc);
    mbar_init(bar(kXFull), 1);
    mbar_init(bar(kXEmpty), NW);
    mbar_init(bar(kXsFull), NW);
    mbar_init(bar(kTFull), 1);
    mbar_init(bar(kTEmpty), NW);
    mbar_init(bar(kAFull), NW);
    mbar_init(bar(kCFull), 1);
    for (int i = 0; i < WS; ++i) {
      mbar_init(bar(kWFull0 + i), 1);
      mbar_init(bar(kWEmpty0 + i), 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(L::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    const float* g_shift = reinterpret_cast<const float*>(pack + (size_t)kFrontMaxGroups * L::kBlob);
    for (int i = threadIdx.x; i < (int)(L::kShift / 4); i += L::kThreads) s_shift[i] = __ldg(g_shift + i);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProdWarp) {
    // ---------------- producer
    if (lane == 0 && blockIdx.x < num_tiles) {
      auto load_x = [&](uint32_t tile) {
        mbar_expect_tx(bar(kXFull), L::kXRaw);
        for (int b = 0; b < NKX; ++b) tma_load_2d(base + L::oXRaw + (uint32_t)b * 16384u, &map_x, bar(kXFull), b * 32, (int)tile * kFrontRows);
      };
      // group g's weights into ring slot ws % WS (ws counts groups across this CTA's tiles)
      auto load_w = [&](uint32_t ws, int g) {
        const uint32_t s = ws % WS;
        mbar_wait(bar(kWEmpty0 + (int)s), ((ws / WS) & 1u) ^ 1u);
        mbar_expect_tx(bar(kWFull0 + (int)s), L::kBlob);
        bulk_load_1d(base + L::oW + s * L::kBlob, pack + (size_t)g * L::kBlob, L::kBlob, bar(kWFull0 + (int)s));
      };
      load_x(blockIdx.x);
      uint32_t lt = 0;
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        if (copy_x) {
          // the block input rides along behind the concatenation: stored straight from the tile TMA just delivered
          mbar_wait(bar(kXFull), lt & 1u);
          for (int b = 0; b < NKX; ++b) tma_store_2d(&map_xc, base + L::oXRaw + (uint32_t)b * 16384u, b * 32, (int)tile * kFrontRows);
          tma_store_wait_read();
        }
        const uint32_t ws0 = lt * (uint32_t)n_groups;
        const int first = n_groups < WS ? n_groups : WS;
        for (int g = 0; g < first; ++g) load_w(ws0 + (uint32_t)g, g);
        if (tile + gridDim.x < num_tiles) {
          mbar_wait(bar(kXEmpty), lt & 1u);  // the workers have split this tile's raw box: the next tile's load may land
          load_x(tile + gridDim.x);
        }
        for (int g = first; g < n_groups; ++g) load_w(ws0 + (uint32_t)g, g);
      }
      tma_store_wait_all();
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer
    if (lane == 0 && blockIdx.x < num_tiles) {
      const uint32_t idesc = make_instr_desc_f16(kFrontRows, WP);
      const uint32_t t_hh = tmem_base, t_x = tmem_base + WP, c_hh = tmem_base + 2 * WP, c_x = tmem_base + 3 * WP;
      // operand boxes: hi boxes [0, nk), lo boxes [nk, 2 nk); two k-steps of 16 per 32-half box
      auto issue = [&](uint32_t a_base, uint32_t b_base, int nk, uint32_t d_hh, uint32_t d_x) {
        for (int kc = 0; kc < nk; ++kc) {
          const uint64_t a_hi = make_smem_desc_sw64(a_base + (uint32_t)kc * L::kBox), a_lo = make_smem_desc_sw64(a_base + (uint32_t)(nk + kc) * L::kBox);
          const uint64_t b_hi = make_smem_desc_sw64(b_base + (uint32_t)kc * (WP * 64)), b_lo = make_smem_desc_sw64(b_base + (uint32_t)(nk + kc) * (WP * 64));
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);  // 32 bytes per k-step inside the 64-byte swizzle row
            const uint32_t acc = (kc | k) != 0 ? 1u : 0u;
            umma_f16(d_x, a_lo + adv, b_hi + adv, idesc, acc);
            umma_f16(d_x, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_f16(d_hh, a_hi + adv, b_hi + adv, idesc, acc);
          }
        }
      };
      uint32_t lt = 0;
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const uint32_t ts0 = lt * (uint32_t)n_groups, tc0 = lt * (uint32_t)n_layers;
        mbar_wait(bar(kXsFull), lt & 1u);
        tcgen05_fence_after();
        auto w_base = [&](int g) { return base + L::oW + ((ts0 + (uint32_t)g) % WS) * L::kBlob; };
        auto conv1 = [&](int g) {
          const uint32_t ts = ts0 + (uint32_t)g;
          mbar_wait(bar(kTEmpty), (ts & 1u) ^ 1u);  // the workers have read the previous group's accumulator
          mbar_wait(bar(kWFull0 + (int)(ts % WS)), (ts / WS) & 1u);
          tcgen05_fence_after();
          issue(base + L::oXSplit, w_base(g), NKX, t_hh, t_x);
          umma_commit(bar(kTFull));
          if (g == n_groups - 1) umma_commit(bar(kWEmpty0 + (int)(ts % WS)));  // the last group has no chain layer
        };
        conv1(0);
        for (int g = 0; g < n_layers; ++g) {
          conv1(g + 1);
          mbar_wait(bar(kAFull), (tc0 + (uint32_t)g) & 1u);
          tcgen05_fence_after();
          issue(base + L::oA, w_base(g) + L::kW1, NKA, c_hh, c_x);
          umma_commit(bar(kCFull));
          umma_commit(bar(kWEmpty0 + (int)((ts0 + (uint32_t)g) % WS)));
        }
      }
    }
  } else {
    // ---------------- workers: warp w owns TMEM lanes 32 (w % 4) .. + 31 and columns [CW (w / 4), + CW) of every group
    const int q = warp & 3, hcol = warp >> 2;
    const int c0 = hcol * CW;
    const int row = 32 * q + lane;  // row of the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;
    const uint32_t stg_off = L::oStg + (uint32_t)warp * (L::kStgBufs * L::kStgTile);
    uint32_t n_stores = 0;  // staging tiles alternate when the warp has two
    const float* s_b1 = s_shift;
    const float* s_bc = s_shift + kFrontMaxGroups * WP;
    const int tid = threadIdx.x;  // 0 .. 32 NW - 1
    constexpr int kXPer = 1024 / (32 * NW);  // float4 of a raw box per worker thread

    // accumulator pair (hi*hi at col, cross terms — still scaled by 2^11 — at col + WP) -> v[0 .. CW)
    auto load_acc = [&](uint32_t col, float (&v)[CW]) {
      uint32_t r[CW], r2[CW];
      if constexpr (CW == 32) {
        tmem_ld_32x32b_x32(tmem_base + lane_off + col, r);
        tmem_ld_32x32b_x32(tmem_base + lane_off + col + WP, r2);
      } else {
        tmem_ld_32x32b_x16(tmem_base + lane_off + col, r);
        tmem_ld_32x32b_x16(tmem_base + lane_off + col + WP, r2);
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]) * kLoUnscale;
    };
    // v -> this warp's swizzled staging tile -> one TMA store into group `grp` of z (clipped at the group's width and at M)
    auto store_tile = [&](const float (&v)[CW], int grp, int m0) {
      const uint32_t off = stg_off + (L::kStgBufs == 2 ? (n_stores & 1u) * L::kStgTile : 0u);
      ++n_stores;
      // the store that last used this tile has finished reading it
      if (lane == 0) {
        if constexpr (L::kStgBufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else tma_store_wait_read();
      }
      __syncwarp();
      uint8_t* stg_ptr = base_ptr + off;
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) {
        const int pj = CW == 32 ? (j ^ (lane & 7)) : (j ^ ((lane >> 1) & 3));
        *reinterpret_cast<float4*>(stg_ptr + lane * (CW * 4) + (pj << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) tma_store_3d(&map_z, base + off, c0, grp, m0 + 32 * q);
    };

    uint32_t lt = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int m0 = (int)tile * kFrontRows;
      // ---- x tile: raw fp32 boxes (SWIZZLE_128B as TMA wrote them) -> fp16 hi / lo boxes (SWIZZLE_64B); see k_gemm_tc's splitter
      mbar_wait(bar(kXFull), lt & 1u);
#pragma unroll
      for (int b = 0; b < NKX; ++b) {
        const float4* raw = reinterpret_cast<const float4*>(base_ptr + L::oXRaw + b * 16384);
        uint8_t* hi8 = base_ptr + L::oXSplit + (uint32_t)b * L::kBox;
        uint8_t* lo8 = hi8 + NKX * L::kBox;
        float4 v[kXPer];
#pragma unroll
        for (int j = 0; j < kXPer; ++j) v[j] = raw[tid + 32 * NW * j];
#pragma unroll
        for (int j = 0; j < kXPer; ++j) {
          const int f = tid + 32 * NW * j;
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
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(kXsFull));
        mbar_arrive(bar(kXEmpty));
      }

      float y[CW];
      for (int g = 0; g < n_groups; ++g) {
        const uint32_t ts = lt * (uint32_t)n_groups + (uint32_t)g;
        mbar_wait(bar(kTFull), ts & 1u);
        tcgen05_fence_after();
        float t[CW];
        load_acc(0u + (uint32_t)c0, t);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kTEmpty));
#pragma unroll
        for (int j = 0; j < CW; ++j) t[j] = fmaxf(t[j] + s_b1[g * WP + c0 + j], 0.f);
        if (g >= 1) {
          const uint32_t tc = lt * (uint32_t)n_layers + (uint32_t)(g - 1);
          mbar_wait(bar(kCFull), tc & 1u);
          tcgen05_fence_after();
          load_acc((uint32_t)(2 * WP + c0), y);
#pragma unroll
          for (int j = 0; j < CW; ++j) y[j] = fmaxf(y[j] + s_bc[(g - 1) * WP + c0 + j], 0.f);
          if (g < n_groups - 1) {
#pragma unroll
            for (int j = 0; j < CW; ++j) t[j] += y[j];
          }
        }
        if (g < n_groups - 1) {
          // operand of chain layer g: 8 consecutive k per 16-byte chunk of row `row`
#pragma unroll
          for (int cc = 0; cc < CW / 8; ++cc) {
            const int k = c0 + 8 * cc;
            const uint32_t o = (uint32_t)(k >> 5) * L::kBox + (uint32_t)row * 64u + (uint32_t)(((((k & 31) >> 3)) ^ ((row >> 1) & 3)) << 4);
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __half h0, h1, l0, l1;
              split_h2(t[8 * cc + 2 * e], h0, l0);
              split_h2(t[8 * cc + 2 * e + 1], h1, l1);
              const __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
              hw[e] = *reinterpret_cast<const uint32_t*>(&hh);
              lw[e] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            *reinterpret_cast<uint4*>(base_ptr + L::oA + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(base_ptr + L::oA + NKA * L::kBox + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
          fence_proxy_async();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kAFull));
        }
        // the stores run under the chain MMA just released: chain accumulator -> operand boxes -> MMA is the critical path
        if (g >= 1) store_tile(y, g - 1, m0);
        if (g == n_groups - 1) store_tile(t, g, m0);  // the last group passes through conv1 only
      }
    }
    if (lane == 0) tma_store_wait_all();  // every store of this warp has landed before the CTA exits
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(L::kTmemCols));
  }
}

// Weight blobs in exactly the bytes the kernel's operand boxes hold: per group [W1 hi boxes | W1 lo boxes | Wc hi | Wc lo],
// a box = [WP rows (output channel) x 32 halves (input channel)] in the SWIZZLE_64B pattern; then b1 [8][WP], bc [8][WP].
__global__ void __launch_bounds__(256) k_front_pack(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ wc,
                                                    const float* __restrict__ bc, int w, int c_in, int n_groups, int wp, int k1,
                                                    uint8_t* __restrict__ pack) {
  const int nkx = k1 / 32, nka = wp / 32;
  const uint32_t kw1 = 2u * nkx * wp * 64u, kwc = 2u * nka * wp * 64u, blob = kw1 + kwc;
  const int per_group = (nkx + nka) * wp * 32;  // operand elements (hi / lo pairs) of one group
  const int64_t n_elem = (int64_t)kFrontMaxGroups * per_group;
  const int64_t n_shift = 2 * kFrontMaxGroups * wp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem + n_shift; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n_elem) {
      const int g = (int)(i / per_group);
      int e = (int)(i - (int64_t)g * per_group);
      const bool chain = e >= nkx * wp * 32;
      if (chain) e -= nkx * wp * 32;
      const int kc = e / (wp * 32), n = (e / 32) % wp, kl = e & 31, k = 32 * kc + kl;
      float v = 0.f;
      if (!chain) {
        if (g < n_groups && n < w && k < c_in) v = w1[((int64_t)g * w + n) * c_in + k];
      } else {
        if (g < n_groups - 1 && n < w && k < w) v = wc[((int64_t)g * w + n) * w + k];
      }
      __half hi, lo;
      split_h2(v, hi, lo);
      const int nk = chain ? nka : nkx;
      const uint32_t off = (uint32_t)g * blob + (chain ? kw1 : 0u) + (uint32_t)kc * wp * 64u + (uint32_t)n * 64u +
                           (uint32_t)((((kl >> 3)) ^ ((n >> 1) & 3)) << 4) + (uint32_t)(kl & 7) * 2u;
      *reinterpret_cast<__half*>(pack + off) = hi;
      *reinterpret_cast<__half*>(pack + off + (uint32_t)nk * wp * 64u) = lo;
    } else {
      const int j = (int)(i - n_elem);
      const bool second = j >= kFrontMaxGroups * wp;
      const int jj = second ? j - kFrontMaxGroups * wp : j;
      const int g = jj / wp, c = jj % wp;
      float v = 0.f;
      if (!second) {
        if (g < n_groups && c < w) v = b1[g * w + c];
      } else {
        if (g < n_groups - 1 && c < w) v = bc[g * w + c];
      }
      reinterpret_cast<float*>(pack + (size_t)kFrontMaxGroups * blob)[j] = v;
    }
  }
}

struct FrontConfig {
  int wp, k1;
};
bool front_config(int width, int n_groups, int c_in, FrontConfig* cfg) {
  if (width < 16 || (width & 3) || n_groups < 2 || n_groups > kFrontMaxGroups || c_in < 4 || (c_in & 3)) return false;
  if (width <= 32 && c_in <= 32) { *cfg = {32, 32}; return true; }
  if (width > 32 && width <= 64 && c_in <= 64) { *cfg = {64, 64}; return true; }
  return false;
}
size_t front_pack_bytes(const FrontConfig& c) {
  return c.wp == 64 ? FrontLayout<64, 64>::kPackBytes : FrontLayout<32, 32>::kPackBytes;
}

template <int WP, int K1>
int launch_front(const float* x, int ld_x, int c_in, const uint8_t* pack, int width, int n_groups, int64_t m_rows, float* z, int ld_z,
                 int copy_x, cudaStream_t stream) {
  using L = FrontLayout<WP, K1>;
  static PerDeviceOnce once;
  const int rc_cfg = once.run([]() -> int {
    KP_CUDA_TRY(cudaFuncSetAttribute(k_res2net_front<WP, K1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kTotal));
    return KPREG_OK;
  });
  if (rc_cfg) return rc_cfg;
  if (m_rows >= ((int64_t)1 << 31) - kFrontRows) return KPREG_E_RANGE;  // 32-bit row coordinates in the kernel
  CUtensorMap mx, mz, mxc;
  {
    cuuint64_t dims[2] = {(cuuint64_t)c_in, (cuuint64_t)m_rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld_x * 4};
    cuuint32_t box[2] = {32, kFrontRows};
    if (!encode_f32_map(&mx, x, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return KPREG_E_CUDA;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)n_groups, (cuuint64_t)m_rows};
    cuuint64_t strides[2] = {(cuuint64_t)width * 4, (cuuint64_t)ld_z * 4};
    cuuint32_t box[3] = {(cuuint32_t)L::CW, 1, 32};
    if (!encode_f32_map(&mz, z, 3, dims, strides, box, L::CW == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) return KPREG_E_CUDA;
  }
  if (copy_x) {
    cuuint64_t dims[2] = {(cuuint64_t)c_in, (cuuint64_t)m_rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld_z * 4};
    cuuint32_t box[2] = {32, kFrontRows};
    if (!encode_f32_map(&mxc, z + (int64_t)n_groups * width, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return KPREG_E_CUDA;
  } else {
    memset(&mxc, 0, sizeof(mxc));
  }
  const int64_t tiles = ceil_div(m_rows, (int64_t)kFrontRows);
  const int64_t max_ctas = (int64_t)kNumSMs * L::kCtasPerSm;
  const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
  k_res2net_front<WP, K1><<<grid, L::kThreads, L::kTotal, stream>>>(mx, mz, mxc, pack, n_groups, m_rows, copy_x);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

}  // namespace
}  // namespace kpreg

using namespace kpreg;

extern "C" int kpreg_front_supported(int width, int n_groups, int c_in) {
  FrontConfig c;
  return front_config(width, n_groups, c_in, &c) ? 1 : 0;
}

extern "C" int kpreg_front_pack_bytes(int width, int n_groups, int c_in, size_t* bytes) {
  FrontConfig c;
  if (!bytes || !front_config(width, n_groups, c_in, &c)) return KPREG_E_INVALID;
  *bytes = align_up(front_pack_bytes(c), 256);
  return KPREG_OK;
}

extern "C" int kpreg_front_pack(const float* w1, const float* b1, const float* wc, const float* bc, int width, int n_groups, int c_in,
                                void* pack, size_t pack_bytes, void* stream_) {
  FrontConfig c;
  if (!w1 || !b1 || !wc || !bc || !pack || !front_config(width, n_groups, c_in, &c)) return KPREG_E_INVALID;
  if (pack_bytes < front_pack_bytes(c)) return KPREG_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(pack) & 15) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  k_front_pack<<<2 * kNumSMs, 256, 0, stream>>>(w1, b1, wc, bc, width, c_in, n_groups, c.wp, c.k1, static_cast<uint8_t*>(pack));
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

extern "C" int kpreg_front_forward(const float* x, int ld_x, int c_in, const void* pack, int width, int n_groups, int64_t m_rows,
                                   float* z, int ld_z, int copy_x, void* stream_) {
  FrontConfig c;
  if (!front_config(width, n_groups, c_in, &c) || m_rows < 0) return KPREG_E_INVALID;
  if (m_rows == 0) return KPREG_OK;
  if (!x || !pack || !z || ld_x < c_in || (ld_x & 3) || (ld_z & 3)) return KPREG_E_INVALID;
  if (ld_z < n_groups * width + (copy_x ? c_in : 0)) return KPREG_E_INVALID;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(pack)) & 15) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  ProfScope prof(KPREG_FAM_LINEAR, stream);
  const uint8_t* pk = static_cast<const uint8_t*>(pack);
  if (c.wp == 64) return launch_front<64, 64>(x, ld_x, c_in, pk, width, n_groups, m_rows, z, ld_z, copy_x ? 1 : 0, stream);
  return launch_front<32, 32>(x, ld_x, c_in, pk, width, n_groups, m_rows, z, ld_z, copy_x ? 1 : 0, stream);
}
