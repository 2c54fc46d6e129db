// Rigid KPConv forward / backward on sm_100a (CUDA-core parts).
//
// Replaces KPConv.forward() (reference models/backbone_kpconv/finegrained_kpconv_blocks.py:265-401):
//   out[n,:] = (1/num[n]) * sum_k ( sum_h infl[n,h,k] * x[idx[n,h],:] ) @ W[k]
//   infl     = clamp(1 - sqrt(|s[idx[n,h]] - q[n] - kp[k]|^2)/extent, 0)  (linear; :353-356)
//   num[n]   = max(1, #{h : sum_c x[idx[n,h],c] > 0})                      (:396-399)
// split into
//   k_row_positive  : per support row, is its feature sum positive (the normalisation's predicate)
//   k_kpconv_gather : one warp per query — lanes compute the K influences of 32 neighbours at a
//                     time into shared memory, then every lane owns channels and streams the
//                     gathered feature rows once, accumulating all K kernel points in registers;
//                     writes the [n_q, K*c_in] aggregate (the gathered [n_q,H,c_in] tensor of the
//                     reference is never materialised) and num[n]
//   contraction     : [n_q, K*c_in] x [K*c_in, c_out] — tcgen05 kernel in kpconv_gemm.cu, or the
//                     fp32 CUDA-core GEMM below (also used by the backward pass).
// Shadow neighbours (index >= n_s; the reference's 1e6 point and zero feature row, :296/:375)
// contribute exactly zero and are skipped.
#include <cstdlib>

#include "common.cuh"

namespace kpreg {

// kpconv_gemm.cu: tcgen05 3xTF32 GEMM
int launch_gemm_tc(const float* a, int lda, const float* w_split, float* c, int ldc, int64_t m, int kd, int n,
                   const float* row_scale, const float* col_scale, const float* col_shift, const float* residual, int ld_res,
                   int act, float slope, float* out2, int ld2, const float* addend, int ld_add, const float* post_res, int ld_post,
                   int post_act, cudaStream_t stream);
int kpconv_gemm_tc_prepare_weights(const float* weights, int kd, int n, int transpose, float* w_split, cudaStream_t stream);
size_t kpconv_gemm_tc_weight_bytes(int kd, int n);
bool gemm_tc_supported(int64_t m, int kd, int n, int lda, const void* a);
// kpconv_gemm_tn.cu: tcgen05 split-K weight-gradient GEMM  C += A^T (B * row_scale)
int launch_gemm_tn(const float* a, int lda, const float* b, int ldb, const float* row_scale, float* c, int ldc, int64_t k_rows,
                   int ma_dim, int nb_dim, int c_transposed, void* split_ws, cudaStream_t stream);
size_t gemm_tn_workspace_bytes(int64_t k_rows, int nb_dim);
bool gemm_tn_supported(int64_t k_rows, int ma_dim, int nb_dim, int lda, const void* a);

namespace {

constexpr int KMAX = 16;             // kernel points handled per neighbour (reference ships 15)
constexpr int KSTRIDE = 20;          // floats per neighbour row of influences in shared memory: 16-byte aligned, conflict-free float4 stores
constexpr int kGatherWarps = 4;      // warps (= queries in flight) per CTA
bool g_gather_mma = true;            // aggregate on mma.sync (3xTF32); false = fp32 FFMA kernel (KPREG_GATHER_FFMA=1)
bool g_c1_by_neighbour = false;      // KPREG_C1_BY_NEIGHBOUR=1: the lane-per-neighbour c_in == 1 kernel also for 'sum' aggregation (A/B)
bool g_no_c1 = false;                // KPREG_NO_C1=1: c_in == 1 goes through the generic gather + GEMM path (A/B measurements)
bool g_gather_novec = false;         // KPREG_GATHER_NOVEC=1: scalar-load channel binding even for aligned rows (A/B measurements)
// Measured (profiles/r3_sweep_gather_grid.txt, 64 pairs): 32 CTAs per SM leave a 5-6-wave grid whose last wave runs at falling
// occupancy; 128 per SM with slices of >= 16 queries took the KPConv calls of a step from 20.15 to 19.63 ms, 256+ lose it again.
int g_gather_ctas_per_sm = 128;      // KPREG_GATHER_CTAS_PER_SM: upper bound of the gather-style grids, in CTAs per SM
int g_gather_min_slice = 16;         // KPREG_GATHER_MIN_SLICE: fewest queries a CTA's contiguous slice holds
bool g_gather_groups = true;         // KPREG_GATHER_GROUPS=0: rows wider than 64 channels stay on one warp per query (A/B measurements)

// Grid of the warp-per-query kernels (each CTA walks a contiguous slice of the processing order).
int gather_blocks(int64_t n_q) {
  // small query sets keep one query per warp (a single ModelNet pair must still spread over the SMs)
  const int64_t fine = (n_q + kGatherWarps - 1) / kGatherWarps, sliced = (n_q + g_gather_min_slice - 1) / g_gather_min_slice;
  const int64_t floor_blocks = (int64_t)kNumSMs * 32, cap = (int64_t)kNumSMs * g_gather_ctas_per_sm;
  int64_t blocks = sliced > floor_blocks ? sliced : floor_blocks;
  if (blocks > fine) blocks = fine;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

__global__ void __launch_bounds__(256) k_row_positive(const float* __restrict__ x, int64_t n_s, int c_in,
                                                      unsigned char* __restrict__ pos) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_s) return;
  double acc = 0.0;
  for (int c = lane; c < c_in; c += 32) acc += (double)x[row * c_in + c];
  acc = warp_sum(acc);
  if (lane == 0) pos[row] = (float)acc > 0.0f ? 1 : 0;
}

// The same predicate for 16-byte aligned rows of c_in = 4 * k channels: G = 2^m >= c_in / 4 lanes per row read one float4 each
// (32 / G rows per warp and load instruction), so short rows — 128 bytes at c_in = 32 — still move 512 bytes per instruction.
template <int G>
__global__ void __launch_bounds__(256) k_row_positive_vec(const float* __restrict__ x, int64_t n_s, int c_in,
                                                          unsigned char* __restrict__ pos) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t / G;
  const int sub = (int)(t - row * G);
  double acc = 0.0;
  if (row < n_s) {
    for (int c = 4 * sub; c < c_in; c += 4 * G) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * c_in + c));
      acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (row < n_s && sub == 0) pos[row] = (float)acc > 0.0f ? 1 : 0;
}

// rsqrt.approx.ftz: the argument is clamped to >= 1e-30 (a normal number) by every caller, so the denormal pre-scaling
// that rsqrtf() compiles to (a compare and two conditional multiplies per call) can never trigger.
__device__ __forceinline__ float rsqrt_fast(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Influence of the K kernel points on one neighbour (relative position rx,ry,rz).  INFL >= 0 fixes the influence mode at
// compile time (no per-kernel-point branches); INFL < 0 reads it from `influence`.
template <int INFL = -1>
__device__ __forceinline__ void influences(float rx, float ry, float rz, const float* __restrict__ s_kp, int n_kpts,
                                           float extent, int influence, int aggregation, float* __restrict__ w_out) {
  const float inv_extent = 1.0f / extent;
  float best = 3.4e38f;
  int best_k = 0;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    float w = 0.f;
    if (k < n_kpts) {
      const float dx = rx - s_kp[3 * k], dy = ry - s_kp[3 * k + 1], dz = rz - s_kp[3 * k + 2];
      const float d2 = dx * dx + dy * dy + dz * dz;
      const int mode = INFL >= 0 ? INFL : influence;
      if (mode == 1) w = fmaxf(1.0f - d2 * rsqrt_fast(fmaxf(d2, 1e-30f)) * inv_extent, 0.0f);
      else if (mode == 2) { const float sig = extent * 0.3f; w = expf(-d2 / (2.0f * sig * sig + 1e-9f)); }
      else w = 1.0f;
      if (d2 < best) { best = d2; best_k = k; }
    }
    w_out[k] = w;
  }
  if (aggregation == 1) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) if (k != best_k) w_out[k] = 0.f;
  }
}

template <typename IdxT, int CPL>
__global__ void __launch_bounds__(kGatherWarps * 32) k_kpconv_gather(
    const float* __restrict__ q_pts, const float* __restrict__ s_pts, const IdxT* __restrict__ idx, const float* __restrict__ x,
    const unsigned char* __restrict__ row_pos, const float* __restrict__ kernel_points, int64_t n_q, int64_t n_s, int n_nbrs,
    int n_kpts, int c_in, float extent, int influence, int aggregation, float* __restrict__ agg, float* __restrict__ inv_num,
    const int32_t* __restrict__ order) {
  __shared__ float s_kp[KMAX * 3];
  __shared__ __align__(16) float s_w[kGatherWarps][32][KSTRIDE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < KMAX * 3) s_kp[threadIdx.x] = threadIdx.x < n_kpts * 3 ? kernel_points[threadIdx.x] : 0.f;
  __syncthreads();

  // each CTA owns a CONTIGUOUS slice of the (spatially sorted) processing order, so the neighbourhoods its warps
  // gather overlap and stay in this SM's L1
  const int64_t per_cta = (n_q + gridDim.x - 1) / gridDim.x;
  const int64_t it_end = min(n_q, (int64_t)(blockIdx.x + 1) * per_cta);
  for (int64_t it = (int64_t)blockIdx.x * per_cta + warp; it < it_end; it += kGatherWarps) {
    const int64_t n = order ? (int64_t)order[it] : it;  // processing order only, never the result
    const float qx = q_pts[3 * n], qy = q_pts[3 * n + 1], qz = q_pts[3 * n + 2];
    float acc[CPL][KMAX];
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc)
#pragma unroll
      for (int k = 0; k < KMAX; ++k) acc[cc][k] = 0.f;
    int num = 0;

    for (int h0 = 0; h0 < n_nbrs; h0 += 32) {
      const int h = h0 + lane;
      int64_t j = n_s;
      if (h < n_nbrs) j = (int64_t)idx[n * n_nbrs + h];
      const bool valid = j >= 0 && j < n_s;
      if (valid) {
        float w[KMAX];
        influences(s_pts[3 * j] - qx, s_pts[3 * j + 1] - qy, s_pts[3 * j + 2] - qz, s_kp, n_kpts, extent, influence,
                   aggregation, w);
#pragma unroll
        for (int k = 0; k < KMAX; k += 4)
          *reinterpret_cast<float4*>(&s_w[warp][lane][k]) = make_float4(w[k], w[k + 1], w[k + 2], w[k + 3]);
      }
      num += __popc(__ballot_sync(0xffffffffu, valid && row_pos[j] != 0));
      const unsigned int vmask = __ballot_sync(0xffffffffu, valid);
      __syncwarp();
      const int lim = min(32, n_nbrs - h0);
      for (int hh = 0; hh < lim; ++hh) {
        if (!((vmask >> hh) & 1u)) continue;
        const int64_t jj = __shfl_sync(0xffffffffu, j, hh);
        const float* __restrict__ xr = x + jj * c_in;
        float xv[CPL];
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
          const int c = lane + 32 * cc;
          xv[cc] = c < c_in ? xr[c] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < KMAX; k += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(&s_w[warp][hh][k]);  // broadcast: one wavefront
#pragma unroll
          for (int cc = 0; cc < CPL; ++cc) {
            acc[cc][k + 0] = fmaf(w4.x, xv[cc], acc[cc][k + 0]);
            acc[cc][k + 1] = fmaf(w4.y, xv[cc], acc[cc][k + 1]);
            acc[cc][k + 2] = fmaf(w4.z, xv[cc], acc[cc][k + 2]);
            acc[cc][k + 3] = fmaf(w4.w, xv[cc], acc[cc][k + 3]);
          }
        }
      }
      __syncwarp();
    }
    float* __restrict__ arow = agg + n * (int64_t)n_kpts * c_in;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < n_kpts) {
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
          const int c = lane + 32 * cc;
          if (c < c_in) arow[k * c_in + c] = acc[cc][k];
        }
      }
    }
    if (lane == 0) inv_num[n] = 1.0f / (float)max(num, 1);
  }
}

#ifndef KPREG_GATHER_PREFETCH
#define KPREG_GATHER_PREFETCH 1
#endif
// ---- tensor-core variant of the gather/aggregate step ---------------------------------------------------
// agg[n, k, c] = sum_h infl[n,h,k] * x[idx[n,h], c] is, per query, a [16 x H] x [H x c_in] product.  One warp per
// query runs it on mma.sync m16n8k8 (TF32, fp32 accumulate) with the same 3xTF32 operand split as the
// contraction GEMM (fp32-grade accuracy): M = kernel points (15 + 1 zero row), K = 8 neighbours per step,
// N = 8 channels per tile.  Every lane computes exactly the four influences of its A fragment
// (kernel points g, g+8 x neighbours t, t+4 with g = lane/4, t = lane%4) — no influence is computed twice —
// and loads its B fragment straight from the two neighbours' feature rows (8 lanes read 32 contiguous bytes).
// 3xTF32 operand split x = hi + lo.  hi = x rounded to the nearest TF32 by integer arithmetic on the bit pattern (add half
// an ulp of the 10-bit mantissa, clear the 13 low bits: two instructions; ptxas expands cvt.rna.tf32.f32 into an
// add + Inf/NaN test + select, and the rounding of lo into four more).  lo = x - hi is exact in fp32 and is handed to the
// tensor core as is: HMMA ignores the 13 low mantissa bits of a TF32 operand, i.e. truncates lo to 11 significant bits,
// an error of at most 2^-21 |x|, below the lo*lo term 3xTF32 drops anyway.  (Inf stays Inf in hi and gives NaN in lo; finite
// values within half a TF32 ulp of FLT_MAX round to Inf — neither occurs in feature data.)
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                 "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

template <int INFL = -1>
__device__ __forceinline__ float influence_one(float rx, float ry, float rz, float kx, float ky, float kz, float inv_extent,
                                               float extent, int influence, float& d2_out) {
  const float dx = rx - kx, dy = ry - ky, dz = rz - kz;
  const float d2 = dx * dx + dy * dy + dz * dz;
  d2_out = d2;
  const int mode = INFL >= 0 ? INFL : influence;
  if (mode == 1) return fmaxf(1.0f - d2 * rsqrt_fast(fmaxf(d2, 1e-30f)) * inv_extent, 0.0f);
  if (mode == 2) { const float sig = extent * 0.3f; return expf(-d2 / (2.0f * sig * sig + 1e-9f)); }
  return 1.0f;
}

// VEC (c_in a multiple of 4, 16-byte aligned rows): the MMA's N index is only a label, so column g of tile i is
// bound to channel NT * g + i instead of 8 * i + g.  Lane g then needs NT CONTIGUOUS channels of each of its two
// neighbour rows (float4 loads; the eight lanes sharing a neighbour read one contiguous 32 * NT-byte span), and
// its D fragment covers the 2 * NT contiguous channels [2t * NT, 2t * NT + 2 NT) of rows g and g + 8 (float4 stores).
// The index row of the next query is requested before the current query's work.  (Requesting the feature rows of k-step
// s + 1 before the MMAs of k-step s was measured SLOWER on B200 — 5-9 % — through the registers it costs.)
// INFL: 1 = the 'linear' influence of every shipped config fixed at compile time, -1 = mode read at run time.
// resident CTAs per SM the register allocation must leave room for (the occupancies the kernel was tuned at)
constexpr int gather_min_blocks(int nt) { return nt <= 4 ? 6 : (nt <= 8 ? 5 : (nt <= 16 ? 3 : 2)); }

// GRP (wide rows, c_in > 8 * NT): the work item is (query, channel group of 8 * NT channels) instead of the query, consecutive
// items = the groups of one query, so the warps of a CTA share the query's index row and neighbour coordinates through L1 and
// each recomputes the influences.  One warp holding all of a 128- / 256-channel row needs 64 / 128 accumulator registers
// (3 / 2 CTAs per SM, and the kernel is latency-bound): measured per query and SM, 2 000 / 5 800 cycles against 719 at 64 channels.
template <typename IdxT, int NT, bool VEC, int INFL, bool GRP = false>  // NT = number of 8-channel tiles (c_in <= 8 * NT unless GRP)
__global__ void __launch_bounds__(kGatherWarps * 32, gather_min_blocks(NT)) k_kpconv_gather_mma(
    const float* __restrict__ q_pts, const float* __restrict__ s_pts, const IdxT* __restrict__ idx, const float* __restrict__ x,
    const unsigned char* __restrict__ row_pos, const float* __restrict__ kernel_points, int64_t n_q, int64_t n_s, int n_nbrs,
    int n_kpts, int c_in, float extent, int influence, int aggregation, float* __restrict__ agg, float* __restrict__ inv_num,
    const int32_t* __restrict__ order, int n_groups) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float inv_extent = 1.0f / extent;
  // this lane's two kernel points (rows g and g + 8 of the A fragment); row 15 is the zero pad when K = 15
  const bool k0_ok = g < n_kpts, k1_ok = g + 8 < n_kpts;
  const float k0x = k0_ok ? kernel_points[3 * g] : 0.f, k0y = k0_ok ? kernel_points[3 * g + 1] : 0.f,
              k0z = k0_ok ? kernel_points[3 * g + 2] : 0.f;
  const float k1x = k1_ok ? kernel_points[3 * (g + 8)] : 0.f, k1y = k1_ok ? kernel_points[3 * (g + 8) + 1] : 0.f,
              k1z = k1_ok ? kernel_points[3 * (g + 8) + 2] : 0.f;
  const int n_s32 = (int)n_s;  // the launcher routes n_s >= 2^31 to the FFMA kernel: support rows fit an int here

  // each CTA owns a CONTIGUOUS slice of the (spatially sorted) processing order, so the neighbourhoods its warps
  // gather overlap and stay in this SM's L1
  const int64_t n_items = GRP ? n_q * n_groups : n_q;
  const int64_t per_cta = (n_items + gridDim.x - 1) / gridDim.x;
  const int64_t it_end = min(n_items, (int64_t)(blockIdx.x + 1) * per_cta);
  int64_t it = (int64_t)blockIdx.x * per_cta + warp;
  // index rows (neighbours h = lane and h = lane + 32) are fetched TWO queries ahead: one query ahead their values are in
  // registers, and the support points / feature rows they name are requested into L1 while the current query is processed
  // (the gather is latency-bound: L1 hit rate 78 % without this)
  int64_t n_nx = 0, n_n2 = 0;
  int grp_nx = 0, grp_n2 = 0;  // (GRP) channel group of the item
  IdxT raw_nx[2] = {(IdxT)n_s32, (IdxT)n_s32}, raw_n2[2] = {(IdxT)n_s32, (IdxT)n_s32};
  auto load_row = [&](int64_t i, int64_t& nn, int& gg, IdxT (&rw)[2]) {
    rw[0] = rw[1] = (IdxT)n_s32;
    if (i < it_end) {
      int64_t q = i;
      if constexpr (GRP) {
        q = i / n_groups;
        gg = (int)(i - q * n_groups);
      }
      nn = order ? (int64_t)order[q] : q;  // processing order only, never the result
#pragma unroll
      for (int r = 0; r < 2; ++r)
        if (lane + 32 * r < n_nbrs) rw[r] = idx[nn * n_nbrs + lane + 32 * r];
    }
  };
  load_row(it, n_nx, grp_nx, raw_nx);
  load_row(it + kGatherWarps, n_n2, grp_n2, raw_n2);
  for (; it < it_end; it += kGatherWarps) {
    const int64_t n = n_nx;
    const int c_off = GRP ? grp_nx * 8 * NT : 0;  // first channel of this item's group
    const IdxT raw[2] = {raw_nx[0], raw_nx[1]};
    n_nx = n_n2;
    grp_nx = grp_n2;
    raw_nx[0] = raw_n2[0];
    raw_nx[1] = raw_n2[1];
    load_row(it + 2 * kGatherWarps, n_n2, grp_n2, raw_n2);
#if KPREG_GATHER_PREFETCH
    // (measured: -3 % at c_in = 32, +3-5 % at c_in >= 64 — rows of several cache lines — so narrow rows only)
    if (NT <= 4 && it + kGatherWarps < it_end) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (raw_nx[r] >= 0 && raw_nx[r] < (IdxT)n_s32) {
          const float* fr = x + (int64_t)raw_nx[r] * c_in;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(s_pts + 3 * (int64_t)raw_nx[r]));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(fr));
        }
      }
    }
#endif
    const float qx = q_pts[3 * n], qy = q_pts[3 * n + 1], qz = q_pts[3 * n + 2];
    // relative positions of this lane's two neighbours, fetched once and shuffled per k-step
    int jn[2];
    float rx[2], ry[2], rz[2];
    int num = 0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const bool valid = raw[r] >= 0 && raw[r] < (IdxT)n_s32;
      const int j = valid ? (int)raw[r] : -1;
      jn[r] = j;
      rx[r] = ry[r] = rz[r] = 0.f;
      if (valid) { rx[r] = s_pts[3 * (int64_t)j] - qx; ry[r] = s_pts[3 * (int64_t)j + 1] - qy; rz[r] = s_pts[3 * (int64_t)j + 2] - qz; }
      num += __popc(__ballot_sync(0xffffffffu, valid && row_pos[j] != 0));
    }
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    // neighbour rows of this lane's fragment columns in k-step h0: h0 + t and h0 + t + 4 (-1 = shadow)
    auto step_rows = [&](int h0, int& ja, int& jb) {
      const int ha = h0 + t, hb = h0 + t + 4;  // (h >> 5) is warp-uniform per k-step: h0 is a multiple of 8
      ja = __shfl_sync(0xffffffffu, (ha >> 5) ? jn[1] : jn[0], ha & 31);
      jb = __shfl_sync(0xffffffffu, (hb >> 5) ? jn[1] : jn[0], hb & 31);
    };
    // A fragment of k-step h0: the four influences (k = g | g + 8) x (h = h0 + t | h0 + t + 4), split hi / lo
    auto step_influences = [&](int h0, bool va, bool vb, float (&a_hi)[4], float (&a_lo)[4]) {
      const int ha = h0 + t, hb = h0 + t + 4;
      const int ra = ha >> 5, rb = hb >> 5;
      const float ax = __shfl_sync(0xffffffffu, ra ? rx[1] : rx[0], ha & 31), ay = __shfl_sync(0xffffffffu, ra ? ry[1] : ry[0], ha & 31),
                  az = __shfl_sync(0xffffffffu, ra ? rz[1] : rz[0], ha & 31);
      const float bx = __shfl_sync(0xffffffffu, rb ? rx[1] : rx[0], hb & 31), by = __shfl_sync(0xffffffffu, rb ? ry[1] : ry[0], hb & 31),
                  bz = __shfl_sync(0xffffffffu, rb ? rz[1] : rz[0], hb & 31);
      float w[4], d2[4];
      w[0] = influence_one<INFL>(ax, ay, az, k0x, k0y, k0z, inv_extent, extent, influence, d2[0]);  // (k = g,     h = t)
      w[1] = influence_one<INFL>(ax, ay, az, k1x, k1y, k1z, inv_extent, extent, influence, d2[1]);  // (k = g + 8, h = t)
      w[2] = influence_one<INFL>(bx, by, bz, k0x, k0y, k0z, inv_extent, extent, influence, d2[2]);  // (k = g,     h = t + 4)
      w[3] = influence_one<INFL>(bx, by, bz, k1x, k1y, k1z, inv_extent, extent, influence, d2[3]);  // (k = g + 8, h = t + 4)
      if (!k0_ok) { w[0] = w[2] = 0.f; d2[0] = d2[2] = 3.4e38f; }
      if (!k1_ok) { w[1] = w[3] = 0.f; d2[1] = d2[3] = 3.4e38f; }
      if (!va) w[0] = w[1] = 0.f;
      if (!vb) w[2] = w[3] = 0.f;
      if (aggregation == 1) {
        // 'closest': only the nearest kernel point of each neighbour keeps its influence (first minimum wins)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float best = d2[2 * hh];
          int best_k = g;
          if (d2[2 * hh + 1] < best) { best = d2[2 * hh + 1]; best_k = g + 8; }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {  // reduce over the eight lanes that share this neighbour (same t)
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
            if (ob < best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
          }
          if (best_k != g) w[2 * hh] = 0.f;
          if (best_k != g + 8) w[2 * hh + 1] = 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) tf32_split(w[i], a_hi[i], a_lo[i]);
    };
    auto mma3 = [&](float (&d)[4], const float (&a_hi)[4], const float (&a_lo)[4], float b0, float b1) {
      float b0h, b0l, b1h, b1l;
      tf32_split(b0, b0h, b0l);
      tf32_split(b1, b1h, b1l);
      mma_tf32(d, a_lo, b0h, b1h);
      mma_tf32(d, a_hi, b0l, b1l);
      mma_tf32(d, a_hi, b0h, b1h);
    };
    auto comp4 = [](const float4& v, int e) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); };

    for (int h0 = 0; h0 < n_nbrs; h0 += 8) {
      int ja, jb;
      step_rows(h0, ja, jb);
      const bool va = ja >= 0, vb = jb >= 0;
      if (!__any_sync(0xffffffffu, va || vb)) continue;  // eight shadow neighbours: nothing to add
      float a_hi[4], a_lo[4];
      step_influences(h0, va, vb, a_hi, a_lo);
      const float* __restrict__ xa = x + (int64_t)(va ? ja : 0) * c_in + c_off;
      const float* __restrict__ xb = x + (int64_t)(vb ? jb : 0) * c_in + c_off;
      // (Measured: feeding the B operand pre-split — features split once per layer into interleaved (hi, lo) pairs instead of
      // once per gathering lane, 30 % fewer instructions — made the step's aggregate kernels 40 % SLOWER: the kernel is bound
      // by the gather path (L1 / L2 latency at a 78 % L1 hit rate), and pre-split rows are twice as long.)
      if constexpr (VEC) {
        float4 fa[NT / 4], fb[NT / 4];
#pragma unroll
        for (int m = 0; m < NT / 4; ++m) {
          const int c = NT * g + 4 * m;
          fa[m] = (va && c_off + c < c_in) ? __ldg(reinterpret_cast<const float4*>(xa + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          fb[m] = (vb && c_off + c < c_in) ? __ldg(reinterpret_cast<const float4*>(xb + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int m = 0; m < NT / 4; ++m)
#pragma unroll
          for (int e = 0; e < 4; ++e) mma3(acc[4 * m + e], a_hi, a_lo, comp4(fa[m], e), comp4(fb[m], e));
      } else {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const int c = 8 * i + g;
          mma3(acc[i], a_hi, a_lo, (va && c_off + c < c_in) ? xa[c] : 0.f, (vb && c_off + c < c_in) ? xb[c] : 0.f);
        }
      }
    }
    // D fragment: rows (kernel points) g and g + 8, columns (channels) 8 i + 2 t, + 1
    float* __restrict__ arow = agg + n * (int64_t)n_kpts * c_in + c_off;
    const bool first_group = !GRP || c_off == 0;  // one item per query reports the neighbour count
    if constexpr (VEC) {
      // D fragment under the VEC binding: acc[i][0|2] <-> channel 2t * NT + i, acc[i][1|3] <-> channel (2t + 1) * NT + i
#pragma unroll
      for (int m = 0; m < NT / 4; ++m) {
        const int ce = 2 * t * NT + 4 * m, co = (2 * t + 1) * NT + 4 * m;
        if (c_off + ce < c_in) {
          if (k0_ok) *reinterpret_cast<float4*>(arow + g * c_in + ce) = make_float4(acc[4 * m][0], acc[4 * m + 1][0], acc[4 * m + 2][0], acc[4 * m + 3][0]);
          if (k1_ok) *reinterpret_cast<float4*>(arow + (g + 8) * c_in + ce) = make_float4(acc[4 * m][2], acc[4 * m + 1][2], acc[4 * m + 2][2], acc[4 * m + 3][2]);
        }
        if (c_off + co < c_in) {
          if (k0_ok) *reinterpret_cast<float4*>(arow + g * c_in + co) = make_float4(acc[4 * m][1], acc[4 * m + 1][1], acc[4 * m + 2][1], acc[4 * m + 3][1]);
          if (k1_ok) *reinterpret_cast<float4*>(arow + (g + 8) * c_in + co) = make_float4(acc[4 * m][3], acc[4 * m + 1][3], acc[4 * m + 2][3], acc[4 * m + 3][3]);
        }
      }
      if (lane == 0 && first_group) inv_num[n] = 1.0f / (float)max(num, 1);
      continue;
    }
    const bool pair_ok = (c_in & 1) == 0;  // (row * c_in + even column) is then 8-byte aligned: one float2 store
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int c = 8 * i + 2 * t;
      if (pair_ok && c_off + c + 1 < c_in) {
        if (k0_ok) *reinterpret_cast<float2*>(arow + g * c_in + c) = make_float2(acc[i][0], acc[i][1]);
        if (k1_ok) *reinterpret_cast<float2*>(arow + (g + 8) * c_in + c) = make_float2(acc[i][2], acc[i][3]);
      } else {
        if (k0_ok) {
          if (c_off + c < c_in) arow[g * c_in + c] = acc[i][0];
          if (c_off + c + 1 < c_in) arow[g * c_in + c + 1] = acc[i][1];
        }
        if (k1_ok) {
          if (c_off + c < c_in) arow[(g + 8) * c_in + c] = acc[i][2];
          if (c_off + c + 1 < c_in) arow[(g + 8) * c_in + c + 1] = acc[i][3];
        }
      }
    }
    if (lane == 0 && first_group) inv_num[n] = 1.0f / (float)max(num, 1);
  }
}


// ---- c_in == 1 (the encoder's first block: a single input feature per point) ----------------------------
// With one channel the aggregate is 15 numbers per query and the contraction a [15] x [15, c_out] product, so the whole
// operator is one kernel: a lane per neighbour evaluates the K influences (fp32 FFMA), a transposing butterfly leaves
// the total of kernel point k in lanes 2k / 2k+1 (16 shuffles), and every lane finishes c_out / 32 output channels
// from register-resident weights.  No aggregate, no row predicate pass (sum_c x > 0 is x > 0) and no GEMM launch.
template <typename IdxT, int CPL, int INFL>  // CPL = output channels per lane (c_out <= 32 * CPL); INFL as in k_kpconv_gather_mma
__global__ void __launch_bounds__(kGatherWarps * 32, 8) k_kpconv_c1(
    const float* __restrict__ q_pts, const float* __restrict__ s_pts, const IdxT* __restrict__ idx, const float* __restrict__ x,
    const float* __restrict__ kernel_points, const float* __restrict__ weights, int64_t n_q, int64_t n_s, int n_nbrs, int n_kpts,
    int c_out, float extent, int influence, int aggregation, float* __restrict__ out, const int32_t* __restrict__ order) {
  __shared__ float s_kp[KMAX * 3];
  __shared__ float s_wt[KMAX][32 * CPL];  // weights, zero-padded: lane reads column lane + 32 cc (conflict-free)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < KMAX * 3) s_kp[threadIdx.x] = threadIdx.x < n_kpts * 3 ? kernel_points[threadIdx.x] : 0.f;
  for (int i = threadIdx.x; i < KMAX * 32 * CPL; i += blockDim.x) {
    const int k = i / (32 * CPL), c = i % (32 * CPL);
    s_wt[k][c] = (k < n_kpts && c < c_out) ? weights[k * c_out + c] : 0.f;
  }
  __syncthreads();
  const float inv_extent = 1.0f / extent;
  // neighbours 0..31: a lane each.  Neighbours 32..: a lane each, or — for a short tail (H = 40: eight neighbours) in
  // 'sum' mode — eight lanes per quarter of the kernel points instead of 24 idle lanes.
  const int rem = n_nbrs - 32;
  const bool quarter = rem > 0 && rem <= 8 && aggregation == 0;
  const int h_a = lane, h_b = 32 + (quarter ? (lane & 7) : lane), q4 = (lane >> 3) * 4;
  const int64_t per_cta = (n_q + gridDim.x - 1) / gridDim.x;
  const int64_t it_end = min(n_q, (int64_t)(blockIdx.x + 1) * per_cta);
  int64_t it = (int64_t)blockIdx.x * per_cta + warp;
  // the index row of the NEXT query is fetched while the current one is processed
  int64_t n_nx = 0, ja_nx = n_s, jb_nx = n_s;
  if (it < it_end) {
    n_nx = order ? (int64_t)order[it] : it;
    if (h_a < n_nbrs) ja_nx = (int64_t)idx[n_nx * n_nbrs + h_a];
    if (h_b < n_nbrs) jb_nx = (int64_t)idx[n_nx * n_nbrs + h_b];
  }
  for (; it < it_end; it += kGatherWarps) {
    const int64_t n = n_nx, ja = ja_nx, jb = jb_nx;  // processing order only, never the result
    if (it + kGatherWarps < it_end) {
      n_nx = order ? (int64_t)order[it + kGatherWarps] : it + kGatherWarps;
      ja_nx = h_a < n_nbrs ? (int64_t)idx[n_nx * n_nbrs + h_a] : n_s;
      jb_nx = h_b < n_nbrs ? (int64_t)idx[n_nx * n_nbrs + h_b] : n_s;
    }
    const float qx = q_pts[3 * n], qy = q_pts[3 * n + 1], qz = q_pts[3 * n + 2];
    const bool va = ja >= 0 && ja < n_s, vb = jb >= 0 && jb < n_s;
    float xa = 0.f, ax = 0.f, ay = 0.f, az = 0.f, xb = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
    if (va) { xa = x[ja]; ax = s_pts[3 * ja] - qx; ay = s_pts[3 * ja + 1] - qy; az = s_pts[3 * ja + 2] - qz; }
    if (vb) { xb = x[jb]; bx = s_pts[3 * jb] - qx; by = s_pts[3 * jb + 1] - qy; bz = s_pts[3 * jb + 2] - qz; }
    int num = __popc(__ballot_sync(0xffffffffu, va && xa > 0.f));
    if (rem > 0) num += __popc(__ballot_sync(0xffffffffu, vb && xb > 0.f && (!quarter || lane < 8)));
    float v[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) v[k] = 0.f;
    if (va) {
      float w[KMAX];
      influences<INFL>(ax, ay, az, s_kp, n_kpts, extent, influence, aggregation, w);
#pragma unroll
      for (int k = 0; k < KMAX; ++k) v[k] = w[k] * xa;
    }
    if (quarter) {
      float u[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = q4 + kk;
        float d2;
        const float w = influence_one<INFL>(bx, by, bz, s_kp[3 * k], s_kp[3 * k + 1], s_kp[3 * k + 2], inv_extent, extent, influence, d2);
        u[kk] = (vb && k < n_kpts) ? w * xb : 0.f;
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k) v[k] += ((k & ~3) == q4) ? u[k & 3] : 0.f;
    } else if (vb) {
      float w[KMAX];
      influences<INFL>(bx, by, bz, s_kp, n_kpts, extent, influence, aggregation, w);
#pragma unroll
      for (int k = 0; k < KMAX; ++k) v[k] = fmaf(w[k], xb, v[k]);
    }
    // transposing butterfly: after offsets 16, 8, 4, 2 a lane holds one kernel point's partial, k = lane >> 1
    float v8[8], v4[4], v2[2];
    {
      const bool up = (lane & 16) != 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) v8[i] = (up ? v[i + 8] : v[i]) + __shfl_xor_sync(0xffffffffu, up ? v[i] : v[i + 8], 16);
    }
    {
      const bool up = (lane & 8) != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) v4[i] = (up ? v8[i + 4] : v8[i]) + __shfl_xor_sync(0xffffffffu, up ? v8[i] : v8[i + 4], 8);
    }
    {
      const bool up = (lane & 4) != 0;
#pragma unroll
      for (int i = 0; i < 2; ++i) v2[i] = (up ? v4[i + 2] : v4[i]) + __shfl_xor_sync(0xffffffffu, up ? v4[i] : v4[i + 2], 4);
    }
    float tot;
    {
      const bool up = (lane & 2) != 0;
      tot = (up ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, up ? v2[0] : v2[1], 2);
      tot += __shfl_xor_sync(0xffffffffu, tot, 1);
    }
    float o[CPL];
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) o[cc] = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const float ak = __shfl_sync(0xffffffffu, tot, 2 * k);
#pragma unroll
      for (int cc = 0; cc < CPL; ++cc) o[cc] = fmaf(ak, s_wt[k][lane + 32 * cc], o[cc]);
    }
    const float inv = 1.0f / (float)max(num, 1);
    float* __restrict__ orow = out + n * (int64_t)c_out;
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
      const int c = lane + 32 * cc;
      if (c < c_out) orow[c] = o[cc] * inv;
    }
  }
}

// The same operator with the lanes bound to (kernel point, neighbour parity) instead of neighbours — 'sum' aggregation only.
// k_kpconv_c1 spends ~800 SASS instructions per query, of which only ~250 evaluate influences: the rest is the transposing
// butterfly that turns "a lane per neighbour" into "a lane per kernel point" (it runs at the issue limit: 2.0 ms for 2.65 M
// queries).  Here lanes 2k and 2k+1 own kernel point k and walk the even / odd neighbours, whose relative position and
// feature the warp stages once in shared memory ([h] float4, broadcast reads): 12 instructions per (kernel point, neighbour)
// and one shuffle to join the two halves — lane 2k holds kernel point k's total exactly where the output stage expects it.
template <typename IdxT, int CPL, int INFL>
__global__ void __launch_bounds__(kGatherWarps * 32, 8) k_kpconv_c1_kp(
    const float* __restrict__ q_pts, const float* __restrict__ s_pts, const IdxT* __restrict__ idx, const float* __restrict__ x,
    const float* __restrict__ kernel_points, const float* __restrict__ weights, int64_t n_q, int64_t n_s, int n_nbrs, int n_kpts,
    int c_out, float extent, int influence, float* __restrict__ out, const int32_t* __restrict__ order) {
  __shared__ float s_wt[KMAX][32 * CPL];  // weights, zero-padded: lane reads column lane + 32 cc (conflict-free)
  __shared__ __align__(16) float4 s_nb[kGatherWarps][64];  // per warp: (rx, ry, rz, x) of the query's neighbours
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < KMAX * 32 * CPL; i += blockDim.x) {
    const int k = i / (32 * CPL), c = i % (32 * CPL);
    s_wt[k][c] = (k < n_kpts && c < c_out) ? weights[k * c_out + c] : 0.f;
  }
  __syncthreads();
  const float inv_extent = 1.0f / extent;
  const int kp = lane >> 1, par = lane & 1;
  const bool kp_ok = kp < n_kpts;
  const float kx = kp_ok ? kernel_points[3 * kp] : 0.f, ky = kp_ok ? kernel_points[3 * kp + 1] : 0.f,
              kz = kp_ok ? kernel_points[3 * kp + 2] : 0.f;
  const int h_a = lane, h_b = 32 + lane;
  const int64_t per_cta = (n_q + gridDim.x - 1) / gridDim.x;
  const int64_t it_end = min(n_q, (int64_t)(blockIdx.x + 1) * per_cta);
  int64_t it = (int64_t)blockIdx.x * per_cta + warp;
  // the index row of the NEXT query is fetched while the current one is processed
  int64_t n_nx = 0, ja_nx = n_s, jb_nx = n_s;
  if (it < it_end) {
    n_nx = order ? (int64_t)order[it] : it;  // processing order only, never the result
    if (h_a < n_nbrs) ja_nx = (int64_t)idx[n_nx * n_nbrs + h_a];
    if (h_b < n_nbrs) jb_nx = (int64_t)idx[n_nx * n_nbrs + h_b];
  }
  float4* __restrict__ nb = s_nb[warp];
  for (; it < it_end; it += kGatherWarps) {
    const int64_t n = n_nx, ja = ja_nx, jb = jb_nx;
    if (it + kGatherWarps < it_end) {
      n_nx = order ? (int64_t)order[it + kGatherWarps] : it + kGatherWarps;
      ja_nx = h_a < n_nbrs ? (int64_t)idx[n_nx * n_nbrs + h_a] : n_s;
      jb_nx = h_b < n_nbrs ? (int64_t)idx[n_nx * n_nbrs + h_b] : n_s;
    }
    const float qx = q_pts[3 * n], qy = q_pts[3 * n + 1], qz = q_pts[3 * n + 2];
    const bool va = ja >= 0 && ja < n_s, vb = jb >= 0 && jb < n_s;
    // a shadow neighbour is staged with feature 0: whatever its influence, it adds nothing
    float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (va) pa = make_float4(s_pts[3 * ja] - qx, s_pts[3 * ja + 1] - qy, s_pts[3 * ja + 2] - qz, x[ja]);
    if (vb) pb = make_float4(s_pts[3 * jb] - qx, s_pts[3 * jb + 1] - qy, s_pts[3 * jb + 2] - qz, x[jb]);
    int num = __popc(__ballot_sync(0xffffffffu, va && pa.w > 0.f));
    if (n_nbrs > 32) num += __popc(__ballot_sync(0xffffffffu, vb && pb.w > 0.f));
    __syncwarp();  // the previous query's reads of the staging rows are done
    nb[h_a] = pa;
    if (h_b < 64) nb[h_b] = pb;
    __syncwarp();
    float acc = 0.f;
#pragma unroll 4
    for (int h = par; h < n_nbrs; h += 2) {
      const float4 p = nb[h];
      float d2;
      const float w = influence_one<INFL>(p.x, p.y, p.z, kx, ky, kz, inv_extent, extent, influence, d2);
      acc = fmaf(w, p.w, acc);
    }
    if (!kp_ok) acc = 0.f;
    const float tot = acc + __shfl_xor_sync(0xffffffffu, acc, 1);
    float o[CPL];
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) o[cc] = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const float ak = __shfl_sync(0xffffffffu, tot, 2 * k);
#pragma unroll
      for (int cc = 0; cc < CPL; ++cc) o[cc] = fmaf(ak, s_wt[k][lane + 32 * cc], o[cc]);
    }
    const float inv = 1.0f / (float)max(num, 1);
    float* __restrict__ orow = out + n * (int64_t)c_out;
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
      const int c = lane + 32 * cc;
      if (c < c_out) orow[c] = o[cc] * inv;
    }
  }
}

// d_x[idx[n,h], c] += sum_k infl[n,h,k] * d_agg[n,k,c]
template <typename IdxT, int CPL>
__global__ void __launch_bounds__(kGatherWarps * 32) k_kpconv_scatter(
    const float* __restrict__ q_pts, const float* __restrict__ s_pts, const IdxT* __restrict__ idx,
    const float* __restrict__ kernel_points, const float* __restrict__ d_agg, int64_t n_q, int64_t n_s, int n_nbrs, int n_kpts,
    int c_in, float extent, int influence, int aggregation, float* __restrict__ d_x, const int32_t* __restrict__ order) {
  __shared__ float s_kp[KMAX * 3];
  __shared__ __align__(16) float s_w[kGatherWarps][32][KSTRIDE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < KMAX * 3) s_kp[threadIdx.x] = threadIdx.x < n_kpts * 3 ? kernel_points[threadIdx.x] : 0.f;
  __syncthreads();
  // each CTA owns a CONTIGUOUS slice of the (spatially sorted) processing order, so the neighbourhoods its warps
  // gather overlap and stay in this SM's L1
  const int64_t per_cta = (n_q + gridDim.x - 1) / gridDim.x;
  const int64_t it_end = min(n_q, (int64_t)(blockIdx.x + 1) * per_cta);
  for (int64_t it = (int64_t)blockIdx.x * per_cta + warp; it < it_end; it += kGatherWarps) {
    const int64_t n = order ? (int64_t)order[it] : it;  // processing order only, never the result
    const float qx = q_pts[3 * n], qy = q_pts[3 * n + 1], qz = q_pts[3 * n + 2];
    float da[CPL][KMAX];
    const float* __restrict__ drow = d_agg + n * (int64_t)n_kpts * c_in;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
#pragma unroll
      for (int cc = 0; cc < CPL; ++cc) {
        const int c = lane + 32 * cc;
        da[cc][k] = (k < n_kpts && c < c_in) ? drow[k * c_in + c] : 0.f;
      }
    for (int h0 = 0; h0 < n_nbrs; h0 += 32) {
      const int h = h0 + lane;
      int64_t j = n_s;
      if (h < n_nbrs) j = (int64_t)idx[n * n_nbrs + h];
      const bool valid = j >= 0 && j < n_s;
      if (valid) {
        float w[KMAX];
        influences(s_pts[3 * j] - qx, s_pts[3 * j + 1] - qy, s_pts[3 * j + 2] - qz, s_kp, n_kpts, extent, influence,
                   aggregation, w);
#pragma unroll
        for (int k = 0; k < KMAX; k += 4)
          *reinterpret_cast<float4*>(&s_w[warp][lane][k]) = make_float4(w[k], w[k + 1], w[k + 2], w[k + 3]);
      }
      const unsigned int vmask = __ballot_sync(0xffffffffu, valid);
      __syncwarp();
      const int lim = min(32, n_nbrs - h0);
      for (int hh = 0; hh < lim; ++hh) {
        if (!((vmask >> hh) & 1u)) continue;
        const int64_t jj = __shfl_sync(0xffffffffu, j, hh);
        float v[CPL];
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) v[cc] = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          const float wk = s_w[warp][hh][k];
#pragma unroll
          for (int cc = 0; cc < CPL; ++cc) v[cc] = fmaf(wk, da[cc][k], v[cc]);
        }
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
          const int c = lane + 32 * cc;
          if (c < c_in) atomicAdd(d_x + jj * c_in + c, v[cc]);
        }
      }
      __syncwarp();
    }
  }
}

// ---- fp32 CUDA-core GEMM: C[M,N] (+)= op(A)[M,Kd] * op(B)[Kd,N], optional per-row output scale ----
//   TA: A is stored [Kd,M] (transposed).  TB: B is stored [N,Kd] (transposed).
//   gridDim.z > 1 splits Kd; partial tiles are then combined with atomicAdd (C zeroed by the caller).
constexpr int GM = 64, GN = 64, GK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_gemm_f32(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                  const float* __restrict__ row_scale, int64_t M, int N, int64_t Kd,
                                                  int64_t k_per_split) {
  __shared__ float sA[GK][GM + 4];
  __shared__ float sB[GK][GN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * GM;
  const int n0 = blockIdx.y * GN;
  const int64_t k_begin = (int64_t)blockIdx.z * k_per_split;
  const int64_t k_end = min(Kd, k_begin + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = k_begin; k0 < k_end; k0 += GK) {
    // load A tile: GM x GK elements, 256 threads x 4
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = threadIdx.x + 256 * r;  // 0..1023
      int mm, kk;
      if (TA) { mm = e % GM; kk = e / GM; } else { kk = e % GK; mm = e / GK; }
      const int64_t gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < k_end) v = TA ? A[gk * M + gm] : A[gm * Kd + gk];
      sA[kk][mm] = v;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = threadIdx.x + 256 * r;
      int nn, kk;
      if (TB) { kk = e % GK; nn = e / GK; } else { nn = e % GN; kk = e / GN; }
      const int gn = n0 + nn;
      const int64_t gk = k0 + kk;
      float v = 0.f;
      if (gn < N && gk < k_end) v = TB ? B[(int64_t)gn * Kd + gk] : B[gk * N + gn];
      sB[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
    const float sc = row_scale ? row_scale[gm] : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      if (gridDim.z > 1) atomicAdd(C + gm * N + gn, acc[i][j] * sc);
      else C[gm * N + gn] = acc[i][j] * sc;
    }
  }
}

__global__ void __launch_bounds__(256) k_scale_rows(const float* __restrict__ in, const float* __restrict__ scale, int64_t rows,
                                                    int cols, float* __restrict__ out) {
  const int64_t total = rows * (int64_t)cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i] * scale[i / cols];
}

template <bool TA, bool TB>
int launch_gemm(const float* A, const float* B, float* C, const float* row_scale, int64_t M, int N, int64_t Kd, int splits,
                cudaStream_t stream) {
  if (M == 0 || N == 0) return KPREG_OK;
  if (splits < 1) splits = 1;
  int64_t per = (Kd + splits - 1) / splits;
  per = (per + GK - 1) / GK * GK;
  splits = (int)((Kd + per - 1) / per);
  if (splits < 1) splits = 1;
  if (splits > 1) KP_CUDA_TRY(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * (size_t)N, stream));
  dim3 grid((unsigned)ceil_div(M, GM), (unsigned)ceil_div(N, GN), (unsigned)splits);
  k_gemm_f32<TA, TB><<<grid, 256, 0, stream>>>(A, B, C, row_scale, M, N, Kd, per);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

struct KpconvWs {
  unsigned char* row_pos; float* inv_num; float* agg; float* d_agg; float* g_scaled; float* w_split; void* tn_split; size_t total;
};

KpconvWs carve_kpconv(void* base, int64_t n_q, int64_t n_s, int n_kpts, int c_in, int c_out, int backward) {
  KpconvWs w;
  Carver cv(base);
  const size_t kd = (size_t)n_kpts * (size_t)c_in;
  w.row_pos = cv.take<unsigned char>((size_t)n_s + 1);
  w.inv_num = cv.take<float>((size_t)n_q + 1);
  w.agg = cv.take<float>((size_t)(n_q > 0 ? n_q : 1) * kd);
  w.d_agg = nullptr;
  w.g_scaled = nullptr;
  w.tn_split = nullptr;
  size_t w_bytes = kpconv_gemm_tc_weight_bytes((int)kd, c_out);
  if (backward) {
    w.d_agg = cv.take<float>((size_t)(n_q > 0 ? n_q : 1) * kd);
    w.g_scaled = cv.take<float>((size_t)(n_q > 0 ? n_q : 1) * (size_t)c_out);
    w.tn_split = cv.take<char>(gemm_tn_workspace_bytes(n_q, c_out));
    const size_t wt_bytes = kpconv_gemm_tc_weight_bytes(c_out, (int)kd);  // the weights as the operand of d_agg = g' W^T
    if (wt_bytes > w_bytes) w_bytes = wt_bytes;
  }
  w.w_split = reinterpret_cast<float*>(cv.take<char>(w_bytes));
  w.total = align_up(cv.used, 256);
  return w;
}

// The normalisation's row predicate (feature sum > 0) of the support rows: part of the aggregate step
void launch_row_pass(const float* x, int64_t n_s, int c_in, unsigned char* row_pos, cudaStream_t stream) {
  if (n_s == 0) return;
  if ((c_in & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int q = c_in / 4;
    if (q <= 4) k_row_positive_vec<4><<<ceil_div(n_s * 4, 256), 256, 0, stream>>>(x, n_s, c_in, row_pos);
    else if (q <= 8) k_row_positive_vec<8><<<ceil_div(n_s * 8, 256), 256, 0, stream>>>(x, n_s, c_in, row_pos);
    else if (q <= 16) k_row_positive_vec<16><<<ceil_div(n_s * 16, 256), 256, 0, stream>>>(x, n_s, c_in, row_pos);
    else k_row_positive_vec<32><<<ceil_div(n_s * 32, 256), 256, 0, stream>>>(x, n_s, c_in, row_pos);
  } else {
    k_row_positive<<<ceil_div(n_s * 32, 256), 256, 0, stream>>>(x, n_s, c_in, row_pos);
  }
  count_launches(1);
}

template <typename IdxT>
int launch_gather(const float* q_pts, const float* s_pts, const void* idx, const float* x, const unsigned char* row_pos,
                  const float* kp, int64_t n_q, int64_t n_s, int n_nbrs, int n_kpts, int c_in, float extent, int influence,
                  int aggregation, float* agg, float* inv_num, const int32_t* order, cudaStream_t stream) {
  const int blocks = gather_blocks(n_q);
  const IdxT* ip = static_cast<const IdxT*>(idx);
  ProfScope prof(KPREG_FAM_GATHER, stream);
  if (n_nbrs <= 64 && n_kpts <= 16 && c_in <= 256 && n_s < ((int64_t)1 << 31) && g_gather_mma) {
    // float4 path: whole rows of x and of the aggregate are 16-byte aligned
    const bool vec = (c_in % 4) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(agg)) & 15) == 0 && !g_gather_novec;
    if (c_in > 64 && vec && g_gather_groups) {
      // wide rows: (query, 64-channel group) work items on the 64-channel kernel (see GRP above)
      const int n_groups = (c_in + 63) / 64;
      const int gblocks = gather_blocks(n_q * n_groups);
      if (influence == 1)
        k_kpconv_gather_mma<IdxT, 8, true, 1, true><<<gblocks, kGatherWarps * 32, 0, stream>>>(
            q_pts, s_pts, ip, x, row_pos, kp, n_q, n_s, n_nbrs, n_kpts, c_in, extent, influence, aggregation, agg, inv_num, order, n_groups);
      else
        k_kpconv_gather_mma<IdxT, 8, true, -1, true><<<gblocks, kGatherWarps * 32, 0, stream>>>(
            q_pts, s_pts, ip, x, row_pos, kp, n_q, n_s, n_nbrs, n_kpts, c_in, extent, influence, aggregation, agg, inv_num, order, n_groups);
      KP_LAUNCH_CHECK();
      return KPREG_OK;
    }
#define KP_GATHER_MMA_(NT, VEC, INFL)                                                                                          \
  k_kpconv_gather_mma<IdxT, NT, VEC, INFL><<<blocks, kGatherWarps * 32, 0, stream>>>(q_pts, s_pts, ip, x, row_pos, kp, n_q,    \
                                                                                     n_s, n_nbrs, n_kpts, c_in, extent,        \
                                                                                     influence, aggregation, agg, inv_num, order, 1)
#define KP_GATHER_MMA(NT, VEC)                                                \
  do {                                                                        \
    if (influence == 1) KP_GATHER_MMA_(NT, VEC, 1); else KP_GATHER_MMA_(NT, VEC, -1); \
  } while (0)
    if (c_in <= 8) KP_GATHER_MMA(1, false);
    else if (c_in <= 16) KP_GATHER_MMA(2, false);
    else if (c_in <= 32) { if (vec) KP_GATHER_MMA(4, true); else KP_GATHER_MMA(4, false); }
    else if (c_in <= 64) { if (vec) KP_GATHER_MMA(8, true); else KP_GATHER_MMA(8, false); }
    else if (c_in <= 128) { if (vec) KP_GATHER_MMA(16, true); else KP_GATHER_MMA(16, false); }
    else { if (vec) KP_GATHER_MMA(32, true); else KP_GATHER_MMA(32, false); }
#undef KP_GATHER_MMA_
#undef KP_GATHER_MMA
    KP_LAUNCH_CHECK();
    return KPREG_OK;
  }
#define KP_GATHER(CPL)                                                                                                     \
  k_kpconv_gather<IdxT, CPL><<<blocks, kGatherWarps * 32, 0, stream>>>(q_pts, s_pts, ip, x, row_pos, kp, n_q, n_s, n_nbrs, \
                                                                       n_kpts, c_in, extent, influence, aggregation, agg,  \
                                                                       inv_num, order)
  if (c_in <= 32) KP_GATHER(1);
  else if (c_in <= 64) KP_GATHER(2);
  else if (c_in <= 128) KP_GATHER(4);
  else if (c_in <= 256) KP_GATHER(8);
  else return KPREG_E_INVALID;
#undef KP_GATHER
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

template <typename IdxT>
int launch_scatter(const float* q_pts, const float* s_pts, const void* idx, const float* kp, const float* d_agg, int64_t n_q,
                   int64_t n_s, int n_nbrs, int n_kpts, int c_in, float extent, int influence, int aggregation, float* d_x,
                   const int32_t* order, cudaStream_t stream) {
  const int blocks = gather_blocks(n_q);
  const IdxT* ip = static_cast<const IdxT*>(idx);
#define KP_SCATTER(CPL)                                                                                                  \
  k_kpconv_scatter<IdxT, CPL><<<blocks, kGatherWarps * 32, 0, stream>>>(q_pts, s_pts, ip, kp, d_agg, n_q, n_s, n_nbrs,   \
                                                                        n_kpts, c_in, extent, influence, aggregation, d_x, order)
  if (c_in <= 32) KP_SCATTER(1);
  else if (c_in <= 64) KP_SCATTER(2);
  else if (c_in <= 128) KP_SCATTER(4);
  else if (c_in <= 256) KP_SCATTER(8);
  else return KPREG_E_INVALID;
#undef KP_SCATTER
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

int check_kpconv_args(int64_t n_q, int64_t n_s, int n_nbrs, int n_kpts, int c_in, int c_out, float extent, int influence,
                      int aggregation) {
  if (n_q < 0 || n_s < 0 || n_nbrs < 0 || n_kpts < 1 || n_kpts > KMAX || c_in < 1 || c_in > 256 || c_out < 1) return KPREG_E_INVALID;
  if (!(extent > 0.f) || influence < 0 || influence > 2 || aggregation < 0 || aggregation > 1) return KPREG_E_INVALID;
  return KPREG_OK;
}

}  // namespace
}  // namespace kpreg

using namespace kpreg;

namespace {
struct GatherModeInit {
  GatherModeInit() {
    const char* e = getenv("KPREG_GATHER_FFMA");
    if (e && e[0] == '1') kpreg::g_gather_mma = false;
    e = getenv("KPREG_GATHER_NOVEC");
    if (e && e[0] == '1') kpreg::g_gather_novec = true;
    e = getenv("KPREG_NO_C1");
    if (e && e[0] == '1') kpreg::g_no_c1 = true;
    e = getenv("KPREG_C1_BY_NEIGHBOUR");
    if (e && e[0] == '1') kpreg::g_c1_by_neighbour = true;
    e = getenv("KPREG_GATHER_CTAS_PER_SM");
    if (e && atoi(e) > 0) kpreg::g_gather_ctas_per_sm = atoi(e);
    e = getenv("KPREG_GATHER_GROUPS");
    if (e && e[0] == '0') kpreg::g_gather_groups = false;
    e = getenv("KPREG_GATHER_MIN_SLICE");
    if (e && atoi(e) > 0) kpreg::g_gather_min_slice = atoi(e);
  }
} g_gather_mode_init;
}  // namespace

extern "C" int kpreg_kpconv_workspace_bytes(int64_t n_q, int64_t n_s, int n_kpts, int c_in, int c_out, int backward,
                                            size_t* bytes) {
  if (!bytes || n_q < 0 || n_s < 0 || n_kpts < 1 || c_in < 1 || c_out < 1) return KPREG_E_INVALID;
  *bytes = carve_kpconv(nullptr, n_q, n_s, n_kpts, c_in, c_out, backward).total;
  return KPREG_OK;
}

extern "C" int kpreg_kpconv_forward(const float* q_pts, const float* s_pts, const void* idx, int idx64, const float* x,
                                    const float* weights, const float* kernel_points, int64_t n_q, int64_t n_s, int n_nbrs,
                                    int n_kpts, int c_in, int c_out, float kp_extent, int influence, int aggregation, int gemm,
                                    const int32_t* order, float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  return kpreg_kpconv_forward_rowpos(q_pts, s_pts, idx, idx64, x, weights, kernel_points, n_q, n_s, n_nbrs, n_kpts, c_in, c_out,
                                     kp_extent, influence, aggregation, gemm, order, nullptr, out, workspace, workspace_bytes, stream_);
}

extern "C" int kpreg_kpconv_forward_rowpos(const float* q_pts, const float* s_pts, const void* idx, int idx64, const float* x,
                                           const float* weights, const float* kernel_points, int64_t n_q, int64_t n_s, int n_nbrs,
                                           int n_kpts, int c_in, int c_out, float kp_extent, int influence, int aggregation,
                                           int gemm, const int32_t* order, const unsigned char* row_pos_in, float* out,
                                           void* workspace, size_t workspace_bytes, void* stream_) {
  int rc = check_kpconv_args(n_q, n_s, n_nbrs, n_kpts, c_in, c_out, kp_extent, influence, aggregation);
  if (rc) return rc;
  if (n_q == 0) return KPREG_OK;
  if (!q_pts || !weights || !kernel_points || !out || !workspace) return KPREG_E_INVALID;
  if (n_s > 0 && (!s_pts || !x)) return KPREG_E_INVALID;
  if (n_nbrs > 0 && !idx) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (c_in == 1 && c_out <= 128 && n_nbrs <= 64 && gemm != 2 && n_s > 0 && !g_no_c1) {
    // single input channel: gather, contraction and normalisation in one kernel (raw [K, 1, c_out] weights)
    const int blocks = gather_blocks(n_q);
    ProfScope prof(KPREG_FAM_GATHER, stream);
#define KP_C1_(IdxT, CPL, INFL)                                                                                                       \
  k_kpconv_c1<IdxT, CPL, INFL><<<blocks, kGatherWarps * 32, 0, stream>>>(q_pts, s_pts, static_cast<const IdxT*>(idx), x,               \
                                                                         kernel_points, weights, n_q, n_s, n_nbrs, n_kpts, c_out,      \
                                                                         kp_extent, influence, aggregation, out, order)
#define KP_C1K_(IdxT, CPL, INFL)                                                                                                      \
  k_kpconv_c1_kp<IdxT, CPL, INFL><<<blocks, kGatherWarps * 32, 0, stream>>>(q_pts, s_pts, static_cast<const IdxT*>(idx), x,            \
                                                                            kernel_points, weights, n_q, n_s, n_nbrs, n_kpts, c_out,   \
                                                                            kp_extent, influence, out, order)
#define KP_C1(IdxT, CPL)                                                                      \
  do {                                                                                        \
    if (aggregation == 0 && !g_c1_by_neighbour) {                                             \
      if (influence == 1) KP_C1K_(IdxT, CPL, 1); else KP_C1K_(IdxT, CPL, -1);                 \
    } else {                                                                                  \
      if (influence == 1) KP_C1_(IdxT, CPL, 1); else KP_C1_(IdxT, CPL, -1);                   \
    }                                                                                         \
  } while (0)
    if (idx64) { if (c_out <= 32) KP_C1(int64_t, 1); else if (c_out <= 64) KP_C1(int64_t, 2); else KP_C1(int64_t, 4); }
    else { if (c_out <= 32) KP_C1(int32_t, 1); else if (c_out <= 64) KP_C1(int32_t, 2); else KP_C1(int32_t, 4); }
#undef KP_C1_
#undef KP_C1K_
#undef KP_C1
    KP_LAUNCH_CHECK();
    return KPREG_OK;
  }
  KpconvWs w = carve_kpconv(workspace, n_q, n_s, n_kpts, c_in, c_out, 0);
  if (w.total > workspace_bytes) return KPREG_E_WORKSPACE;
  const unsigned char* row_pos = row_pos_in ? row_pos_in : w.row_pos;
  if (!row_pos_in) {
    ProfScope prof_rows(KPREG_FAM_GATHER, stream);  // the row pass belongs to the aggregate step
    launch_row_pass(x, n_s, c_in, w.row_pos, stream);
    KP_CUDA_TRY(cudaPeekAtLastError());
  }
  rc = idx64 ? launch_gather<int64_t>(q_pts, s_pts, idx, x, row_pos, kernel_points, n_q, n_s, n_nbrs, n_kpts, c_in, kp_extent,
                                      influence, aggregation, w.agg, w.inv_num, order, stream)
             : launch_gather<int32_t>(q_pts, s_pts, idx, x, row_pos, kernel_points, n_q, n_s, n_nbrs, n_kpts, c_in, kp_extent,
                                      influence, aggregation, w.agg, w.inv_num, order, stream);
  if (rc) return rc;
  const int kd = n_kpts * c_in;
  ProfScope prof(KPREG_FAM_CONTRACT, stream);
  if (gemm == 2) {
    // `weights` already is the split operand pair produced by kpreg_split_weights(transpose = 1)
    if (!gemm_tc_supported(n_q, kd, c_out, kd, w.agg)) return KPREG_E_INVALID;
    return launch_gemm_tc(w.agg, kd, weights, out, c_out, n_q, kd, c_out, w.inv_num, nullptr, nullptr, nullptr, 0, 0, 0.f,
                          nullptr, 0, nullptr, 0, nullptr, 0, 0, stream);
  }
  if (gemm == 1 && gemm_tc_supported(n_q, kd, c_out, kd, w.agg)) {
    rc = kpconv_gemm_tc_prepare_weights(weights, kd, c_out, 1, w.w_split, stream);
    if (rc) return rc;
    return launch_gemm_tc(w.agg, kd, w.w_split, out, c_out, n_q, kd, c_out, w.inv_num, nullptr, nullptr, nullptr, 0, 0, 0.f,
                          nullptr, 0, nullptr, 0, nullptr, 0, 0, stream);
  }
  return launch_gemm<false, false>(w.agg, weights, out, w.inv_num, n_q, c_out, kd, 1, stream);
}

extern "C" int kpreg_kpconv_backward(const float* q_pts, const float* s_pts, const void* idx, int idx64, const float* x,
                                     const float* weights, const float* kernel_points, const float* grad_out, int64_t n_q,
                                     int64_t n_s, int n_nbrs, int n_kpts, int c_in, int c_out, float kp_extent, int influence,
                                     int aggregation, const int32_t* order, float* d_x, float* d_weights, void* workspace,
                                     size_t workspace_bytes, void* stream_) {
  int rc = check_kpconv_args(n_q, n_s, n_nbrs, n_kpts, c_in, c_out, kp_extent, influence, aggregation);
  if (rc) return rc;
  if (!d_weights || !weights || !kernel_points || !workspace) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int kd = n_kpts * c_in;
  KP_CUDA_TRY(cudaMemsetAsync(d_weights, 0, sizeof(float) * (size_t)kd * (size_t)c_out, stream));
  if (n_s > 0) {
    if (!d_x) return KPREG_E_INVALID;
    KP_CUDA_TRY(cudaMemsetAsync(d_x, 0, sizeof(float) * (size_t)n_s * (size_t)c_in, stream));
  }
  if (n_q == 0 || n_s == 0) return KPREG_OK;
  if (!q_pts || !s_pts || !x || !grad_out || (n_nbrs > 0 && !idx)) return KPREG_E_INVALID;
  KpconvWs w = carve_kpconv(workspace, n_q, n_s, n_kpts, c_in, c_out, 1);
  if (w.total > workspace_bytes) return KPREG_E_WORKSPACE;
  // recompute the aggregate and the normalisation (cheaper than keeping [n_q, K*c_in] alive per layer)
  launch_row_pass(x, n_s, c_in, w.row_pos, stream);
  KP_CUDA_TRY(cudaPeekAtLastError());
  rc = idx64 ? launch_gather<int64_t>(q_pts, s_pts, idx, x, w.row_pos, kernel_points, n_q, n_s, n_nbrs, n_kpts, c_in, kp_extent,
                                      influence, aggregation, w.agg, w.inv_num, order, stream)
             : launch_gather<int32_t>(q_pts, s_pts, idx, x, w.row_pos, kernel_points, n_q, n_s, n_nbrs, n_kpts, c_in, kp_extent,
                                      influence, aggregation, w.agg, w.inv_num, order, stream);
  if (rc) return rc;
  // g' = grad_out / num
  {
    int blocks = ceil_div(n_q * (int64_t)c_out, 256);
    if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
    k_scale_rows<<<blocks, 256, 0, stream>>>(grad_out, w.inv_num, n_q, c_out, w.g_scaled);
    KP_LAUNCH_CHECK();
  }
  static const bool bwd_fp32 = [] { const char* e = getenv("KPREG_BACKWARD_FP32"); return e && e[0] == '1'; }();  // A/B measurements
  if (!bwd_fp32 && gemm_tn_supported(n_q, kd, c_out, kd, w.agg) && gemm_tc_supported(n_q, c_out, kd, c_out, w.g_scaled)) {
    // both contractions of the backward pass on tcgen05 (3xTF32):
    //   d_weights[K*c_in, c_out] += agg^T (grad_out / num)     split over the queries, operands taken MN-major as they lie
    //   d_agg[n_q, K*c_in]        = (grad_out / num) W^T       the forward kernel with W^T as its K-major operand
    ProfScope prof(KPREG_FAM_CONTRACT, stream);
    rc = launch_gemm_tn(w.agg, kd, grad_out, c_out, w.inv_num, d_weights, c_out, n_q, kd, c_out, 0, w.tn_split, stream);
    if (rc) return rc;
    rc = kpconv_gemm_tc_prepare_weights(weights, c_out, kd, 0, w.w_split, stream);  // W as [n = K*c_in, k = c_out]
    if (rc) return rc;
    rc = launch_gemm_tc(w.g_scaled, c_out, w.w_split, w.d_agg, kd, n_q, c_out, kd, nullptr, nullptr, nullptr, nullptr, 0, 0, 0.f,
                        nullptr, 0, nullptr, 0, nullptr, 0, 0, stream);
    if (rc) return rc;
    return idx64 ? launch_scatter<int64_t>(q_pts, s_pts, idx, kernel_points, w.d_agg, n_q, n_s, n_nbrs, n_kpts, c_in, kp_extent,
                                           influence, aggregation, d_x, order, stream)
                 : launch_scatter<int32_t>(q_pts, s_pts, idx, kernel_points, w.d_agg, n_q, n_s, n_nbrs, n_kpts, c_in, kp_extent,
                                           influence, aggregation, d_x, order, stream);
  }
  // d_weights[K*c_in, c_out] = agg^T g'   (reduction over the queries, split across CTAs)
  {
    const int tiles = ceil_div(kd, GM) * ceil_div(c_out, GN);
    int splits = (2 * kNumSMs + tiles - 1) / tiles;
    const int max_splits = ceil_div(n_q, 4 * GK);
    if (splits > max_splits) splits = max_splits;
    rc = launch_gemm<true, false>(w.agg, w.g_scaled, d_weights, nullptr, kd, c_out, n_q, splits, stream);
    if (rc) return rc;
  }
  // d_agg[n_q, K*c_in] = g' W^T
  rc = launch_gemm<false, true>(w.g_scaled, weights, w.d_agg, nullptr, n_q, kd, c_out, 1, stream);
  if (rc) return rc;
  return idx64 ? launch_scatter<int64_t>(q_pts, s_pts, idx, kernel_points, w.d_agg, n_q, n_s, n_nbrs, n_kpts, c_in, kp_extent,
                                         influence, aggregation, d_x, order, stream)
               : launch_scatter<int32_t>(q_pts, s_pts, idx, kernel_points, w.d_agg, n_q, n_s, n_nbrs, n_kpts, c_in, kp_extent,
                                         influence, aggregation, d_x, order, stream);
}
