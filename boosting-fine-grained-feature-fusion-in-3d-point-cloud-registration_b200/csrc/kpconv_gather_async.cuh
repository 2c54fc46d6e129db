// EXPERIMENTAL — compiled only with -DKPREG_EXPERIMENTAL_GATHER_ASYNC (make EXTRA=-DKPREG_EXPERIMENTAL_GATHER_ASYNC) and
// selected at run time with KPREG_GATHER_ASYNC=1.  NOT validated on hardware yet: written at the end of round 1 from the
// stall profile of k_kpconv_gather_mma (profiles/r1f_ncu_full_summary.md), after the round's GPU budget was spent.
// It is not part of the default libkpreg_b200.so.
//
// Same operator, same fragment bindings and arithmetic as k_kpconv_gather_mma<.., VEC = true, ..> (kpconv.cu); the only
// change is how the feature rows reach the B fragments.  There, every lane loads its float4s for k-step s and the warp
// waits on them (long_scoreboard = 35 % of the stall samples at 4-5 warps per scheduler; holding the next k-step in
// registers cost a CTA of occupancy and was slower).  Here each warp owns a 3-stage ring in shared memory; the eight
// rows of k-step s + 2 are requested with cp.async (16-byte chunks, zero-fill for shadow neighbours) while k-step s is
// computed, and the B fragments are read back with LDS.128.  A chunk (row r, 16-byte column c) lives at column
// c ^ swz(r) of its row; swz toggles the two low column bits that do NOT separate a lane's two partner chunks
// (NT = 4: bits 1,2; NT = 8: bits 0,2; NT = 16: bits 0,1), which makes every quarter-warp LDS.128 conflict-free.
//
// Included by kpconv.cu inside namespace kpreg::{anonymous}, after mma_tf32 / tf32_split / influence_one.
#pragma once

constexpr int kRingStages = 3;

__device__ __forceinline__ void cp_async_16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int src_bytes = valid ? 16 : 0;  // src-size 0: nothing is read, the 16 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT>
__device__ __forceinline__ int ring_swizzle(int row) {
  static_assert(NT == 4 || NT == 8 || NT == 16, "ring swizzle is derived for 4, 8 and 16 channel tiles");
  const int r = row & 3;
  if (NT == 4) return r << 1;                  // bits 1, 2
  if (NT == 8) return (r & 1) | ((r & 2) << 1);  // bits 0, 2
  return r;                                    // bits 0, 1
}

template <typename IdxT, int NT, int INFL>
__global__ void __launch_bounds__(kGatherWarps * 32, gather_min_blocks(NT)) k_kpconv_gather_async(
    const float* __restrict__ q_pts, const float* __restrict__ s_pts, const IdxT* __restrict__ idx, const float* __restrict__ x,
    const unsigned char* __restrict__ row_pos, const float* __restrict__ kernel_points, int64_t n_q, int64_t n_s, int n_nbrs,
    int n_kpts, int c_in, float extent, int influence, int aggregation, float* __restrict__ agg, float* __restrict__ inv_num,
    const int32_t* __restrict__ order) {
  constexpr int kRowChunks = 2 * NT;                 // 16-byte chunks per feature row (8 NT floats)
  constexpr int kStageChunks = 8 * kRowChunks;       // eight neighbour rows per k-step
  extern __shared__ float4 s_ring_all[];             // [kGatherWarps][kRingStages][kStageChunks]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* ring = s_ring_all + (size_t)warp * kRingStages * kStageChunks;
  const int g = lane >> 2, t = lane & 3;
  const float inv_extent = 1.0f / extent;
  const bool k0_ok = g < n_kpts, k1_ok = g + 8 < n_kpts;
  const float k0x = k0_ok ? kernel_points[3 * g] : 0.f, k0y = k0_ok ? kernel_points[3 * g + 1] : 0.f,
              k0z = k0_ok ? kernel_points[3 * g + 2] : 0.f;
  const float k1x = k1_ok ? kernel_points[3 * (g + 8)] : 0.f, k1y = k1_ok ? kernel_points[3 * (g + 8) + 1] : 0.f,
              k1z = k1_ok ? kernel_points[3 * (g + 8) + 2] : 0.f;
  const int n_s32 = (int)n_s;
  const int n_steps = (n_nbrs + 7) >> 3;
  const int row_chunks_valid = c_in >> 2;            // chunks of a row that exist in x (c_in is a multiple of 4)

  const int64_t per_cta = (n_q + gridDim.x - 1) / gridDim.x;
  const int64_t it_end = min(n_q, (int64_t)(blockIdx.x + 1) * per_cta);
  int64_t it = (int64_t)blockIdx.x * per_cta + warp;
  int64_t n_nx = 0;
  IdxT raw_nx[2] = {(IdxT)n_s32, (IdxT)n_s32};
  if (it < it_end) {
    n_nx = order ? (int64_t)order[it] : it;
#pragma unroll
    for (int r = 0; r < 2; ++r)
      if (lane + 32 * r < n_nbrs) raw_nx[r] = idx[n_nx * n_nbrs + lane + 32 * r];
  }
  for (; it < it_end; it += kGatherWarps) {
    const int64_t n = n_nx;
    const IdxT raw[2] = {raw_nx[0], raw_nx[1]};
    if (it + kGatherWarps < it_end) {
      n_nx = order ? (int64_t)order[it + kGatherWarps] : it + kGatherWarps;
#pragma unroll
      for (int r = 0; r < 2; ++r) raw_nx[r] = lane + 32 * r < n_nbrs ? idx[n_nx * n_nbrs + lane + 32 * r] : (IdxT)n_s32;
    }
    int jn[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) jn[r] = (raw[r] >= 0 && raw[r] < (IdxT)n_s32) ? (int)raw[r] : -1;

    // request the eight feature rows of k-step `step` into ring slot step % kRingStages (all lanes, kStageChunks / 32 chunks each)
    auto request_step = [&](int step) {
      float4* slot = ring + (step % kRingStages) * kStageChunks;
      const int h0 = step << 3;
      const int jsel = (h0 >> 5) ? jn[1] : jn[0];     // h0 .. h0 + 7 lie in one 32-neighbour block: warp-uniform select
#pragma unroll
      for (int sidx = 0; sidx < kStageChunks / 32; ++sidx) {
        const int i = lane + 32 * sidx;
        const int r = i / kRowChunks, c = i % kRowChunks;
        const int j = __shfl_sync(0xffffffffu, jsel, (h0 + r) & 31);
        const bool ok = j >= 0 && c < row_chunks_valid && h0 + r < n_nbrs;
        const float* src = x + (int64_t)(ok ? j : 0) * c_in + 4 * (ok ? c : 0);
        cp_async_16_zfill(slot + r * kRowChunks + (c ^ ring_swizzle<NT>(r)), src, ok);
      }
    };
    __syncwarp();  // every lane has finished reading the ring for the previous query
    request_step(0);
    cp_async_commit();
    if (n_steps > 1) request_step(1);
    cp_async_commit();

    // relative positions and the normaliser (these loads overlap the row requests above)
    const float qx = q_pts[3 * n], qy = q_pts[3 * n + 1], qz = q_pts[3 * n + 2];
    float rx[2], ry[2], rz[2];
    int num = 0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int j = jn[r];
      const bool valid = j >= 0;
      rx[r] = ry[r] = rz[r] = 0.f;
      if (valid) { rx[r] = s_pts[3 * (int64_t)j] - qx; ry[r] = s_pts[3 * (int64_t)j + 1] - qy; rz[r] = s_pts[3 * (int64_t)j + 2] - qz; }
      num += __popc(__ballot_sync(0xffffffffu, valid && row_pos[j] != 0));
    }
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    for (int step = 0; step < n_steps; ++step) {
      const int h0 = step << 3;
      if (step + 2 < n_steps) {
        __syncwarp();  // slot (step + 2) % 3 was read in iteration step - 1: nobody may still be reading it
        request_step(step + 2);
      }
      cp_async_commit();   // one group per iteration (possibly empty) keeps the pending-group count uniform
      cp_async_wait<2>();  // groups step + 1 and step + 2 may be in flight; group `step` has landed for this lane
      __syncwarp();        // ... and for every other lane of the warp
      const int ha = h0 + t, hb = h0 + t + 4;
      const int ra = ha >> 5, rb = hb >> 5;
      const int ja = __shfl_sync(0xffffffffu, ra ? jn[1] : jn[0], ha & 31);
      const int jb = __shfl_sync(0xffffffffu, rb ? jn[1] : jn[0], hb & 31);
      const bool va = ja >= 0, vb = jb >= 0;
      if (!__any_sync(0xffffffffu, va || vb)) continue;  // eight shadow neighbours: nothing to add
      const float ax = __shfl_sync(0xffffffffu, ra ? rx[1] : rx[0], ha & 31), ay = __shfl_sync(0xffffffffu, ra ? ry[1] : ry[0], ha & 31),
                  az = __shfl_sync(0xffffffffu, ra ? rz[1] : rz[0], ha & 31);
      const float bx = __shfl_sync(0xffffffffu, rb ? rx[1] : rx[0], hb & 31), by = __shfl_sync(0xffffffffu, rb ? ry[1] : ry[0], hb & 31),
                  bz = __shfl_sync(0xffffffffu, rb ? rz[1] : rz[0], hb & 31);
      float w[4], d2[4];
      w[0] = influence_one<INFL>(ax, ay, az, k0x, k0y, k0z, inv_extent, extent, influence, d2[0]);
      w[1] = influence_one<INFL>(ax, ay, az, k1x, k1y, k1z, inv_extent, extent, influence, d2[1]);
      w[2] = influence_one<INFL>(bx, by, bz, k0x, k0y, k0z, inv_extent, extent, influence, d2[2]);
      w[3] = influence_one<INFL>(bx, by, bz, k1x, k1y, k1z, inv_extent, extent, influence, d2[3]);
      if (!k0_ok) { w[0] = w[2] = 0.f; d2[0] = d2[2] = 3.4e38f; }
      if (!k1_ok) { w[1] = w[3] = 0.f; d2[1] = d2[3] = 3.4e38f; }
      if (!va) w[0] = w[1] = 0.f;
      if (!vb) w[2] = w[3] = 0.f;
      if (aggregation == 1) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float best = d2[2 * hh];
          int best_k = g;
          if (d2[2 * hh + 1] < best) { best = d2[2 * hh + 1]; best_k = g + 8; }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
            if (ob < best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
          }
          if (best_k != g) w[2 * hh] = 0.f;
          if (best_k != g + 8) w[2 * hh + 1] = 0.f;
        }
      }
      float a_hi[4], a_lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) tf32_split(w[i], a_hi[i], a_lo[i]);
      // B fragments from the ring: rows t and t + 4 of the slot, channels NT g + 4 m .. + 3 = chunk (NT / 4) g + m
      const float4* slot = ring + (step % kRingStages) * kStageChunks;
      const int swz = ring_swizzle<NT>(t);  // rows t and t + 4 share it
#pragma unroll
      for (int m = 0; m < NT / 4; ++m) {
        const int c = (NT / 4) * g + m;
        const float4 fa = slot[t * kRowChunks + (c ^ swz)];
        const float4 fb = slot[(t + 4) * kRowChunks + (c ^ swz)];
        const float av[4] = {fa.x, fa.y, fa.z, fa.w}, bv[4] = {fb.x, fb.y, fb.z, fb.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float b0h, b0l, b1h, b1l;
          tf32_split(av[e], b0h, b0l);
          tf32_split(bv[e], b1h, b1l);
          mma_tf32(acc[4 * m + e], a_lo, b0h, b1h);
          mma_tf32(acc[4 * m + e], a_hi, b0l, b1l);
          mma_tf32(acc[4 * m + e], a_hi, b0h, b1h);
        }
      }
    }
    cp_async_wait<0>();  // nothing of this query is left in flight before the ring is reused
    float* __restrict__ arow = agg + n * (int64_t)n_kpts * c_in;
#pragma unroll
    for (int m = 0; m < NT / 4; ++m) {
      const int ce = 2 * t * NT + 4 * m, co = (2 * t + 1) * NT + 4 * m;
      if (ce < c_in) {
        if (k0_ok) *reinterpret_cast<float4*>(arow + g * c_in + ce) = make_float4(acc[4 * m][0], acc[4 * m + 1][0], acc[4 * m + 2][0], acc[4 * m + 3][0]);
        if (k1_ok) *reinterpret_cast<float4*>(arow + (g + 8) * c_in + ce) = make_float4(acc[4 * m][2], acc[4 * m + 1][2], acc[4 * m + 2][2], acc[4 * m + 3][2]);
      }
      if (co < c_in) {
        if (k0_ok) *reinterpret_cast<float4*>(arow + g * c_in + co) = make_float4(acc[4 * m][1], acc[4 * m + 1][1], acc[4 * m + 2][1], acc[4 * m + 3][1]);
        if (k1_ok) *reinterpret_cast<float4*>(arow + (g + 8) * c_in + co) = make_float4(acc[4 * m][3], acc[4 * m + 1][3], acc[4 * m + 2][3], acc[4 * m + 3][3]);
      }
    }
    if (lane == 0) inv_num[n] = 1.0f / (float)max(num, 1);
  }
}
