// Radius neighbours of a stacked batch on sm_100a: uniform-grid cell binning + warp-per-query sweep.
//
// Replaces batch_nanoflann_neighbors() (reference cpp_neighbors/neighbors/neighbors.cpp:211-332).
// The reference's KD-tree is only an accelerator; what defines the result is
//   - the metric   d2 = (dx*dx + dy*dy) + dz*dz in fp32 without FMA   (nanoflann.hpp:432-440)
//   - the test     d2 < r*r, strict                                    (nanoflann.hpp:249-253)
//   - the order    ascending d2 (nanoflann.hpp:1280-1289); ties by index here
//   - the row      local index + cloud offset, padded with n_supports (neighbors.cpp:304-327)
// and those are reproduced bit for bit.
//
// Grid build (once per support set):
//   cell key = cloud<<42 | cz<<28 | cy<<14 | cx, coordinates relative to the batch bounding box,
//   computed in fp64 so that a neighbour within r can never be more than one cell away;
//   supports are radix-sorted by key into a float4 array (xyz + local index); every occupied cell
//   gets a [start,end) range in an open-addressing hash table (64-bit keys, linear probing).
// Query, hot path (k_grid_query_tq, row width <= 56): ONE THREAD PER QUERY, 32 consecutive queries of the cell-sorted
// processing order per warp.
//   * The warp's queries fall into a handful of x-adjacent cells.  Lanes whose cells lie within a small window of the
//     first pending lane's cell form a group; the group's candidate set is the bounding box of its cells grown by one
//     cell.  Because cx is the lowest key field, the occupied cells of one (cloud, cz, cy) row of the box are ONE
//     contiguous span of the sorted support array: 9 .. 25 spans per group instead of 27 hash look-ups + a per-candidate
//     binary search per query.  (Cells are probed once per group, lanes in parallel; span ends land in shared memory.)
//   * Every lane then tests every candidate of the box against its own query: the candidate's float4 is one
//     warp-uniform load (a broadcast, served by L1 — the spans are shared by the CTA's warps and by consecutive
//     iterations), ~12 instructions per candidate for 32 queries at once.  A warp per query spent ~50 instructions per
//     32 candidates on ONE query: the old kernel was instruction-issue bound at 2 % of the HBM roofline.
//   * Hits are appended to the lane's own list in shared memory ([lane][CAP+1] 64-bit (d2 bits << 32 | index) keys —
//     the odd pitch makes both the per-lane and the per-row access patterns conflict-free), sorted per lane by
//     insertion (32 lists in lock step), and the rows are written out cooperatively: one row = one contiguous,
//     coalesced store of the whole warp.
//   * A lane with more than CAP (= 64) hits is finished by the warp-cooperative exact routine below (exact_query_warp).
// Query, general path (k_grid_query, any width, also the count-only pass): one warp per query through
// exact_query_warp: lanes 0..26 look up the 27 surrounding cells; the warp then walks the concatenated ranges 32
//   candidates at a time (coalesced 16-byte loads), compacts the hits with a ballot into shared
//   memory as 64-bit (d2 bits << 32 | index) keys, ranks them by counting and writes each index at its rank
//   (a row is one contiguous 4*width-byte span, so the scattered 4-byte stores of a warp fall into 1-2 lines).
#include <cub/cub.cuh>

#include <cstdlib>

#include "common.cuh"

namespace kpreg {
namespace {

constexpr int kCellBits = 14;
constexpr int kCellMax = (1 << kCellBits) - 1;
constexpr int kCloudShift = 3 * kCellBits;
constexpr uint64_t kEmptyKey = ~0ull;
constexpr int kWarpsPerBlock = 8;
constexpr int kHitCap = 256;  // hits buffered per query in shared memory; beyond: recount path

struct GridHeader {  // lives at the start of the grid workspace (device memory)
  double min[3];
  double inv_cell;
  int32_t dims[3];
  int32_t status;
};

// one hash-table slot: the cell key and its [start, end) range of the sorted array travel in ONE 16-byte load
struct __align__(16) Slot {
  unsigned long long key;
  int2 val;
};

struct GridWs {
  GridHeader* hdr; int64_t* off; int64_t* q_off; unsigned int* bbox;
  uint64_t* keys0; uint64_t* keys1; uint32_t* idx0; uint32_t* idx1;
  float4* sorted; Slot* tab;
  void* cub_tmp; size_t cub_tmp_bytes; uint32_t tab_cap; size_t total;
  int64_t n; int n_clouds;
};

GridWs carve_grid(void* base, int64_t n, int n_clouds) {
  GridWs w;
  Carver cv(base);
  const size_t np = (size_t)(n > 0 ? n : 1);
  w.hdr = cv.take<GridHeader>(1);
  w.off = cv.take<int64_t>((size_t)n_clouds + 1);
  w.q_off = cv.take<int64_t>((size_t)n_clouds + 1);
  w.bbox = cv.take<unsigned int>(8);
  w.keys0 = cv.take<uint64_t>(np);
  w.keys1 = cv.take<uint64_t>(np);
  w.idx0 = cv.take<uint32_t>(np);
  w.idx1 = cv.take<uint32_t>(np);
  w.sorted = cv.take<float4>(np);
  uint32_t cap = 64;
  while ((size_t)cap < 2 * np) cap <<= 1;
  w.tab_cap = cap;
  w.tab = cv.take<Slot>(cap);
  w.cub_tmp_bytes = (size_t)(8u << 20) + np * 16;
  w.cub_tmp = cv.take<char>(w.cub_tmp_bytes);
  w.total = align_up(cv.used, 256);
  w.n = n;
  w.n_clouds = n_clouds;
  return w;
}

__device__ __forceinline__ uint32_t hash_key(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (uint32_t)k;
}

__global__ void k_grid_init(unsigned int* __restrict__ bbox, GridHeader* __restrict__ hdr) {
  if (threadIdx.x < 6) bbox[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
  if (threadIdx.x == 0) hdr->status = 0;
}

__global__ void __launch_bounds__(256) k_grid_bbox(const float* __restrict__ pts, int64_t n, unsigned int* __restrict__ bbox) {
  float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float v = pts[3 * i + d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  if (lane_id() == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      atomicMin(bbox + d, float_to_ordered(mn[d]));
      atomicMax(bbox + 3 + d, float_to_ordered(mx[d]));
    }
  }
}

// The cell edge is `cell` widened by 1e-6 (fp32 rounding of d2 and r*r can never admit a point that
// is a full widened cell away), and widened further if the batch extent would overflow 14 bits/axis.
__global__ void k_grid_header(const unsigned int* __restrict__ bbox, float cell, GridHeader* __restrict__ hdr) {
  if (threadIdx.x != 0) return;
  double edge = (double)cell * (1.0 + 1e-6);
  double ext = 0.0;
  for (int d = 0; d < 3; ++d) {
    double mn = (double)ordered_to_float(bbox[d]);
    double mx = (double)ordered_to_float(bbox[3 + d]);
    hdr->min[d] = mn;
    ext = fmax(ext, mx - mn);
  }
  if (!(ext / edge < (double)(kCellMax - 1))) edge = ext / (double)(kCellMax - 2);
  if (!(edge > 0.0) || !isfinite(edge)) { edge = 1.0; hdr->status = KPREG_E_RANGE; }
  hdr->inv_cell = 1.0 / edge;
  for (int d = 0; d < 3; ++d) {
    double mx = (double)ordered_to_float(bbox[3 + d]);
    hdr->dims[d] = (int)floor((mx - hdr->min[d]) * hdr->inv_cell) + 1;
  }
}

__device__ __forceinline__ void cell_coords(const GridHeader& h, float x, float y, float z, int& cx, int& cy, int& cz) {
  cx = (int)floor(((double)x - h.min[0]) * h.inv_cell);
  cy = (int)floor(((double)y - h.min[1]) * h.inv_cell);
  cz = (int)floor(((double)z - h.min[2]) * h.inv_cell);
}

__device__ __forceinline__ uint64_t make_key(int cloud, int cx, int cy, int cz) {
  return ((uint64_t)cloud << kCloudShift) | ((uint64_t)cz << (2 * kCellBits)) | ((uint64_t)cy << kCellBits) | (uint64_t)cx;
}

__global__ void __launch_bounds__(256) k_cell_keys(const float* __restrict__ pts, const int64_t* __restrict__ off, int n_clouds,
                                                   int64_t n, const GridHeader* __restrict__ hdr, uint64_t* __restrict__ keys,
                                                   uint32_t* __restrict__ idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const GridHeader h = *hdr;
  int c = cloud_of(off, n_clouds, i);
  int cx, cy, cz;
  cell_coords(h, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], cx, cy, cz);
  cx = min(max(cx, 0), kCellMax);
  cy = min(max(cy, 0), kCellMax);
  cz = min(max(cz, 0), kCellMax);
  keys[i] = make_key(c, cx, cy, cz);
  idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_table_clear(Slot* __restrict__ tab, uint32_t cap) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) tab[i].key = kEmptyKey;
}

// Gather the sorted float4 array (xyz + LOCAL index bits) and register every cell in the hash table.
__global__ void __launch_bounds__(256) k_grid_fill(const float* __restrict__ pts, const int64_t* __restrict__ off,
                                                   const uint64_t* __restrict__ keys_sorted, const uint32_t* __restrict__ idx_sorted,
                                                   int64_t n, float4* __restrict__ sorted, Slot* __restrict__ tab, uint32_t cap) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint64_t key = keys_sorted[p];
  const uint32_t i = idx_sorted[p];
  const int cloud = (int)(key >> kCloudShift);
  sorted[p] = make_float4(pts[3 * (int64_t)i], pts[3 * (int64_t)i + 1], pts[3 * (int64_t)i + 2],
                          __int_as_float((int)((int64_t)i - off[cloud])));
  if (p > 0 && keys_sorted[p - 1] == key) return;
  int64_t e = p + 1;
  while (e < n && keys_sorted[e] == key) ++e;
  uint32_t slot = hash_key(key) & (cap - 1);
  while (true) {
    unsigned long long prev = atomicCAS(&tab[slot].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (prev == kEmptyKey) { tab[slot].val = make_int2((int)p, (int)e); break; }
    slot = (slot + 1) & (cap - 1);
  }
}

__device__ __forceinline__ int2 table_find(const Slot* __restrict__ tab, uint32_t cap, uint64_t key) {
  uint32_t slot = hash_key(key) & (cap - 1);
  while (true) {
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(tab + slot));
    const uint64_t k = ((uint64_t)e.y << 32) | e.x;
    if (k == key) return make_int2((int)e.z, (int)e.w);
    if (k == kEmptyKey) return make_int2(0, 0);
    slot = (slot + 1) & (cap - 1);
  }
}

__device__ __forceinline__ float dist2_ref(float qx, float qy, float qz, float sx, float sy, float sz) {
  // nanoflann L2_Simple_Adaptor::evalMetric: result = 0; result += diff*diff for x, y, z — no FMA.
  const float dx = __fsub_rn(qx, sx), dy = __fsub_rn(qy, sy), dz = __fsub_rn(qz, sz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// One query, the whole warp: the exact neighbour row of (qx,qy,qz) for ANY hit count.  All arguments are warp-uniform.
// `hits` is a per-warp shared-memory buffer of hit_cap keys, s_cell_start / s_cell_prefix hold 32 ints each.
// Writes the row (indices + padding) and returns the number of supports within the radius.
template <typename OutT>
__device__ __forceinline__ int exact_query_warp(const GridHeader& h, const float4* __restrict__ sorted,
                                                const Slot* __restrict__ tab, uint32_t cap,
                                                float qx, float qy, float qz, int cloud, int cx, int cy, int cz, float r2, int width,
                                                OutT* __restrict__ row, int64_t cloud_base, int64_t n_supports,
                                                unsigned long long* __restrict__ hits, int hit_cap, int* __restrict__ s_cell_start,
                                                int* __restrict__ s_cell_prefix) {
  const int lane = threadIdx.x & 31;
  // lanes 0..26: one surrounding cell each
  int start = 0, len = 0;
  if (lane < 27) {
    const int nx = cx + (lane % 3) - 1, ny = cy + ((lane / 3) % 3) - 1, nz = cz + (lane / 9) - 1;
    if (nx >= 0 && ny >= 0 && nz >= 0 && nx < h.dims[0] && ny < h.dims[1] && nz < h.dims[2]) {
      const int2 rng = table_find(tab, cap, make_key(cloud, nx, ny, nz));
      start = rng.x;
      len = rng.y - rng.x;
    }
  }
  const int incl = warp_scan_inclusive(len);
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  __syncwarp();
  s_cell_start[lane] = start;
  s_cell_prefix[lane] = incl - len;  // exclusive prefix
  __syncwarp();

  int count = 0;
  for (int t0 = 0; t0 < total; t0 += 32) {
    const int t = t0 + lane;
    bool hit = false;
    unsigned long long packed = 0ull;
    if (t < total) {
      // last cell whose exclusive prefix is <= t (prefixes are non-decreasing; empty cells repeat)
      int lo = 0, hi = 27;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_cell_prefix[mid] <= t) lo = mid; else hi = mid;
      }
      const float4 sp = sorted[s_cell_start[lo] + (t - s_cell_prefix[lo])];
      const float d2 = dist2_ref(qx, qy, qz, sp.x, sp.y, sp.z);
      hit = d2 < r2;
      packed = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)__float_as_int(sp.w);
    }
    const unsigned int mask = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int pos = count + __popc(mask & ((1u << lane) - 1u));
      if (pos < hit_cap) hits[pos] = packed;
    }
    count += __popc(mask);
  }
  __syncwarp();
  if (count <= hit_cap) {
    // rank by counting: keys are distinct (distinct indices), so rank = number of smaller keys.  (A warp-wide
    // bitonic sort of the buffer was measured slower: its ~21-28 dependent shared-memory stages are latency-bound,
    // while these comparisons are independent and pipeline.)
    for (int e = lane; e < count; e += 32) {
      const unsigned long long mine = hits[e];
      int rank = 0;
#pragma unroll 4
      for (int j = 0; j < count; ++j) rank += (hits[j] < mine) ? 1 : 0;
      if (rank < width) row[rank] = (OutT)((int64_t)(unsigned int)(mine & 0xffffffffull) + cloud_base);
    }
  } else {
    // more hits than the shared buffer holds: recount from the candidate ranges (rare, exact, slow)
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + lane;
      bool hit = false;
      unsigned long long mine = 0ull;
      if (t < total) {
        int lo = 0, hi = 27;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (s_cell_prefix[mid] <= t) lo = mid; else hi = mid;
        }
        const float4 sp = sorted[s_cell_start[lo] + (t - s_cell_prefix[lo])];
        const float d2 = dist2_ref(qx, qy, qz, sp.x, sp.y, sp.z);
        hit = d2 < r2;
        mine = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)__float_as_int(sp.w);
      }
      if (hit) {
        int rank = 0;
        for (int cidx = 0; cidx < 27; ++cidx) {
          const int cs = s_cell_start[cidx];
          const int ce = cs + ((cidx < 26 ? s_cell_prefix[cidx + 1] : total) - s_cell_prefix[cidx]);
          for (int j = cs; j < ce; ++j) {
            const float4 o = sorted[j];
            const float od2 = dist2_ref(qx, qy, qz, o.x, o.y, o.z);
            const unsigned long long ok = ((unsigned long long)__float_as_uint(od2) << 32) | (unsigned int)__float_as_int(o.w);
            rank += (od2 < r2 && ok < mine) ? 1 : 0;
          }
        }
        if (rank < width) row[rank] = (OutT)((int64_t)(unsigned int)(mine & 0xffffffffull) + cloud_base);
      }
    }
  }
  for (int e = count + lane; e < width; e += 32) row[e] = (OutT)n_supports;
  __syncwarp();
  return count;
}

// General path: one warp per query (any row width, also the count-only pass with width 0).
template <typename OutT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_grid_query(
    const GridHeader* __restrict__ hdr, const int64_t* __restrict__ s_off, int n_clouds, int64_t n_supports,
    const float4* __restrict__ sorted, const Slot* __restrict__ tab, uint32_t cap,
    const float* __restrict__ queries, const int64_t* __restrict__ q_off, int64_t n_queries, float radius, float r2, int width,
    OutT* __restrict__ out, int32_t* __restrict__ out_counts, int32_t* __restrict__ out_stats, const int32_t* __restrict__ order) {
  __shared__ unsigned long long s_hits[kWarpsPerBlock][kHitCap];
  __shared__ int s_cell_start[kWarpsPerBlock][32];
  __shared__ int s_cell_prefix[kWarpsPerBlock][32];
  __shared__ int s_block_max;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_block_max = 0;
  __syncthreads();
  const GridHeader h = *hdr;
  int my_max = 0;

  // each CTA owns a contiguous slice of the (cell-sorted) processing order: its queries share cells -> L1 hits
  const int64_t per_cta = (n_queries + gridDim.x - 1) / gridDim.x;
  const int64_t it_end = min(n_queries, (int64_t)(blockIdx.x + 1) * per_cta);
  for (int64_t it = (int64_t)blockIdx.x * per_cta + warp; it < it_end; it += kWarpsPerBlock) {
    const int64_t qi = order ? (int64_t)order[it] : it;  // processing order only, never the result
    const float qx = queries[3 * qi], qy = queries[3 * qi + 1], qz = queries[3 * qi + 2];
    const int cloud = cloud_of(q_off, n_clouds, qi);
    int cx, cy, cz;
    cell_coords(h, qx, qy, qz, cx, cy, cz);
    cx = min(max(cx, -2), kCellMax + 2);
    cy = min(max(cy, -2), kCellMax + 2);
    cz = min(max(cz, -2), kCellMax + 2);
    const int count = exact_query_warp<OutT>(h, sorted, tab, cap, qx, qy, qz, cloud, cx, cy, cz, r2, width,
                                             out + qi * (int64_t)width, s_off[cloud], n_supports, s_hits[warp], kHitCap,
                                             s_cell_start[warp], s_cell_prefix[warp]);
    if (out_counts != nullptr && lane == 0) out_counts[qi] = count;
    my_max = max(my_max, count);
  }
  if (lane == 0 && my_max > 0) atomicMax(&s_block_max, my_max);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_block_max > 0) atomicMax(out_stats, s_block_max);
    if (h.status != 0) atomicMax(out_stats + 1, h.status);
    // the 27-cell sweep is only exhaustive while the query radius does not exceed the cell edge
    if ((double)radius * (1.0 + 5e-7) * h.inv_cell > 1.0) atomicMax(out_stats + 1, (int32_t)KPREG_E_RANGE);
  }
}

// ---- hot path: one thread per query ------------------------------------------------------------------------------
// warps (= 32-query blocks) per CTA.  Measured at 64 pairs (radius queries of a step): 8 warps 8.7 ms, 4 warps 6.96, 2 warps 6.84,
// 1 warp 6.80 — a CTA's slot is only refilled when its slowest warp is done, and that costs more than the L1 lines
// consecutive blocks of one CTA share.
constexpr int kTqWarps = 1;
constexpr int kTqCap = 48;        // a lane buffers up to kTqCap - 1 hits; more -> exact_query_warp
constexpr int kTqPitch = kTqCap + 1;  // odd pitch (in 4-byte entries): lane-own and row-cooperative accesses are both conflict-free
constexpr int kTqMaxWidth = 47;   // row widths served by this kernel
constexpr int kTqSpan = 6;        // a group's cells lie within +-kTqSpan (x) / +-1 (y, z) of its first lane's cell
// per warp: 32 hit lists of kTqPitch candidate positions (4 bytes: the kernel is latency-bound, and at 8 bytes per hit
// shared memory allowed 16 warps per SM instead of 32 — measured 292 vs 247 us on the finest level) + 64 span ends
constexpr int kTqWarpBytes = 32 * kTqPitch * 4 + 64 * 4;
constexpr size_t kTqSmemBytes = (size_t)kTqWarps * kTqWarpBytes;
static_assert(kTqWarpBytes % 16 == 0, "per-warp shared-memory slices stay 16-byte aligned");

template <typename OutT>
__global__ void __launch_bounds__(kTqWarps * 32, 32 / kTqWarps) k_grid_query_tq(
    const GridHeader* __restrict__ hdr, const int64_t* __restrict__ s_off, int n_clouds, int64_t n_supports,
    const float4* __restrict__ sorted, const Slot* __restrict__ tab, uint32_t cap,
    const float* __restrict__ queries, const int64_t* __restrict__ q_off, int64_t n_queries, float radius, float r2, int width,
    OutT* __restrict__ out, int32_t* __restrict__ out_counts, int32_t* __restrict__ out_stats, const int32_t* __restrict__ order) {
  extern __shared__ __align__(16) unsigned char tq_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* const wbase = tq_smem + (size_t)warp * kTqWarpBytes;
  int* const warp_hits = reinterpret_cast<int*>(wbase);  // [32][kTqPitch] positions in sorted[]
  int* const my_hits = warp_hits + lane * kTqPitch;
  int* const s_row_start = reinterpret_cast<int*>(wbase + 32 * kTqPitch * 4);
  int* const s_row_end = s_row_start + 32;
  const GridHeader h = *hdr;

  // one 32-query block of the (cell-sorted) processing order per warp
  const int64_t blk = (int64_t)blockIdx.x * kTqWarps + warp;
  const int64_t it = (blk << 5) + lane;
  const bool valid = it < n_queries;
  if (!__any_sync(0xffffffffu, valid)) return;
  int64_t qi = 0;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  int cloud = -1, cx = 0, cy = 0, cz = 0;
  int64_t cloud_base = 0;
  if (valid) {
    qi = order ? (int64_t)order[it] : it;  // processing order only, never the result
    qx = queries[3 * qi]; qy = queries[3 * qi + 1]; qz = queries[3 * qi + 2];
    cloud = cloud_of(q_off, n_clouds, qi);
    cloud_base = s_off[cloud];
    cell_coords(h, qx, qy, qz, cx, cy, cz);
    cx = min(max(cx, -2), kCellMax + 2);
    cy = min(max(cy, -2), kCellMax + 2);
    cz = min(max(cz, -2), kCellMax + 2);
  }
  // hits are appended through a shared-memory cursor that saturates at slot kTqCap (a scratch slot): a lane whose cursor
  // reaches it holds >= kTqCap hits and is finished by exact_query_warp
  const uint32_t hits_lo = (uint32_t)__cvta_generic_to_shared(my_hits), hits_hi = hits_lo + kTqCap * 4;
  uint32_t cursor = hits_lo;
  unsigned int pending = __ballot_sync(0xffffffffu, valid);
  while (pending) {
    // group = the pending lanes whose cell lies in a small window around the first pending lane's cell
    const int head = __ffs(pending) - 1;
    const int hcloud = __shfl_sync(0xffffffffu, cloud, head);
    const int hcx = __shfl_sync(0xffffffffu, cx, head), hcy = __shfl_sync(0xffffffffu, cy, head),
              hcz = __shfl_sync(0xffffffffu, cz, head);
    const bool member = ((pending >> lane) & 1u) && cloud == hcloud && abs(cx - hcx) <= kTqSpan && abs(cy - hcy) <= 1 &&
                        abs(cz - hcz) <= 1;
    pending &= ~__ballot_sync(0xffffffffu, member);
    // bounding box of the group's cells, grown by one cell, clipped to the grid
    const int bx0 = max(__reduce_min_sync(0xffffffffu, member ? cx : 0x7fffffff) - 1, 0);
    const int by0 = max(__reduce_min_sync(0xffffffffu, member ? cy : 0x7fffffff) - 1, 0);
    const int bz0 = max(__reduce_min_sync(0xffffffffu, member ? cz : 0x7fffffff) - 1, 0);
    const int bx1 = min(__reduce_max_sync(0xffffffffu, member ? cx : -0x7fffffff) + 1, h.dims[0] - 1);
    const int by1 = min(__reduce_max_sync(0xffffffffu, member ? cy : -0x7fffffff) + 1, h.dims[1] - 1);
    const int bz1 = min(__reduce_max_sync(0xffffffffu, member ? cz : -0x7fffffff) + 1, h.dims[2] - 1);
    if (bx0 > bx1 || by0 > by1 || bz0 > bz1) continue;  // the whole group lies outside the grid: no neighbours
    const int nbx = bx1 - bx0 + 1, nby = by1 - by0 + 1, n_rows = nby * (bz1 - bz0 + 1);
    // every (cz, cy) row of the box: the span of the sorted array its occupied cells cover
    __syncwarp();
    s_row_start[lane] = 0x7fffffff;
    s_row_end[lane] = 0;
    __syncwarp();
    const int n_cells = n_rows * nbx;
    for (int c = lane; c < n_cells; c += 32) {
      const int r = c / nbx, xi = c - r * nbx;
      const int rz = r / nby, ry = r - rz * nby;
      const int2 rng = table_find(tab, cap, make_key(hcloud, bx0 + xi, by0 + ry, bz0 + rz));
      if (rng.y > rng.x) {
        atomicMin(&s_row_start[r], rng.x);
        atomicMax(&s_row_end[r], rng.y);
      }
    }
    __syncwarp();
    // every lane tests every candidate of the box against its own query (lanes outside the group: radius -1)
    const float my_r2 = member ? r2 : -1.0f;
    auto test = [&](const float4& sp, int pos) {
      if (dist2_ref(qx, qy, qz, sp.x, sp.y, sp.z) < my_r2) {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(cursor), "r"(pos) : "memory");
        cursor = min(cursor + 4u, hits_hi);
      }
    };
    for (int r = 0; r < n_rows; ++r) {
      const int rs = s_row_start[r], re = s_row_end[r];  // warp-uniform
      if (rs >= re) continue;
      int j = rs;
      for (; j + 8 <= re; j += 8) {  // eight independent broadcast loads in flight, then eight tests
        const float4 c0 = __ldg(sorted + j), c1 = __ldg(sorted + j + 1), c2 = __ldg(sorted + j + 2), c3 = __ldg(sorted + j + 3);
        const float4 c4 = __ldg(sorted + j + 4), c5 = __ldg(sorted + j + 5), c6 = __ldg(sorted + j + 6), c7 = __ldg(sorted + j + 7);
        test(c0, j); test(c1, j + 1); test(c2, j + 2); test(c3, j + 3);
        test(c4, j + 4); test(c5, j + 5); test(c6, j + 6); test(c7, j + 7);
      }
      // (software-pipelining the batches — next batch in flight while this one is tested — cost 20 % more instructions
      // for the register copies; at 64 pairs per step the kernel issues ~3 of 4 instructions per cycle, so it lost)
      if (j + 4 <= re) {
        const float4 c0 = __ldg(sorted + j), c1 = __ldg(sorted + j + 1), c2 = __ldg(sorted + j + 2), c3 = __ldg(sorted + j + 3);
        test(c0, j); test(c1, j + 1); test(c2, j + 2); test(c3, j + 3);
        j += 4;
      }
      if (j < re) {  // one to three left: load them together
        const int j1 = min(j + 1, re - 1), j2 = min(j + 2, re - 1);
        const float4 c0 = __ldg(sorted + j), c1 = __ldg(sorted + j1), c2 = __ldg(sorted + j2);
        test(c0, j);
        if (j + 1 < re) test(c1, j1);
        if (j + 2 < re) test(c2, j2);
      }
    }
  }
  __syncwarp();
  // rows leave one at a time, the whole warp on one row: lanes fetch the row's buffered candidates back (L1 / L2),
  // rebuild {d2, index}, rank their key by counting over the row's d2 values (broadcast shared-memory reads; a row
  // holding two equal d2 is re-ranked on (d2, index)) and write the index at its rank — one contiguous span per row
  const bool overflow = cursor == hits_hi;
  int count = (int)((cursor - hits_lo) >> 2);
  const unsigned int n_supports32 = (unsigned int)n_supports;
  float* const s_d2 = reinterpret_cast<float*>(s_row_start);  // 64 floats: the row's d2 values (the span table is dead)
  const unsigned int* const sd = reinterpret_cast<const unsigned int*>(s_d2);
  const int qi32 = (int)qi, base32 = (int)cloud_base;  // both < 2^31 (kpreg_grid_build / kpreg_grid_query check)
  const unsigned int writable = __ballot_sync(0xffffffffu, valid && !overflow);
  // (a) rows of at most 32 hits — nearly all of them: one key per lane.  The next row's candidates are fetched while
  // this row is ranked (the fetch is an L1 / L2 round trip, and a warp walks its 32 rows one after the other).
  {
    unsigned int rem = writable & __ballot_sync(0xffffffffu, count <= 32);
    int r = 0, rn = 0;
    float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](unsigned int mask, int& r_out, int& rn_out, float4& sp_out) {
      r_out = __ffs(mask) - 1;
      rn_out = __shfl_sync(0xffffffffu, count, r_out);
      if (lane < rn_out) sp_out = __ldg(sorted + warp_hits[r_out * kTqPitch + lane]);
    };
    if (rem) fetch(rem, r, rn, sp);
    while (rem) {
      rem &= rem - 1;
      int r_nx = 0, rn_nx = 0;
      float4 sp_nx = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rem) fetch(rem, r_nx, rn_nx, sp_nx);
      const float rx = __shfl_sync(0xffffffffu, qx, r), ry = __shfl_sync(0xffffffffu, qy, r), rz = __shfl_sync(0xffffffffu, qz, r);
      const unsigned int rbase = (unsigned int)__shfl_sync(0xffffffffu, base32, r);
      OutT* __restrict__ row = out + (int64_t)__shfl_sync(0xffffffffu, qi32, r) * width;
      // +inf for the idle lanes: never smaller than a real key
      const float d0 = lane < rn ? dist2_ref(rx, ry, rz, sp.x, sp.y, sp.z) : __int_as_float(0x7f800000);
      const unsigned int i0 = __float_as_uint(sp.w);
      __syncwarp();
      s_d2[lane] = d0;
      __syncwarp();
      // d2 >= 0, so the float order is the order of the bit patterns (the key order of the reference sort)
      const unsigned int u0 = __float_as_uint(d0);
      int rank0 = 0;
#pragma unroll 4
      for (int j = 0; j < rn; ++j)
        asm("{.reg .pred p; setp.lt.u32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(rank0) : "r"(sd[j]), "r"(u0));
      // strict-less ranks of distinct keys add up to rn (rn - 1) / 2; equal d2 (rare) make the sum smaller -> re-rank on (d2, index)
      if (__reduce_add_sync(0xffffffffu, lane < rn ? rank0 : 0) * 2 != rn * (rn - 1)) {
        rank0 = 0;
        for (int j = 0; j < rn; ++j) {
          const unsigned int uj = sd[j], ij = __shfl_sync(0xffffffffu, i0, j);
          rank0 += (uj < u0 || (uj == u0 && ij < i0)) ? 1 : 0;
        }
      }
      if (lane < rn) {
        if (rank0 < width) row[rank0] = (OutT)(i0 + rbase);
      } else {
        if (lane < width) row[lane] = (OutT)n_supports32;
      }
      if (lane + 32 < width) row[lane + 32] = (OutT)n_supports32;  // (kTqMaxWidth < 64)
      r = r_nx; rn = rn_nx; sp = sp_nx;
    }
  }
  // (b) rows of 33 .. kTqCap - 1 hits: two keys per lane
  for (unsigned int rem = writable & __ballot_sync(0xffffffffu, count > 32); rem; rem &= rem - 1) {
    const int r = __ffs(rem) - 1;
    const int rn = __shfl_sync(0xffffffffu, count, r);
    const float rx = __shfl_sync(0xffffffffu, qx, r), ry = __shfl_sync(0xffffffffu, qy, r), rz = __shfl_sync(0xffffffffu, qz, r);
    const unsigned int rbase = (unsigned int)__shfl_sync(0xffffffffu, base32, r);
    OutT* __restrict__ row = out + (int64_t)__shfl_sync(0xffffffffu, qi32, r) * width;
    const int* __restrict__ rh = warp_hits + r * kTqPitch;
    float d1 = __int_as_float(0x7f800000);
    unsigned int i1 = 0xffffffffu;
    const float4 sp0 = __ldg(sorted + rh[lane]);
    const float d0 = dist2_ref(rx, ry, rz, sp0.x, sp0.y, sp0.z);
    const unsigned int i0 = __float_as_uint(sp0.w);
    if (lane + 32 < rn) {
      const float4 sp = __ldg(sorted + rh[lane + 32]);
      d1 = dist2_ref(rx, ry, rz, sp.x, sp.y, sp.z);
      i1 = __float_as_uint(sp.w);
    }
    __syncwarp();
    s_d2[lane] = d0;
    s_d2[lane + 32] = d1;
    __syncwarp();
    const unsigned int u0 = __float_as_uint(d0), u1 = __float_as_uint(d1);
    int rank0 = 0, rank1 = 0;
#pragma unroll 4
    for (int j = 0; j < rn; ++j) {
      const unsigned int uj = sd[j];
      asm("{.reg .pred p; setp.lt.u32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(rank0) : "r"(uj), "r"(u0));
      asm("{.reg .pred p; setp.lt.u32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(rank1) : "r"(uj), "r"(u1));
    }
    if (__reduce_add_sync(0xffffffffu, rank0 + (lane + 32 < rn ? rank1 : 0)) * 2 != rn * (rn - 1)) {
      rank0 = rank1 = 0;
      for (int j = 0; j < rn; ++j) {
        const unsigned int uj = sd[j];
        const unsigned int ij = j < 32 ? __shfl_sync(0xffffffffu, i0, j) : __shfl_sync(0xffffffffu, i1, j - 32);
        rank0 += (uj < u0 || (uj == u0 && ij < i0)) ? 1 : 0;
        rank1 += (uj < u1 || (uj == u1 && ij < i1)) ? 1 : 0;
      }
    }
    if (rank0 < width) row[rank0] = (OutT)(i0 + rbase);
    if (lane + 32 < rn && rank1 < width) row[rank1] = (OutT)(i1 + rbase);
    for (int e = rn + lane; e < width; e += 32) row[e] = (OutT)n_supports32;
  }
  __syncwarp();
  // lanes that buffered too many hits: the exact warp-cooperative routine, one query at a time (the lists are free now)
  unsigned int ov = __ballot_sync(0xffffffffu, valid && overflow);
  while (ov) {
    const int l = __ffs(ov) - 1;
    ov &= ov - 1;
    const int64_t oq = __shfl_sync(0xffffffffu, qi, l), obase = __shfl_sync(0xffffffffu, cloud_base, l);
    const int exact = exact_query_warp<OutT>(h, sorted, tab, cap, __shfl_sync(0xffffffffu, qx, l),
                                             __shfl_sync(0xffffffffu, qy, l), __shfl_sync(0xffffffffu, qz, l),
                                             __shfl_sync(0xffffffffu, cloud, l), __shfl_sync(0xffffffffu, cx, l),
                                             __shfl_sync(0xffffffffu, cy, l), __shfl_sync(0xffffffffu, cz, l), r2, width,
                                             out + oq * (int64_t)width, obase, n_supports,
                                             reinterpret_cast<unsigned long long*>(warp_hits), (32 * kTqPitch * 4) / 8,
                                             s_row_start, s_row_end);
    if (lane == l) count = exact;
  }
  if (out_counts != nullptr && valid) out_counts[qi] = count;
  const int warp_max = __reduce_max_sync(0xffffffffu, count);
  if (lane == 0) {
    if (warp_max > 0) atomicMax(out_stats, warp_max);
    if (blk == 0) {
      if (h.status != 0) atomicMax(out_stats + 1, h.status);
      // the one-cell margin of a group's box is only exhaustive while the query radius does not exceed the cell edge
      if ((double)radius * (1.0 + 5e-7) * h.inv_cell > 1.0) atomicMax(out_stats + 1, (int32_t)KPREG_E_RANGE);
    }
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256) k_pack_rows(const int32_t* __restrict__ in, int64_t n_rows, int in_width, int out_width,
                                                   OutT* __restrict__ out) {
  const int64_t total = n_rows * (int64_t)out_width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / out_width;
    const int cidx = (int)(i - r * out_width);
    out[i] = (OutT)in[r * (int64_t)in_width + cidx];
  }
}

}  // namespace
}  // namespace kpreg

using namespace kpreg;

extern "C" int kpreg_grid_workspace_bytes(int64_t n_supports, int n_clouds, size_t* bytes) {
  if (!bytes || n_supports < 0 || n_clouds < 0) return KPREG_E_INVALID;
  *bytes = carve_grid(nullptr, n_supports, n_clouds).total;
  return KPREG_OK;
}

extern "C" int kpreg_grid_build(const float* supports, const int32_t* s_lens, int64_t n, int n_clouds, float cell,
                                void* grid, size_t grid_bytes, int32_t* out_order, void* stream_) {
  if (!s_lens || !grid || n < 0 || n_clouds < 1 || !(cell > 0.f)) return KPREG_E_INVALID;
  if (n > 0 && !supports) return KPREG_E_INVALID;
  if (n >= (int64_t)0x7fffffff || n_clouds >= (1 << (63 - kCloudShift))) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  GridWs w = carve_grid(grid, n, n_clouds);
  if (w.total > grid_bytes) return KPREG_E_WORKSPACE;
  ProfScope prof(KPREG_FAM_GRID_BUILD, stream);
  int rc = launch_cloud_offsets(s_lens, n_clouds, w.off, stream);
  if (rc) return rc;
  k_grid_init<<<1, 32, 0, stream>>>(w.bbox, w.hdr);
  KP_LAUNCH_CHECK();
  const int blocks = n > 0 ? ceil_div(n, 256) : 1;
  if (n > 0) {
    k_grid_bbox<<<blocks < 4 * kNumSMs ? blocks : 4 * kNumSMs, 256, 0, stream>>>(supports, n, w.bbox);
    KP_LAUNCH_CHECK();
  }
  k_grid_header<<<1, 32, 0, stream>>>(w.bbox, cell, w.hdr);
  KP_LAUNCH_CHECK();
  k_table_clear<<<ceil_div(w.tab_cap, 1024) < 8 * kNumSMs ? ceil_div(w.tab_cap, 1024) : 8 * kNumSMs, 256, 0, stream>>>(w.tab, w.tab_cap);
  KP_LAUNCH_CHECK();
  if (n == 0) return KPREG_OK;
  k_cell_keys<<<blocks, 256, 0, stream>>>(supports, w.off, n_clouds, n, w.hdr, w.keys0, w.idx0);
  KP_LAUNCH_CHECK();
  cub::DoubleBuffer<uint64_t> dk(w.keys0, w.keys1);
  cub::DoubleBuffer<uint32_t> dv(w.idx0, w.idx1);
  const int end_bit = kCloudShift + bits_for((uint64_t)n_clouds);
  size_t need = 0;
  KP_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)n, 0, end_bit, stream));
  if (need > w.cub_tmp_bytes) return KPREG_E_WORKSPACE;
  KP_CUDA_TRY(cub::DeviceRadixSort::SortPairs(w.cub_tmp, need, dk, dv, (int)n, 0, end_bit, stream));
  count_launches((unsigned long long)(2 + (end_bit + 7) / 8));
  k_grid_fill<<<blocks, 256, 0, stream>>>(supports, w.off, dk.Current(), dv.Current(), n, w.sorted, w.tab, w.tab_cap);
  KP_LAUNCH_CHECK();
  if (out_order)  // the cell-sorted permutation of the supports: a spatially coherent processing order for later kernels
    KP_CUDA_TRY(cudaMemcpyAsync(out_order, dv.Current(), sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
  return KPREG_OK;
}

extern "C" int kpreg_grid_query(const void* grid, int64_t n, int n_clouds, const float* queries, const int32_t* q_lens,
                                int64_t n_queries, float radius, int width, int idx64, const int32_t* order, void* out_idx,
                                int32_t* out_counts, int32_t* out_stats, void* stream_) {
  if (!grid || !q_lens || !out_stats || n < 0 || n_clouds < 1 || n_queries < 0 || width < 0 || !(radius > 0.f)) return KPREG_E_INVALID;
  if (n_queries == 0) return KPREG_OK;
  if (!queries || (width > 0 && !out_idx) || n_queries >= (int64_t)0x7fffffff) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  GridWs w = carve_grid(const_cast<void*>(grid), n, n_clouds);
  int64_t* q_off = w.q_off;
  int rc = launch_cloud_offsets(q_lens, n_clouds, q_off, stream);
  if (rc) return rc;
  const float r2 = radius * radius;  // neighbors.cpp:226, fp32
  ProfScope prof(KPREG_FAM_GRID_QUERY, stream);
  static const bool force_general = [] { const char* e = getenv("KPREG_QUERY_GENERAL"); return e && e[0] == '1'; }();  // A/B measurements
  // The thread-per-query kernel needs ~150 k queries to fill the GPU (148 SMs x 32 warps x 32 queries) and a warp walks
  // its 32 queries through long dependent phases (measured 190-250 us for ANY launch below one wave); smaller query
  // sets are served faster by the warp-per-query kernel (7 us per query and warp, ~1.3 ns per query in bulk).
  static const int64_t tq_min = [] { const char* e = getenv("KPREG_QUERY_TQ_MIN"); return e ? (int64_t)atoll(e) : (int64_t)131072; }();
  if (width >= 1 && width <= kTqMaxWidth && n_queries >= tq_min && !force_general) {
    // hot path: one thread per query, 32-query blocks of the processing order per warp
    const int blocks = ceil_div(ceil_div(n_queries, 32), kTqWarps);  // n_queries < 2^31 rows (checked above): fits a grid
    if (idx64) {
      static PerDeviceOnce once;
      const int rc_cfg = once.run([]() -> int {
        KP_CUDA_TRY(cudaFuncSetAttribute(k_grid_query_tq<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTqSmemBytes));
        return KPREG_OK;
      });
      if (rc_cfg) return rc_cfg;
      k_grid_query_tq<int64_t><<<blocks, kTqWarps * 32, kTqSmemBytes, stream>>>(w.hdr, w.off, n_clouds, n, w.sorted, w.tab,
                                                                                 w.tab_cap, queries, q_off, n_queries, radius, r2, width,
                                                                                 static_cast<int64_t*>(out_idx), out_counts, out_stats, order);
    } else {
      static PerDeviceOnce once;
      const int rc_cfg = once.run([]() -> int {
        KP_CUDA_TRY(cudaFuncSetAttribute(k_grid_query_tq<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTqSmemBytes));
        return KPREG_OK;
      });
      if (rc_cfg) return rc_cfg;
      k_grid_query_tq<int32_t><<<blocks, kTqWarps * 32, kTqSmemBytes, stream>>>(w.hdr, w.off, n_clouds, n, w.sorted, w.tab,
                                                                                 w.tab_cap, queries, q_off, n_queries, radius, r2, width,
                                                                                 static_cast<int32_t*>(out_idx), out_counts, out_stats, order);
    }
    KP_LAUNCH_CHECK();
    return KPREG_OK;
  }
  int blocks = ceil_div(n_queries, kWarpsPerBlock);
  const int max_blocks = kNumSMs * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  if (idx64) {
    k_grid_query<int64_t><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(w.hdr, w.off, n_clouds, n, w.sorted, w.tab,
                                                                     w.tab_cap, queries, q_off, n_queries, radius, r2, width,
                                                                     static_cast<int64_t*>(out_idx), out_counts, out_stats, order);
  } else {
    k_grid_query<int32_t><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(w.hdr, w.off, n_clouds, n, w.sorted, w.tab,
                                                                     w.tab_cap, queries, q_off, n_queries, radius, r2, width,
                                                                     static_cast<int32_t*>(out_idx), out_counts, out_stats, order);
  }
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

extern "C" int kpreg_pack_rows(const int32_t* in, int64_t n_rows, int in_width, int out_width, int idx64, void* out,
                               void* stream_) {
  if (n_rows < 0 || in_width < 0 || out_width < 0 || out_width > in_width) return KPREG_E_INVALID;
  if (n_rows == 0 || out_width == 0) return KPREG_OK;
  if (!in || !out) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t total = n_rows * (int64_t)out_width;
  int blocks = ceil_div(total, 256);
  if (blocks > 32 * kNumSMs) blocks = 32 * kNumSMs;
  ProfScope prof(KPREG_FAM_OTHER, stream);
  if (idx64) k_pack_rows<int64_t><<<blocks, 256, 0, stream>>>(in, n_rows, in_width, out_width, static_cast<int64_t*>(out));
  else k_pack_rows<int32_t><<<blocks, 256, 0, stream>>>(in, n_rows, in_width, out_width, static_cast<int32_t*>(out));
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}
