// Voxel-grid barycentre subsampling of a stacked batch on sm_100a.
//
// Replaces batch_grid_subsampling() (reference cpp_subsampling/grid_subsampling/grid_subsampling.cpp:109-211,
// per-cloud body :5-106).  Results are bit-identical to the reference, including the order of the
// subsampled points, which in the reference is the iteration order of a libstdc++
// std::unordered_map<size_t, SampledData> (:48, :85-87).
//
// Pipeline (all clouds of the batch in the same launches, nothing synchronises the host):
//   1. per-cloud bounding box             k_bbox          (reference :25-27)
//   2. grid origin / dimensions           k_origin        (reference :27-31, fp32 arithmetic)
//   3. voxel key per point                k_keys          (reference :53-56), key = cloud<<40 | voxel
//   4. stable radix sort by key           cub::DeviceRadixSort (points of one voxel stay in input order)
//   5. first-appearance rank of each voxel: flag the first point of every voxel, prefix-sum the
//      flags in input order            k_heads + cub::DeviceScan
//   6. segmented mean                     k_voxels: one thread per voxel adds its points
//      sequentially in input order (fp32, like SampledData::update_points, grid_subsampling.h:74-79)
//      and scales by (float)(1.0/count) (reference :87)
//   7. hash-table order                   k_order: one CTA (or, for large clouds, one 8-CTA cluster) per cloud replays the table's growth
//      13 -> 29 -> 59 -> ... (libstdc++ prime policy, load factor 1) as ~log2(m) parallel rounds.
//      For a fixed bucket count the list order is a pure function of the insertion sequence S:
//      buckets by descending first appearance in S, inside a bucket by descending position in S;
//      a rehash re-inserts the current list front to back, then the new keys follow.
#include <cub/cub.cuh>

#include "common.cuh"

namespace kpreg {
namespace {

constexpr int kVoxelBits = 40;
constexpr uint64_t kVoxelMask = (1ull << kVoxelBits) - 1ull;

// Bucket counts of std::unordered_map<size_t,T> as it grows from empty at load factor 1 (libstdc++
// _Prime_rehash_policy: 13, then the next listed prime >= 2x).  The test suite re-derives the
// sequence from libstdc++ itself and checks this table against it.
__constant__ unsigned int c_table_sizes[27] = {
    13u,       29u,       59u,       127u,      257u,       541u,       1109u,
    2357u,     5087u,     10273u,    20753u,    42043u,     85229u,     172933u,
    351061u,   712697u,   1447153u,  2938679u,  5967347u,   12117689u,  24607243u,
    49969847u, 101473717u, 206062531u, 418451333u, 849749479u, 1725587117u};

__global__ void k_cloud_offsets(const int32_t* __restrict__ lens, int n_clouds, int64_t* __restrict__ off) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t acc = 0;
    for (int c = 0; c < n_clouds; ++c) {
      off[c] = acc;
      acc += lens[c] > 0 ? lens[c] : 0;
    }
    off[n_clouds] = acc;
  }
}

__global__ void k_init_bbox(unsigned int* __restrict__ bbox, int n_clouds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_clouds * 6) bbox[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}

// bbox[c*6 + 0..2] = min xyz, [c*6 + 3..5] = max xyz (ordered-uint encoding).
__global__ void __launch_bounds__(256) k_bbox(const float* __restrict__ pts, const int64_t* __restrict__ off,
                                              int n_clouds, int64_t n, unsigned int* __restrict__ bbox) {
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = base + threadIdx.x;
    bool live = i < n;
    int c = live ? cloud_of(off, n_clouds, i) : -1;
    float x = 0.f, y = 0.f, z = 0.f;
    if (live) { x = pts[3 * i]; y = pts[3 * i + 1]; z = pts[3 * i + 2]; }
    int c0 = __shfl_sync(0xffffffffu, c, 0);
    bool uniform = __all_sync(0xffffffffu, c == c0);
    if (uniform) {
      float mnx = x, mny = y, mnz = z, mxx = x, mxy = y, mxz = z;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
      }
      if (lane_id() == 0 && c0 >= 0) {
        unsigned int* bb = bbox + c0 * 6;
        atomicMin(bb + 0, float_to_ordered(mnx));
        atomicMin(bb + 1, float_to_ordered(mny));
        atomicMin(bb + 2, float_to_ordered(mnz));
        atomicMax(bb + 3, float_to_ordered(mxx));
        atomicMax(bb + 4, float_to_ordered(mxy));
        atomicMax(bb + 5, float_to_ordered(mxz));
      }
    } else if (live) {
      unsigned int* bb = bbox + c * 6;
      atomicMin(bb + 0, float_to_ordered(x));
      atomicMin(bb + 1, float_to_ordered(y));
      atomicMin(bb + 2, float_to_ordered(z));
      atomicMax(bb + 3, float_to_ordered(x));
      atomicMax(bb + 4, float_to_ordered(y));
      atomicMax(bb + 5, float_to_ordered(z));
    }
  }
}

struct GridDesc {  // per cloud
  float org[3];
  unsigned int pad;
  unsigned long long nx, nxy;
};

// origin = floor(min * (1/dl)) * dl; NX = floor((max.x - origin.x)/dl) + 1 ... all fp32, no contraction.
__global__ void k_origin(const unsigned int* __restrict__ bbox, const int64_t* __restrict__ off, int n_clouds,
                         float dl, GridDesc* __restrict__ grid, int32_t* __restrict__ status) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clouds) return;
  GridDesc g;
  g.pad = 0;
  if (off[c + 1] == off[c]) {
    g.org[0] = g.org[1] = g.org[2] = 0.f;
    g.nx = g.nxy = 1;
    grid[c] = g;
    return;
  }
  float inv = __fdiv_rn(1.0f, dl);
  float ext[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float mn = ordered_to_float(bbox[c * 6 + d]);
    float mx = ordered_to_float(bbox[c * 6 + 3 + d]);
    g.org[d] = __fmul_rn(floorf(__fmul_rn(mn, inv)), dl);
    ext[d] = floorf(__fdiv_rn(__fsub_rn(mx, g.org[d]), dl));
  }
  unsigned long long nx = (unsigned long long)ext[0] + 1ull;
  unsigned long long ny = (unsigned long long)ext[1] + 1ull;
  unsigned long long nz = (unsigned long long)ext[2] + 1ull;
  g.nx = nx;
  g.nxy = nx * ny;
  grid[c] = g;
  // the composite sort key keeps 40 bits for the voxel index
  double cells = (double)nx * (double)ny * (double)nz;
  if (!(cells < (double)(1ull << kVoxelBits))) atomicMax(status, (int32_t)KPREG_E_RANGE);
}

__global__ void __launch_bounds__(256) k_keys(const float* __restrict__ pts, const int64_t* __restrict__ off,
                                              int n_clouds, int64_t n, float dl, const GridDesc* __restrict__ grid,
                                              uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cloud_of(off, n_clouds, i);
  const GridDesc g = grid[c];
  unsigned long long ix = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(pts[3 * i + 0], g.org[0]), dl));
  unsigned long long iy = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(pts[3 * i + 1], g.org[1]), dl));
  unsigned long long iz = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(pts[3 * i + 2], g.org[2]), dl));
  unsigned long long v = ix + g.nx * iy + g.nxy * iz;
  keys[i] = ((uint64_t)c << kVoxelBits) | (v & kVoxelMask);
  idx[i] = (uint32_t)i;
}

// sorted position p starts a voxel iff its key differs from p-1; the voxel's first point (in input
// order, the sort being stable) gets flag 1.
__global__ void __launch_bounds__(256) k_heads(const uint64_t* __restrict__ keys_sorted, const uint32_t* __restrict__ idx_sorted,
                                               int64_t n, uint32_t* __restrict__ first_flag) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  bool head = (p == 0) || (keys_sorted[p] != keys_sorted[p - 1]);
  first_flag[idx_sorted[p]] = head ? 1u : 0u;
}

// voff[c] = number of voxels before cloud c; out_counts / out offsets with the max_p cap.
__global__ void k_voxel_offsets(const int64_t* __restrict__ off, int n_clouds, int64_t n, const uint32_t* __restrict__ first_flag,
                                const uint32_t* __restrict__ rank, int max_p, int64_t* __restrict__ voff,
                                int64_t* __restrict__ ooff, int32_t* __restrict__ out_counts, const int32_t* __restrict__ status) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t total = n > 0 ? (int64_t)rank[n - 1] + (int64_t)first_flag[n - 1] : 0;
  for (int c = 0; c <= n_clouds; ++c) voff[c] = (off[c] < n) ? (int64_t)rank[off[c]] : total;
  int64_t acc = 0;
  for (int c = 0; c < n_clouds; ++c) {
    int64_t m = voff[c + 1] - voff[c];
    if (max_p >= 1 && m > max_p) m = max_p;
    ooff[c] = acc;
    out_counts[c] = (int32_t)m;
    acc += m;
  }
  ooff[n_clouds] = acc;
  out_counts[n_clouds] = (int32_t)acc;
  out_counts[n_clouds + 1] = *status;
}

// One thread per voxel: sequential fp32 sum of its points in input order, then * (float)(1.0/count).
__global__ void __launch_bounds__(256) k_voxels(const float* __restrict__ pts, const uint64_t* __restrict__ keys_sorted,
                                                const uint32_t* __restrict__ idx_sorted, const uint32_t* __restrict__ rank,
                                                int64_t n, float* __restrict__ bary, uint64_t* __restrict__ vkey) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint64_t key = keys_sorted[p];
  if (p > 0 && keys_sorted[p - 1] == key) return;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  int count = 0;
  for (int64_t j = p; j < n && keys_sorted[j] == key; ++j) {
    const float* q = pts + 3 * (int64_t)idx_sorted[j];
    sx = __fadd_rn(sx, q[0]);
    sy = __fadd_rn(sy, q[1]);
    sz = __fadd_rn(sz, q[2]);
    ++count;
  }
  const float a = (float)(1.0 / (double)count);
  const int64_t v = rank[idx_sorted[p]];
  bary[3 * v + 0] = __fmul_rn(sx, a);
  bary[3 * v + 1] = __fmul_rn(sy, a);
  bary[3 * v + 2] = __fmul_rn(sz, a);
  vkey[v] = key & kVoxelMask;
}

template <int THREADS>
__device__ __forceinline__ int block_scan_inclusive(int v, int* smem_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = warp_scan_inclusive(v);
  if (lane == 31) smem_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < THREADS / 32 ? smem_warp[lane] : 0;
    int ws = warp_scan_inclusive(w);
    smem_warp[lane] = ws;  // inclusive sums of warp totals (32 slots)
  }
  __syncthreads();
  int prev = warp > 0 ? smem_warp[warp - 1] : 0;
  total = smem_warp[THREADS / 32 - 1];
  __syncthreads();
  return incl + prev;
}

// Replay of the unordered_map's list order: one TEAM per cloud — a single CTA, or for large clouds a thread-block
// cluster of CS CTAs that synchronise with barrier.cluster between the passes of a growth round (the early rounds,
// whose tables are tiny, are run by the cluster's first CTA alone).  Scratch arrays are per point/cloud:
//   list0/list1/nxt/start/bkt : capacity len_c, at offset off[c]
//   b_first/b_count/b_head    : capacity 3*len_c + 16, at offset 3*off[c] + 16*c
//   part                      : 16 ints per cloud (per-CTA scan totals of a cluster)
constexpr unsigned int kSoloTable = 5087;  // growth rounds up to this table size stay on one CTA

template <int THREADS, int CS>
__global__ void __launch_bounds__(THREADS) k_order(
    const int64_t* __restrict__ off, const int64_t* __restrict__ voff, const uint64_t* __restrict__ vkey,
    const float* __restrict__ bary, const int64_t* __restrict__ ooff, const int32_t* __restrict__ out_counts,
    int32_t* __restrict__ list0, int32_t* __restrict__ list1, int32_t* __restrict__ nxt, int32_t* __restrict__ start,
    uint32_t* __restrict__ bkt, int32_t* __restrict__ b_first, int32_t* __restrict__ b_count,
    int32_t* __restrict__ b_head, int32_t* __restrict__ part_all, float* __restrict__ out_pts) {
  __shared__ int s_warp[32];
  const int c = blockIdx.x / CS;
  const int rank = blockIdx.x % CS;  // == %cluster_ctarank for a 1-D cluster of CS CTAs
  const int64_t vbase = voff[c];
  const int m = (int)(voff[c + 1] - vbase);
  if (m == 0) return;  // uniform over the cluster
  const int64_t pbase = off[c];
  const int64_t bbase = 3 * pbase + 16 * (int64_t)c;
  int32_t* cur = list0 + pbase;
  int32_t* nxl = list1 + pbase;
  int32_t* chain = nxt + pbase;
  int32_t* st = start + pbase;
  uint32_t* bk = bkt + pbase;
  int32_t* bF = b_first + bbase;
  int32_t* bC = b_count + bbase;
  int32_t* bH = b_head + bbase;
  int32_t* part = part_all + 16 * (int64_t)c;
  const uint64_t* key = vkey + vbase;

  auto cluster_sync = [&]() {
    if constexpr (CS > 1) {
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      __syncthreads();
    }
  };

  int n_prev = 0;
  bool team_mode = false;  // false: this round is run by CTA 0 of the cluster alone
  for (int g = 0;; ++g) {
    const unsigned int nb = c_table_sizes[g];
    const int n = (unsigned int)m < nb ? m : (int)nb;
    if (CS > 1 && !team_mode && nb > kSoloTable) {
      cluster_sync();  // hand-over: CTA 0's lists become visible to the whole cluster
      team_mode = true;
    }
    const int team = team_mode ? CS : 1;
    const bool active = team_mode || rank == 0;
    const int tid = (team_mode ? rank : 0) * THREADS + (int)threadIdx.x;
    const int nthreads = team * THREADS;
    auto team_sync = [&]() { if (team_mode) cluster_sync(); else __syncthreads(); };

    if (active) {
      for (unsigned int b = tid; b < nb; b += nthreads) {
        bF[b] = 0x7fffffff;
        bC[b] = 0;
        bH[b] = -1;
      }
      team_sync();
      // insertion sequence S: the previous list front to back, then the new voxels by first appearance
      for (int pos = tid; pos < n; pos += nthreads) {
        const int r = pos < n_prev ? cur[pos] : pos;
        const unsigned int b = (unsigned int)(key[r] % (uint64_t)nb);
        bk[pos] = b;
        atomicMin(&bF[b], pos);
        atomicAdd(&bC[b], 1);
        chain[pos] = atomicExch(&bH[b], pos);
      }
      team_sync();
      // start[pos] for bucket-first positions = number of elements in buckets that appear later in S: an exclusive
      // scan over the REVERSED positions; every CTA of the team scans one contiguous chunk, then adds the totals of
      // the chunks before it
      const int chunk = (n + team - 1) / team;
      const int lo = (team_mode ? rank : 0) * chunk;
      const int hi = min(n, lo + chunk);
      int carry = 0;
      for (int t0 = lo; t0 < hi; t0 += THREADS) {
        const int rpos = t0 + (int)threadIdx.x;
        const int pos = n - 1 - rpos;
        int v = 0;
        if (rpos < hi) {
          const unsigned int b = bk[pos];
          v = (bF[b] == pos) ? bC[b] : 0;
        }
        int total;
        const int incl = block_scan_inclusive<THREADS>(v, s_warp, total);
        if (rpos < hi) st[pos] = carry + incl - v;
        carry += total;
      }
      if (team_mode) {
        if (threadIdx.x == 0) part[rank] = carry;
        cluster_sync();
        int before = 0;
        for (int r = 0; r < rank; ++r) before += part[r];
        if (before != 0)
          for (int rpos = lo + (int)threadIdx.x; rpos < hi; rpos += THREADS) st[n - 1 - rpos] += before;
      }
      team_sync();
      for (int pos = tid; pos < n; pos += nthreads) {
        const int r = pos < n_prev ? cur[pos] : pos;
        const unsigned int b = bk[pos];
        int later = 0;
        for (int j = bH[b]; j != -1; j = chain[j]) later += (j > pos) ? 1 : 0;
        nxl[st[bF[b]] + later] = r;
      }
      team_sync();
    }
    int32_t* t = cur; cur = nxl; nxl = t;
    n_prev = n;
    if (n == m) break;
  }
  if (CS > 1 && !team_mode) cluster_sync();  // small cloud finished by CTA 0 alone: publish its list to the cluster
  const int keep = out_counts[c];
  const int64_t obase = ooff[c];
  for (int j = rank * THREADS + (int)threadIdx.x; j < keep; j += CS * THREADS) {
    const int64_t v = vbase + cur[j];
    out_pts[3 * (obase + j) + 0] = bary[3 * v + 0];
    out_pts[3 * (obase + j) + 1] = bary[3 * v + 1];
    out_pts[3 * (obase + j) + 2] = bary[3 * v + 2];
  }
}

struct SubsampleWs {
  int64_t* off; int64_t* voff; int64_t* ooff; unsigned int* bbox; GridDesc* grid; int32_t* status;
  uint64_t* keys0; uint64_t* keys1; uint32_t* idx0; uint32_t* idx1;
  uint32_t* first_flag; uint32_t* rank; float* bary; uint64_t* vkey;
  int32_t* list0; int32_t* list1; int32_t* nxt; int32_t* start; uint32_t* bkt;
  int32_t* b_first; int32_t* b_count; int32_t* b_head; int32_t* part;
  void* cub_tmp; size_t cub_tmp_bytes; size_t total;
};

SubsampleWs carve_subsample(void* base, int64_t n, int n_clouds) {
  SubsampleWs w;
  Carver cv(base);
  const size_t np = (size_t)(n > 0 ? n : 1);
  const size_t nc = (size_t)n_clouds;
  w.off = cv.take<int64_t>(nc + 1);
  w.voff = cv.take<int64_t>(nc + 1);
  w.ooff = cv.take<int64_t>(nc + 1);
  w.bbox = cv.take<unsigned int>(nc * 6);
  w.grid = cv.take<GridDesc>(nc);
  w.status = cv.take<int32_t>(4);
  w.keys0 = cv.take<uint64_t>(np);
  w.keys1 = cv.take<uint64_t>(np);
  w.idx0 = cv.take<uint32_t>(np);
  w.idx1 = cv.take<uint32_t>(np);
  w.first_flag = cv.take<uint32_t>(np);
  w.rank = cv.take<uint32_t>(np);
  w.bary = cv.take<float>(np * 3);
  w.vkey = cv.take<uint64_t>(np);
  w.list0 = cv.take<int32_t>(np);
  w.list1 = cv.take<int32_t>(np);
  w.nxt = cv.take<int32_t>(np);
  w.start = cv.take<int32_t>(np);
  w.bkt = cv.take<uint32_t>(np);
  const size_t nbk = 3 * np + 16 * nc + 16;
  w.b_first = cv.take<int32_t>(nbk);
  w.b_count = cv.take<int32_t>(nbk);
  w.b_head = cv.take<int32_t>(nbk);
  w.part = cv.take<int32_t>(16 * nc + 16);
  w.cub_tmp_bytes = (size_t)(8u << 20) + np * 16;
  w.cub_tmp = cv.take<char>(w.cub_tmp_bytes);
  w.total = align_up(cv.used, 256);
  return w;
}

}  // namespace

int launch_cloud_offsets(const int32_t* lens, int n_clouds, int64_t* off, cudaStream_t stream) {
  k_cloud_offsets<<<1, 32, 0, stream>>>(lens, n_clouds, off);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

}  // namespace kpreg

using namespace kpreg;

extern "C" int kpreg_subsample_workspace_bytes(int64_t n_points, int n_clouds, size_t* bytes) {
  if (!bytes || n_points < 0 || n_clouds < 0) return KPREG_E_INVALID;
  *bytes = carve_subsample(nullptr, n_points, n_clouds).total;
  return KPREG_OK;
}

extern "C" int kpreg_subsample_batch(const float* pts, const int32_t* lens, int64_t n, int n_clouds, float dl,
                                     int max_p, float* out_pts, int32_t* out_counts, void* workspace,
                                     size_t workspace_bytes, void* stream_) {
  if (!lens || !out_counts || !workspace || n < 0 || n_clouds < 1 || !(dl > 0.f)) return KPREG_E_INVALID;
  if (n > 0 && (!pts || !out_pts)) return KPREG_E_INVALID;
  if (n >= (int64_t)0x7fffffff || n_clouds >= (1 << (63 - kVoxelBits))) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  SubsampleWs w = carve_subsample(workspace, n, n_clouds);
  if (w.total > workspace_bytes) return KPREG_E_WORKSPACE;

  ProfScope prof(KPREG_FAM_SUBSAMPLE, stream);
  KP_CUDA_TRY(cudaMemsetAsync(w.status, 0, 4 * sizeof(int32_t), stream));
  int rc = launch_cloud_offsets(lens, n_clouds, w.off, stream);
  if (rc) return rc;
  if (n == 0) {
    KP_CUDA_TRY(cudaMemsetAsync(out_counts, 0, (n_clouds + 2) * sizeof(int32_t), stream));
    return KPREG_OK;
  }
  k_init_bbox<<<ceil_div(n_clouds * 6, 256), 256, 0, stream>>>(w.bbox, n_clouds);
  KP_LAUNCH_CHECK();
  const int pt_blocks = ceil_div(n, 256);
  k_bbox<<<pt_blocks < 8 * kNumSMs ? pt_blocks : 8 * kNumSMs, 256, 0, stream>>>(pts, w.off, n_clouds, n, w.bbox);
  KP_LAUNCH_CHECK();
  k_origin<<<ceil_div(n_clouds, 128), 128, 0, stream>>>(w.bbox, w.off, n_clouds, dl, w.grid, w.status);
  KP_LAUNCH_CHECK();
  k_keys<<<pt_blocks, 256, 0, stream>>>(pts, w.off, n_clouds, n, dl, w.grid, w.keys0, w.idx0);
  KP_LAUNCH_CHECK();

  cub::DoubleBuffer<uint64_t> dk(w.keys0, w.keys1);
  cub::DoubleBuffer<uint32_t> dv(w.idx0, w.idx1);
  const int end_bit = kVoxelBits + bits_for((uint64_t)n_clouds);
  size_t need = 0;
  KP_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)n, 0, end_bit, stream));
  if (need > w.cub_tmp_bytes) return KPREG_E_WORKSPACE;
  KP_CUDA_TRY(cub::DeviceRadixSort::SortPairs(w.cub_tmp, need, dk, dv, (int)n, 0, end_bit, stream));
  count_launches((unsigned long long)(2 + (end_bit + 7) / 8));
  const uint64_t* ks = dk.Current();
  const uint32_t* is = dv.Current();

  k_heads<<<pt_blocks, 256, 0, stream>>>(ks, is, n, w.first_flag);
  KP_LAUNCH_CHECK();
  KP_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, w.first_flag, w.rank, (int)n, stream));
  if (need > w.cub_tmp_bytes) return KPREG_E_WORKSPACE;
  KP_CUDA_TRY(cub::DeviceScan::ExclusiveSum(w.cub_tmp, need, w.first_flag, w.rank, (int)n, stream));
  count_launches(2);
  k_voxel_offsets<<<1, 32, 0, stream>>>(w.off, n_clouds, n, w.first_flag, w.rank, max_p, w.voff, w.ooff, out_counts,
                                        w.status);
  KP_LAUNCH_CHECK();
  k_voxels<<<pt_blocks, 256, 0, stream>>>(pts, ks, is, w.rank, n, w.bary, w.vkey);
  KP_LAUNCH_CHECK();
  if (n / n_clouds > 32768) {
    // large clouds: a cluster of 8 CTAs per cloud (the rounds of the replay are parallel over the cloud's voxels)
    constexpr int kCluster = 8;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clouds * kCluster));
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    KP_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_order<1024, kCluster>, (const int64_t*)w.off, (const int64_t*)w.voff,
                                   (const uint64_t*)w.vkey, (const float*)w.bary, (const int64_t*)w.ooff,
                                   (const int32_t*)out_counts, w.list0, w.list1, w.nxt, w.start, w.bkt, w.b_first, w.b_count,
                                   w.b_head, w.part, out_pts));
  } else {
    k_order<1024, 1><<<n_clouds, 1024, 0, stream>>>(w.off, w.voff, w.vkey, w.bary, w.ooff, out_counts, w.list0, w.list1,
                                                   w.nxt, w.start, w.bkt, w.b_first, w.b_count, w.b_head, w.part, out_pts);
  }
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}
