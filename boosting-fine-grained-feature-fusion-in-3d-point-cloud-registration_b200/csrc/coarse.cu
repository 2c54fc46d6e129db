// The steps on either side of the KPConv path on sm_100a — all HBM-bound index / gather work:
//
//   k_overlap_pool   compute_overlaps' masked mean over the pooling rows
//                    (reference models/backbone_kpconv/finegrained_kpconv.py:545-571)
//   k_sine_embed     PositionEmbeddingCoordsSine on stacked points
//                    (reference models/transformer/position_embedding.py:8-49)
//   k_pack_coarse    split_src_tgt + pad_sequence of the projected coarse features and of their position embedding,
//                    with the padding masks, in one pass over the coarse level
//                    (reference utils/seq_manipulation.py:6-48 as used at models/finegrained_regtr.py:149-172)
//   k_shuffle_gather ShufflePoints' permutation gather of points + overlap mask and its reverse index
//                    (reference data_loaders/transforms.py:95-131)
//   k_remap_pairs    ShufflePoints' remapping of the correspondence index pairs through the reverse indices
#include "common.cuh"

namespace kpreg {
namespace {

// out[n] = clamp( sum_{h valid} level[idx[n,h]] / #valid, 0, 1 ),  valid = idx < n_s.  One warp per pooling row.
// A row without a valid entry gives 0/0 = NaN like the reference's torch expression.
template <typename IdxT>
__global__ void __launch_bounds__(256) k_overlap_pool(const float* __restrict__ level, const IdxT* __restrict__ idx, int64_t n_q,
                                                      int64_t n_s, int n_nbrs, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= n_q) return;
  float sum = 0.f;
  int cnt = 0;
  for (int h = lane; h < n_nbrs; h += 32) {
    const int64_t j = (int64_t)idx[n * n_nbrs + h];
    if (j >= 0 && j < n_s) { sum += level[j]; ++cnt; }
  }
  sum = warp_sum(sum);
  cnt = warp_sum(cnt);
  if (lane == 0) {
    const float v = sum / (float)cnt;
    out[n] = v != v ? v : fminf(fmaxf(v, 0.f), 1.f);  // torch.clamp keeps NaN
  }
}

// pos_emb[(d * F + k)] = k even ? sin(x_d * scale / dim_t[k]) : cos(x_d * scale / dim_t[k]);  columns >= n_dim * F are 0.
// dim_t[k] = temperature^(2 floor(k/2) / F) is computed by the host exactly as the reference does (torch, fp32) and
// passed in, so that the arguments of sin / cos are bit-identical to the reference's.
__device__ __forceinline__ float sine_embed_value(const float* __restrict__ xyz, int n_dim, int num_feats, float scale,
                                                  const float* __restrict__ dim_t, int col) {
  const int d = col / num_feats;
  if (d >= n_dim) return 0.f;
  const int k = col - d * num_feats;
  const float arg = __fdiv_rn(__fmul_rn(xyz[d], scale), dim_t[k]);
  return (k & 1) ? cosf(arg) : sinf(arg);
}

__global__ void __launch_bounds__(256) k_sine_embed(const float* __restrict__ xyz, int64_t n_rows, int n_dim, int d_model,
                                                    int num_feats, float scale, const float* __restrict__ dim_t,
                                                    float* __restrict__ out) {
  const int64_t total = n_rows * (int64_t)d_model;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / d_model;
    out[i] = sine_embed_value(xyz + row * n_dim, n_dim, num_feats, scale, dim_t, (int)(i - row * d_model));
  }
}

// One CTA per padded position (j, half): for every cloud b of that half, row (j * B + b) of the padded feature tensor
// [n_max, B, d] receives feats[off[b] + j] (zeros beyond the cloud's length), the padded position embedding its sine
// code, and mask[b, j] = (j >= len[b]).  Threads run over (b, column): coalesced reads of the stacked rows, coalesced
// writes of the padded rows.
__global__ void __launch_bounds__(256) k_pack_coarse(const float* __restrict__ feats, int ld_f, const float* __restrict__ xyz,
                                                     const int64_t* __restrict__ off, int n_pairs, int d_model, int num_feats,
                                                     float scale, const float* __restrict__ dim_t, int ns_max, int nt_max,
                                                     float* __restrict__ src_f, float* __restrict__ tgt_f,
                                                     float* __restrict__ src_pe, float* __restrict__ tgt_pe,
                                                     unsigned char* __restrict__ src_mask, unsigned char* __restrict__ tgt_mask) {
  const int half = blockIdx.x >= ns_max ? 1 : 0;
  const int j = half ? blockIdx.x - ns_max : blockIdx.x;
  const int n_max = half ? nt_max : ns_max;
  float* __restrict__ out_f = half ? tgt_f : src_f;
  float* __restrict__ out_pe = half ? tgt_pe : src_pe;
  unsigned char* __restrict__ out_m = half ? tgt_mask : src_mask;
  const int total = n_pairs * d_model;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int b = i / d_model, c = i - b * d_model;
    const int cloud = half * n_pairs + b;
    const int64_t start = off[cloud];
    const bool valid = (int64_t)j < off[cloud + 1] - start;
    const int64_t row = start + j;
    const int64_t o = ((int64_t)j * n_pairs + b) * d_model + c;
    if (out_f) out_f[o] = valid ? feats[row * ld_f + c] : 0.f;
    if (out_pe) out_pe[o] = valid ? sine_embed_value(xyz + row * 3, 3, num_feats, scale, dim_t, c) : 0.f;
    if (c == 0 && out_m) out_m[(int64_t)b * n_max + j] = valid ? 0 : 1;
  }
}

// out_pts[i] = pts[perm[i]], out_mask[i] = mask[perm[i]], rev[perm[i]] = i (rev pre-filled with -1 by the caller's memset)
__global__ void __launch_bounds__(256) k_shuffle_gather(const float* __restrict__ pts, const unsigned char* __restrict__ mask,
                                                        const int64_t* __restrict__ perm, int64_t n_out, int64_t n_in,
                                                        float* __restrict__ out_pts, unsigned char* __restrict__ out_mask,
                                                        int64_t* __restrict__ rev, int32_t* __restrict__ status) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = perm[i];
    if (j < 0 || j >= n_in) { atomicMax(status, (int32_t)KPREG_E_RANGE); continue; }
    out_pts[3 * i] = pts[3 * j];
    out_pts[3 * i + 1] = pts[3 * j + 1];
    out_pts[3 * i + 2] = pts[3 * j + 2];
    if (mask) out_mask[i] = mask[j];
    if (rev) rev[j] = i;
  }
}

// pair p: (a, b) = (rev_src[corr[0, p]], rev_tgt[corr[1, p]]); keep[p] = both >= 0 (the caller compacts in order)
__global__ void __launch_bounds__(256) k_remap_pairs(const int64_t* __restrict__ corr, int64_t n_pairs, const int64_t* __restrict__ rev_src,
                                                     int64_t n_src, const int64_t* __restrict__ rev_tgt, int64_t n_tgt,
                                                     int64_t* __restrict__ out, unsigned char* __restrict__ keep) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = corr[p], t = corr[n_pairs + p];
    const int64_t a = (s >= 0 && s < n_src) ? rev_src[s] : -1, b = (t >= 0 && t < n_tgt) ? rev_tgt[t] : -1;
    out[p] = a;
    out[n_pairs + p] = b;
    keep[p] = (a >= 0 && b >= 0) ? 1 : 0;
  }
}

int grid_for(int64_t n, int per_block) {
  int blocks = ceil_div(n > 0 ? n : 1, per_block);
  return blocks > 32 * kNumSMs ? 32 * kNumSMs : blocks;
}

}  // namespace
}  // namespace kpreg

using namespace kpreg;

extern "C" int kpreg_overlap_pool(const float* level, const void* idx, int idx64, int64_t n_q, int64_t n_s, int n_nbrs,
                                  float* out, void* stream_) {
  if (n_q < 0 || n_s < 0 || n_nbrs < 0) return KPREG_E_INVALID;
  if (n_q == 0) return KPREG_OK;
  if (!out || (n_nbrs > 0 && !idx) || (n_s > 0 && !level)) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int blocks = ceil_div(n_q * 32, 256);
  ProfScope prof(KPREG_FAM_OTHER, stream);
  if (idx64) k_overlap_pool<int64_t><<<blocks, 256, 0, stream>>>(level, static_cast<const int64_t*>(idx), n_q, n_s, n_nbrs, out);
  else k_overlap_pool<int32_t><<<blocks, 256, 0, stream>>>(level, static_cast<const int32_t*>(idx), n_q, n_s, n_nbrs, out);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

extern "C" int kpreg_sine_embed(const float* xyz, int64_t n_rows, int n_dim, int d_model, int num_feats, float scale,
                                const float* dim_t, float* out, void* stream_) {
  if (n_rows < 0 || n_dim < 1 || d_model < 1 || num_feats < 0 || n_dim * num_feats > d_model) return KPREG_E_INVALID;
  if (n_rows == 0) return KPREG_OK;
  if (!xyz || !out || (num_feats > 0 && !dim_t)) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  ProfScope prof(KPREG_FAM_OTHER, stream);
  k_sine_embed<<<grid_for(n_rows * (int64_t)d_model, 256), 256, 0, stream>>>(xyz, n_rows, n_dim, d_model, num_feats, scale, dim_t, out);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

extern "C" int kpreg_pack_coarse(const float* feats, int ld_f, const float* xyz, const int32_t* lens, int n_pairs, int d_model,
                                 int num_feats, float scale, const float* dim_t, int ns_max, int nt_max, float* src_feats,
                                 float* tgt_feats, float* src_pe, float* tgt_pe, unsigned char* src_mask, unsigned char* tgt_mask,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
  if (n_pairs < 1 || d_model < 1 || ns_max < 0 || nt_max < 0 || !lens || !workspace) return KPREG_E_INVALID;
  if ((src_feats || tgt_feats) && (!feats || ld_f < d_model)) return KPREG_E_INVALID;
  if ((src_pe || tgt_pe) && (!xyz || num_feats < 0 || 3 * num_feats > d_model || (num_feats > 0 && !dim_t))) return KPREG_E_INVALID;
  if (workspace_bytes < sizeof(int64_t) * (size_t)(2 * n_pairs + 1)) return KPREG_E_WORKSPACE;
  if (ns_max + nt_max == 0) return KPREG_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  int64_t* off = static_cast<int64_t*>(workspace);
  int rc = launch_cloud_offsets(lens, 2 * n_pairs, off, stream);
  if (rc) return rc;
  ProfScope prof(KPREG_FAM_OTHER, stream);
  k_pack_coarse<<<ns_max + nt_max, 256, 0, stream>>>(feats, ld_f, xyz, off, n_pairs, d_model, num_feats, scale, dim_t, ns_max, nt_max,
                                                    src_feats, tgt_feats, src_pe, tgt_pe, src_mask, tgt_mask);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

extern "C" int kpreg_pack_coarse_workspace_bytes(int n_pairs, size_t* bytes) {
  if (!bytes || n_pairs < 0) return KPREG_E_INVALID;
  *bytes = align_up(sizeof(int64_t) * (size_t)(2 * n_pairs + 1), 256);
  return KPREG_OK;
}

extern "C" int kpreg_shuffle_gather(const float* pts, const unsigned char* mask, const int64_t* perm, int64_t n_out, int64_t n_in,
                                    float* out_pts, unsigned char* out_mask, int64_t* rev, int32_t* status, void* stream_) {
  if (n_out < 0 || n_in < 0 || !status) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rev && n_in > 0) KP_CUDA_TRY(cudaMemsetAsync(rev, 0xff, sizeof(int64_t) * (size_t)n_in, stream));  // -1
  if (n_out == 0) return KPREG_OK;
  if (!pts || !perm || !out_pts || (mask && !out_mask)) return KPREG_E_INVALID;
  ProfScope prof(KPREG_FAM_OTHER, stream);
  k_shuffle_gather<<<grid_for(n_out, 256), 256, 0, stream>>>(pts, mask, perm, n_out, n_in, out_pts, out_mask, rev, status);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

extern "C" int kpreg_remap_pairs(const int64_t* corr, int64_t n_pairs, const int64_t* rev_src, int64_t n_src, const int64_t* rev_tgt,
                                 int64_t n_tgt, int64_t* out, unsigned char* keep, void* stream_) {
  if (n_pairs < 0 || n_src < 0 || n_tgt < 0) return KPREG_E_INVALID;
  if (n_pairs == 0) return KPREG_OK;
  if (!corr || !rev_src || !rev_tgt || !out || !keep) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  ProfScope prof(KPREG_FAM_OTHER, stream);
  k_remap_pairs<<<grid_for(n_pairs, 256), 256, 0, stream>>>(corr, n_pairs, rev_src, n_src, rev_tgt, n_tgt, out, keep);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}
