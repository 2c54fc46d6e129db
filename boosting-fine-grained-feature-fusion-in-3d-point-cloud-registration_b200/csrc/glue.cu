// Encoder-block glue on sm_100a: the dense layers and per-cloud normalisation that
// sit around KPConv inside the reference's blocks.
//
//   kpreg_linear_forward        y = act((x W^T) * col_scale + col_shift + residual)
//       replaces nn.Linear(bias=False) [+ eval-mode nn.BatchNorm1d folded into col_scale/col_shift] [+ ReLU]
//       of UnaryBlock.mlp (reference finegrained_kpconv_blocks.py:521-555) and of my_Bottle2neck's
//       conv1/bn1, convs[i]/bns[i], conv3/bn3, downsample (reference res2net.py:84-159, 231-265);
//       runs on the tcgen05 3xTF32 GEMM of kpconv_gemm.cu (fp32-grade accuracy), fp32 CUDA cores otherwise.
//   kpreg_segment_norm_forward  y = act((x - mean_c) * rstd_c + residual) per cloud c and channel
//       replaces BatchNormBlock's per-cloud nn.InstanceNorm1d (affine=False, eps 1e-5, biased variance;
//       reference finegrained_kpconv_blocks.py:462-518, a Python loop over clouds there), optionally fused
//       with the LeakyReLU / shortcut-add that follows it (:552-554, :632-634, :712-725).
#include "common.cuh"

namespace kpreg {

int launch_gemm_tc(const float* a, int lda, const float* w_split, float* c, int ldc, int64_t m, int kd, int n,
                   const float* row_scale, const float* col_scale, const float* col_shift, const float* residual, int ld_res,
                   int act, float slope, float* out2, int ld2, const float* addend, int ld_add, const float* post_res, int ld_post,
                   int post_act, cudaStream_t stream);
int launch_gemm_tc_pair(const float* a, int lda, int k1, const float* a2, int lda2, const float* w_split, float* c, int ldc, int64_t m,
                        int kd, int n, const float* row_scale, const float* col_scale, const float* col_shift, const float* residual,
                        int ld_res, int act, float slope, float* out2, int ld2, const float* addend, int ld_add, const float* post_res,
                        int ld_post, int post_act, cudaStream_t stream);
int kpconv_gemm_tc_prepare_weights(const float* weights, int kd, int n, int transpose, float* w_split, cudaStream_t stream);
size_t kpconv_gemm_tc_weight_bytes(int kd, int n);
bool gemm_tc_supported(int64_t m, int kd, int n, int lda, const void* a);
int gemm_tc_block_n(int kd, int n);
int launch_gemm_tn(const float* a, int lda, const float* b, int ldb, const float* row_scale, float* c, int ldc, int64_t k_rows,
                   int ma_dim, int nb_dim, int c_transposed, void* split_ws, cudaStream_t stream);
size_t gemm_tn_workspace_bytes(int64_t k_rows, int nb_dim);
bool gemm_tn_supported(int64_t k_rows, int ma_dim, int nb_dim, int lda, const void* a);

namespace {

__device__ __forceinline__ float activate(float x, int act, float slope) {
  if (act == 1) return fmaxf(x, 0.f);
  if (act == 2) return x > 0.f ? x : x * slope;
  return x;
}

// fp32 CUDA-core fallback: one thread per output element (shapes the TMA path cannot address are tiny).
__global__ void __launch_bounds__(256) k_linear_simple(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                                       int64_t m_rows, int k_dim, int n_dim, const float* __restrict__ col_scale,
                                                       const float* __restrict__ col_shift, const float* __restrict__ residual,
                                                       int ld_res, int act, float slope, float* __restrict__ out, int ldc,
                                                       float* __restrict__ out2, int ld2, const float* __restrict__ addend,
                                                       int ld_add, const float* __restrict__ post_res, int ld_post, int post_act) {
  const int64_t total = m_rows * (int64_t)n_dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / n_dim;
    const int n = (int)(i - m * n_dim);
    float acc = 0.f;
    for (int k = 0; k < k_dim; ++k) acc = fmaf(x[m * ldx + k], w[(int64_t)n * k_dim + k], acc);
    if (col_scale) acc *= col_scale[n];
    if (col_shift) acc += col_shift[n];
    if (residual) acc += residual[m * ld_res + n];
    acc = activate(acc, act, slope);
    if (post_res) acc = activate(acc + post_res[m * ld_post + n], post_act, slope);
    out[m * ldc + n] = acc;
    if (out2) out2[m * ld2 + n] = acc + addend[m * ld_add + n];
  }
}

// One launch in front of the statistics pass: cloud offsets (a serial prefix over a few hundred lengths) and the zeroed
// fp64 moment buffers — instead of an offsets kernel plus one or two memsets per normalisation (25 normalisations per step).
__global__ void __launch_bounds__(256) k_norm_prologue(const int32_t* __restrict__ lens, int n_clouds, int64_t* __restrict__ off,
                                                       double* __restrict__ zero_a, double* __restrict__ zero_b, int64_t n_zero) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t acc = 0;
    for (int c = 0; c < n_clouds; ++c) {
      off[c] = acc;
      acc += lens[c] > 0 ? lens[c] : 0;
    }
    off[n_clouds] = acc;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += (int64_t)gridDim.x * blockDim.x) {
    zero_a[i] = 0.0;
    if (zero_b) zero_b[i] = 0.0;
  }
}

constexpr int kStatRows = 256;  // rows per CTA
constexpr int kStatCh = 64;     // channels per CTA: 16 threads x float4; 16 row groups

__device__ __forceinline__ void stat_flush(double* __restrict__ stats, int cloud, int channels, int ch, const double (&sum)[4],
                                           const double (&sq)[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    double* dst = stats + ((int64_t)cloud * channels + ch + e) * 2;
    atomicAdd(dst, sum[e]);
    atomicAdd(dst + 1, sq[e]);
  }
}

// stats[(cloud * C + ch) * 2 + {0,1}] += {sum x, sum x^2} in fp64.  Each thread streams float4 (4 channels) of 16
// rows; a CTA covers 256 rows x 64 channels and issues one atomic pair per channel and cloud segment of its rows.
// TX = threads per row (8 for rows of <= 32 channels: no idle lanes), 256 / TX row groups.
template <int TX>
__global__ void __launch_bounds__(256) k_segnorm_stats(const float* __restrict__ x, int ldx, const int64_t* __restrict__ off,
                                                       int n_clouds, int64_t n_rows, int channels, double* __restrict__ stats,
                                                       int rows_per_cta = kStatRows) {
  constexpr int RY = 256 / TX, CH = 4 * TX;
  __shared__ double s_sum[RY][CH], s_sq[RY][CH];
  const int cx = threadIdx.x % TX, ry = threadIdx.x / TX;
  const int ch = blockIdx.y * CH + cx * 4;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(n_rows, r0 + rows_per_cta);
  const int c_first = cloud_of(off, n_clouds, r0);
  const int c_last = cloud_of(off, n_clouds, r1 - 1);
  const bool live = ch < channels;  // channels is a multiple of 4
  // one pass per cloud segment of the chunk (coarse levels: clouds of a few hundred rows, most chunks hold two or three):
  // per-thread partial sums -> shared-memory reduction over the row groups -> one atomic pair per channel and segment
  for (int c = c_first; c <= c_last; ++c) {
    const int64_t s0 = max(r0, off[c]), s1 = min(r1, off[c + 1]);
    double sum[4] = {0.0, 0.0, 0.0, 0.0}, sq[4] = {0.0, 0.0, 0.0, 0.0};
    if (live) {
#pragma unroll 4
      for (int64_t r = s0 + ry; r < s1; r += RY) {
        const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + ch);
        sum[0] += (double)v.x; sq[0] += (double)v.x * (double)v.x;
        sum[1] += (double)v.y; sq[1] += (double)v.y * (double)v.y;
        sum[2] += (double)v.z; sq[2] += (double)v.z * (double)v.z;
        sum[3] += (double)v.w; sq[3] += (double)v.w * (double)v.w;
      }
    }
    if (c != c_first) __syncthreads();  // the previous segment's reduction has read the tiles
#pragma unroll
    for (int e = 0; e < 4; ++e) { s_sum[ry][cx * 4 + e] = sum[e]; s_sq[ry][cx * 4 + e] = sq[e]; }
    __syncthreads();
    if (threadIdx.x < CH && s1 > s0) {
      const int cc = blockIdx.y * CH + threadIdx.x;
      if (cc < channels) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int g = 0; g < RY; ++g) { a += s_sum[g][threadIdx.x]; b += s_sq[g][threadIdx.x]; }
        double* dst = stats + ((int64_t)c * channels + cc) * 2;
        atomicAdd(dst, a);
        atomicAdd(dst + 1, b);
      }
    }
  }
}

// mean / rstd per (cloud, channel) from the fp64 moments: biased variance, eps inside the sqrt.
__global__ void __launch_bounds__(256) k_segnorm_finalize(const double* __restrict__ stats, const int64_t* __restrict__ off,
                                                          int n_clouds, int channels, float eps, float2* __restrict__ mr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_clouds * channels) return;
  const int c = i / channels;
  const double n = (double)max((int64_t)1, off[c + 1] - off[c]);
  const double mean = stats[2 * i] / n;
  double var = stats[2 * i + 1] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  mr[i] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

// y = act((x - mean) * rstd + residual).  Same tiling as k_segnorm_stats (a CTA = 256 rows x 64 channels, a thread = 4
// channels of every 16th row), so a thread's channels are fixed and its cloud changes at most a few times: mean / rstd are
// derived from the fp64 moments in registers when the cloud changes (no separate finalize launch, no per-element
// cloud search, no per-element statistics loads) — biased variance, eps inside the sqrt.  TX = threads per row
// (4 TX channels per CTA column): 8 for narrow rows so that no lane idles at 32 channels.
template <int TX>
__global__ void __launch_bounds__(256) k_segnorm_apply(const float* __restrict__ x, int ldx, const int64_t* __restrict__ off,
                                                       int n_clouds, int64_t n_rows, int channels, const double* __restrict__ stats,
                                                       float eps, const float* __restrict__ residual, int ld_res, int act, float slope,
                                                       float* __restrict__ out, int ldo, int rows_per_cta,
                                                       unsigned char* __restrict__ row_pos) {
  constexpr int RY = 256 / TX;
  const int cx = threadIdx.x % TX, ry = threadIdx.x / TX;
  const int ch = blockIdx.y * (4 * TX) + cx * 4;
  if (ch >= channels) return;  // channels is a multiple of 4
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(n_rows, r0 + rows_per_cta);
  int c = -1;
  int64_t c_end = 0;
  float mean[4], rstd[4];
  for (int64_t r = r0 + ry; r < r1; r += RY) {
    if (c < 0 || r >= c_end) {
      c = cloud_of(off, n_clouds, r);
      c_end = off[c + 1];
      // (one fp64 division per cloud change; the moments and the cancellation-prone E[x^2] - mean^2 stay in fp64, the
      // reciprocal square root runs in fp32 — at the coarse levels a thread changes cloud every few rows)
      const double inv_n = 1.0 / (double)max((int64_t)1, c_end - off[c]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double* st = stats + ((int64_t)c * channels + ch + e) * 2;
        const double m = st[0] * inv_n;
        double var = st[1] * inv_n - m * m;
        if (var < 0.0) var = 0.0;
        mean[e] = (float)m;
        rstd[e] = 1.0f / sqrtf((float)(var + (double)eps));
      }
    }
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + ch);
    float4 y = make_float4((v.x - mean[0]) * rstd[0], (v.y - mean[1]) * rstd[1], (v.z - mean[2]) * rstd[2], (v.w - mean[3]) * rstd[3]);
    if (residual) {
      const float4 q = *reinterpret_cast<const float4*>(residual + r * ld_res + ch);
      y.x += q.x; y.y += q.y; y.z += q.z; y.w += q.w;
    }
    y.x = activate(y.x, act, slope); y.y = activate(y.y, act, slope);
    y.z = activate(y.z, act, slope); y.w = activate(y.w, act, slope);
    *reinterpret_cast<float4*>(out + r * ldo + ch) = y;
    if (row_pos) {
      // KPConv's normalisation predicate of the row just written (sum of its features > 0, fp64 sum — the arithmetic and
      // the summation order of k_row_positive_vec): the launcher passes row_pos only when channels == 4 * TX, i.e. the TX
      // consecutive lanes of this row hold all of it and leave the loop together
      double s = ((double)y.x + (double)y.y) + ((double)y.z + (double)y.w);
      const unsigned int grp = (TX == 32 ? 0xffffffffu : ((1u << TX) - 1u)) << ((threadIdx.x & 31) & ~(TX - 1));
#pragma unroll
      for (int o = TX / 2; o > 0; o >>= 1) s += __shfl_xor_sync(grp, s, o);
      if (cx == 0) row_pos[r] = (float)s > 0.0f ? 1 : 0;
    }
  }
}

// Backward of the per-cloud instance norm, y = (x - mean) * rstd:
//   dx = rstd * (dy - mean_n(dy) - xhat * mean_n(dy * xhat)),  xhat = (x - mean) * rstd, means over the cloud's rows.
// bstats[(cloud * C + ch) * 2 + {0,1}] += {sum dy, sum dy * xhat} in fp64 — same tiling as k_segnorm_stats.
__global__ void __launch_bounds__(256) k_segnorm_bstats(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int ld_dy,
                                                        const int64_t* __restrict__ off, int n_clouds, int64_t n_rows, int channels,
                                                        const float2* __restrict__ mr, double* __restrict__ bstats) {
  const int cx = threadIdx.x & 15, ry = threadIdx.x >> 4;
  const int ch = blockIdx.y * kStatCh + cx * 4;
  if (ch >= channels) return;  // channels is a multiple of 4
  const int64_t r0 = (int64_t)blockIdx.x * kStatRows;
  const int64_t r1 = min(n_rows, r0 + kStatRows);
  int c = -1;
  float2 m[4];
  double sum[4] = {0.0, 0.0, 0.0, 0.0}, sq[4] = {0.0, 0.0, 0.0, 0.0};
  const bool one_cloud = cloud_of(off, n_clouds, r0) == cloud_of(off, n_clouds, r1 - 1);
  for (int64_t r = r0 + ry; r < r1; r += 16) {
    const int cr = (one_cloud && c >= 0) ? c : cloud_of(off, n_clouds, r);
    if (cr != c) {
      if (c >= 0) stat_flush(bstats, c, channels, ch, sum, sq);
      c = cr;
#pragma unroll
      for (int e = 0; e < 4; ++e) { sum[e] = sq[e] = 0.0; m[e] = mr[(int64_t)c * channels + ch + e]; }
    }
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + ch);
    const float4 g = *reinterpret_cast<const float4*>(dy + r * ld_dy + ch);
    sum[0] += (double)g.x; sq[0] += (double)g.x * (double)((v.x - m[0].x) * m[0].y);
    sum[1] += (double)g.y; sq[1] += (double)g.y * (double)((v.y - m[1].x) * m[1].y);
    sum[2] += (double)g.z; sq[2] += (double)g.z * (double)((v.z - m[2].x) * m[2].y);
    sum[3] += (double)g.w; sq[3] += (double)g.w * (double)((v.w - m[3].x) * m[3].y);
  }
  if (c >= 0) stat_flush(bstats, c, channels, ch, sum, sq);
}

__global__ void __launch_bounds__(256) k_segnorm_bapply(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int ld_dy,
                                                        const int64_t* __restrict__ off, int n_clouds, int64_t n_rows, int channels,
                                                        const float2* __restrict__ mr, const double* __restrict__ bstats,
                                                        float* __restrict__ dx, int ld_dx) {
  const int c4 = channels >> 2;
  const int64_t total = n_rows * (int64_t)c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int ch = (int)(i - r * c4) * 4;
    const int c = cloud_of(off, n_clouds, r);
    const float inv_n = 1.0f / (float)max((int64_t)1, off[c + 1] - off[c]);
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + ch);
    const float4 g = *reinterpret_cast<const float4*>(dy + r * ld_dy + ch);
    const float vv[4] = {v.x, v.y, v.z, v.w}, gg[4] = {g.x, g.y, g.z, g.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 m = mr[(int64_t)c * channels + ch + e];
      const double* b = bstats + ((int64_t)c * channels + ch + e) * 2;
      const float xhat = (vv[e] - m.x) * m.y;
      o[e] = m.y * (gg[e] - (float)b[0] * inv_n - xhat * ((float)b[1] * inv_n));
    }
    *reinterpret_cast<float4*>(dx + r * ld_dx + ch) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

struct NormWs { int64_t* off; double* stats; float2* mr; double* bstats; size_t total; };
NormWs carve_norm(void* base, int n_clouds, int channels) {
  NormWs w;
  Carver cv(base);
  w.off = cv.take<int64_t>((size_t)n_clouds + 1);
  w.stats = cv.take<double>((size_t)n_clouds * channels * 2);
  w.mr = cv.take<float2>((size_t)n_clouds * channels);
  w.bstats = cv.take<double>((size_t)n_clouds * channels * 2);
  w.total = align_up(cv.used, 256);
  return w;
}

}  // namespace
}  // namespace kpreg

using namespace kpreg;

extern "C" int kpreg_gemm_supported(int64_t m_rows, int k_dim, int n_dim, int ldx, const void* x) {
  return gemm_tc_supported(m_rows, k_dim, n_dim, ldx, x) ? 1 : 0;
}

extern "C" int kpreg_linear_tile_cols(int k_dim, int n_dim) {
  if (k_dim < 1 || n_dim < 1) return 0;
  return gemm_tc_block_n(k_dim, n_dim);
}

extern "C" int kpreg_linear_workspace_bytes(int k_dim, int n_dim, size_t* bytes) {
  if (!bytes || k_dim < 1 || n_dim < 1) return KPREG_E_INVALID;
  *bytes = kpconv_gemm_tc_weight_bytes(k_dim, n_dim) + 256;
  return KPREG_OK;
}

// Split a weight matrix once (inference: weights do not change between calls) into the hi/lo TF32 operand pair
// the tensor-core GEMM consumes.  weight is [n_dim, k_dim] (transpose = 0, nn.Linear) or [k_dim, n_dim]
// (transpose = 1, KPConv's [K*c_in, c_out]).  out must hold kpreg_linear_workspace_bytes(k_dim, n_dim) bytes.
extern "C" int kpreg_split_weights(const float* weight, int k_dim, int n_dim, int transpose, void* out, size_t out_bytes,
                                   void* stream_) {
  if (!weight || !out || k_dim < 1 || n_dim < 1) return KPREG_E_INVALID;
  if (out_bytes < kpconv_gemm_tc_weight_bytes(k_dim, n_dim)) return KPREG_E_WORKSPACE;
  return kpconv_gemm_tc_prepare_weights(weight, k_dim, n_dim, transpose ? 1 : 0, static_cast<float*>(out), (cudaStream_t)stream_);
}

extern "C" int kpreg_linear_forward(const float* x, int ldx, const float* weight, int64_t m_rows, int k_dim, int n_dim,
                                    const float* col_scale, const float* col_shift, const float* residual, int ld_res, int act,
                                    float slope, float* out, int ldc, float* out2, int ld2, const float* addend, int ld_add,
                                    const float* post_residual, int ld_post, int post_act, int gemm, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
  if (m_rows < 0 || k_dim < 1 || n_dim < 1 || ldx < k_dim || ldc < n_dim || act < 0 || act > 2) return KPREG_E_INVALID;
  if (post_act < 0 || post_act > 2 || (post_residual && ld_post < n_dim)) return KPREG_E_INVALID;
  if (m_rows == 0) return KPREG_OK;
  if (!x || !weight || !out || (out2 && !addend)) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  ProfScope prof(KPREG_FAM_LINEAR, stream);
  if (gemm == 2) {
    // `weight` already is the split operand pair produced by kpreg_split_weights
    if (!gemm_tc_supported(m_rows, k_dim, n_dim, ldx, x)) return KPREG_E_INVALID;
    return launch_gemm_tc(x, ldx, weight, out, ldc, m_rows, k_dim, n_dim, nullptr, col_scale, col_shift, residual, ld_res, act,
                          slope, out2, ld2, addend, ld_add, post_residual, ld_post, post_act, stream);
  }
  if (gemm == 1 && gemm_tc_supported(m_rows, k_dim, n_dim, ldx, x)) {
    if (!workspace || workspace_bytes < kpconv_gemm_tc_weight_bytes(k_dim, n_dim)) return KPREG_E_WORKSPACE;
    float* w_split = static_cast<float*>(workspace);
    int rc = kpconv_gemm_tc_prepare_weights(weight, k_dim, n_dim, 0, w_split, stream);
    if (rc) return rc;
    return launch_gemm_tc(x, ldx, w_split, out, ldc, m_rows, k_dim, n_dim, nullptr, col_scale, col_shift, residual, ld_res, act,
                          slope, out2, ld2, addend, ld_add, post_residual, ld_post, post_act, stream);
  }
  int blocks = ceil_div(m_rows * (int64_t)n_dim, 256);
  if (blocks > 32 * kNumSMs) blocks = 32 * kNumSMs;
  k_linear_simple<<<blocks, 256, 0, stream>>>(x, ldx, weight, m_rows, k_dim, n_dim, col_scale, col_shift, residual, ld_res, act,
                                              slope, out, ldc, out2, ld2, addend, ld_add, post_residual, ld_post, post_act);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

extern "C" int kpreg_linear_pair_forward(const float* x1, int ld1, int k1, const float* x2, int ld2, int k2, const float* w_split,
                                         int64_t m_rows, int n_dim, const float* col_scale, const float* col_shift, int act,
                                         float slope, const float* post_residual, int ld_post, int post_act, float* out, int ldc,
                                         void* stream_) {
  if (m_rows < 0 || k1 < 4 || k2 < 1 || n_dim < 8 || ld1 < k1 || ld2 < k2 || ldc < n_dim || act < 0 || act > 2) return KPREG_E_INVALID;
  if (post_act < 0 || post_act > 2 || (post_residual && ld_post < n_dim)) return KPREG_E_INVALID;
  if (m_rows == 0) return KPREG_OK;
  if (!x1 || !x2 || !w_split || !out) return KPREG_E_INVALID;
  const int kd = (k1 + 31) / 32 * 32 + k2;
  if (!gemm_tc_supported(m_rows, kd, n_dim, ld1, x1)) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  ProfScope prof(KPREG_FAM_LINEAR, stream);
  return launch_gemm_tc_pair(x1, ld1, k1, x2, ld2, w_split, out, ldc, m_rows, kd, n_dim, nullptr, col_scale, col_shift, nullptr, 0, act,
                             slope, nullptr, 0, nullptr, 0, post_residual, ld_post, post_act, stream);
}

extern "C" int kpreg_segment_norm_workspace_bytes(int n_clouds, int channels, size_t* bytes) {
  if (!bytes || n_clouds < 1 || channels < 1) return KPREG_E_INVALID;
  *bytes = carve_norm(nullptr, n_clouds, channels).total;
  return KPREG_OK;
}

extern "C" int kpreg_segment_norm_forward(const float* x, int ldx, const int32_t* lens, int n_clouds, int64_t n_rows,
                                          int channels, float eps, const float* residual, int ld_res, int act, float slope,
                                          float* out, int ldo, void* workspace, size_t workspace_bytes, void* stream_) {
  return kpreg_segment_norm_forward_rowpos(x, ldx, lens, n_clouds, n_rows, channels, eps, residual, ld_res, act, slope, out, ldo,
                                           nullptr, workspace, workspace_bytes, stream_);
}

extern "C" int kpreg_segment_norm_rowpos_supported(int channels) { return channels == 32 || channels == 64; }

extern "C" int kpreg_segment_norm_forward_rowpos(const float* x, int ldx, const int32_t* lens, int n_clouds, int64_t n_rows,
                                                 int channels, float eps, const float* residual, int ld_res, int act, float slope,
                                                 float* out, int ldo, unsigned char* row_pos, void* workspace,
                                                 size_t workspace_bytes, void* stream_) {
  if (row_pos && !kpreg_segment_norm_rowpos_supported(channels)) return KPREG_E_INVALID;
  if (n_rows < 0 || n_clouds < 1 || channels < 4 || (channels & 3) || (ldx & 3) || (ldo & 3) || act < 0 || act > 2) return KPREG_E_INVALID;
  if (residual && (ld_res & 3)) return KPREG_E_INVALID;
  if (n_rows == 0) return KPREG_OK;
  if (!x || !lens || !out || !workspace) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  NormWs w = carve_norm(workspace, n_clouds, channels);
  if (w.total > workspace_bytes) return KPREG_E_WORKSPACE;
  ProfScope prof(KPREG_FAM_NORM, stream);
  {
    const int64_t n_zero = (int64_t)n_clouds * channels * 2;
    int pb = ceil_div(n_zero, 256);
    if (pb > 4 * kNumSMs) pb = 4 * kNumSMs;
    k_norm_prologue<<<pb, 256, 0, stream>>>(lens, n_clouds, w.off, w.stats, nullptr, n_zero);
    KP_LAUNCH_CHECK();
  }
  // rows per CTA: a 256-row chunk of a narrow tensor is 32-64 KB, which an SM streams in under a microsecond — the CTA's
  // fixed work (two cloud searches, the shared-memory reduction, its atomics) then sets the rate; large row counts get longer chunks
  static const int rows_env = [] { const char* e = getenv("KPREG_NORM_ROWS"); return e ? atoi(e) : 0; }();  // A/B measurements
  int rows = kStatRows;
  const int col_ctas = ceil_div(channels, channels <= 32 ? 32 : kStatCh);
  while (rows < 2048 && ceil_div(n_rows, 2 * rows) * (int64_t)col_ctas >= (int64_t)kNumSMs * 16) rows *= 2;  // (profiles/r3_sweep_norm_rows.txt)
  if (rows_env >= 256) rows = rows_env;
  dim3 grid((unsigned)ceil_div(n_rows, rows), (unsigned)col_ctas);
  if (channels <= 32) k_segnorm_stats<8><<<grid, 256, 0, stream>>>(x, ldx, w.off, n_clouds, n_rows, channels, w.stats, rows);
  else k_segnorm_stats<16><<<grid, 256, 0, stream>>>(x, ldx, w.off, n_clouds, n_rows, channels, w.stats, rows);
  KP_LAUNCH_CHECK();
  if (channels <= 32)
    k_segnorm_apply<8><<<grid, 256, 0, stream>>>(x, ldx, w.off, n_clouds, n_rows, channels, w.stats, eps, residual, ld_res, act, slope, out, ldo, rows, row_pos);
  else
    k_segnorm_apply<16><<<grid, 256, 0, stream>>>(x, ldx, w.off, n_clouds, n_rows, channels, w.stats, eps, residual, ld_res, act, slope, out, ldo, rows, row_pos);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

extern "C" int kpreg_segment_norm_backward(const float* x, int ldx, const float* dy, int ld_dy, const int32_t* lens, int n_clouds,
                                           int64_t n_rows, int channels, float eps, float* dx, int ld_dx, void* workspace,
                                           size_t workspace_bytes, void* stream_) {
  if (n_rows < 0 || n_clouds < 1 || channels < 4 || (channels & 3) || (ldx & 3) || (ld_dy & 3) || (ld_dx & 3)) return KPREG_E_INVALID;
  if (n_rows == 0) return KPREG_OK;
  if (!x || !dy || !lens || !dx || !workspace) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  NormWs w = carve_norm(workspace, n_clouds, channels);
  if (w.total > workspace_bytes) return KPREG_E_WORKSPACE;
  ProfScope prof(KPREG_FAM_NORM, stream);
  // the forward statistics are recomputed from x (one extra pass; nothing is kept alive between forward and backward)
  {
    const int64_t n_zero = (int64_t)n_clouds * channels * 2;
    int pb = ceil_div(n_zero, 256);
    if (pb > 4 * kNumSMs) pb = 4 * kNumSMs;
    k_norm_prologue<<<pb, 256, 0, stream>>>(lens, n_clouds, w.off, w.stats, w.bstats, n_zero);
    KP_LAUNCH_CHECK();
  }
  dim3 grid((unsigned)ceil_div(n_rows, kStatRows), (unsigned)ceil_div(channels, kStatCh));
  if (channels <= 32) {
    dim3 grid_s((unsigned)ceil_div(n_rows, kStatRows), (unsigned)ceil_div(channels, 32));
    k_segnorm_stats<8><<<grid_s, 256, 0, stream>>>(x, ldx, w.off, n_clouds, n_rows, channels, w.stats);
  } else {
    k_segnorm_stats<16><<<grid, 256, 0, stream>>>(x, ldx, w.off, n_clouds, n_rows, channels, w.stats);
  }
  KP_LAUNCH_CHECK();
  k_segnorm_finalize<<<ceil_div((int64_t)n_clouds * channels, 256), 256, 0, stream>>>(w.stats, w.off, n_clouds, channels, eps, w.mr);
  KP_LAUNCH_CHECK();
  k_segnorm_bstats<<<grid, 256, 0, stream>>>(x, ldx, dy, ld_dy, w.off, n_clouds, n_rows, channels, w.mr, w.bstats);
  KP_LAUNCH_CHECK();
  int blocks = ceil_div(n_rows * (int64_t)(channels >> 2), 256);
  if (blocks > 32 * kNumSMs) blocks = 32 * kNumSMs;
  k_segnorm_bapply<<<blocks, 256, 0, stream>>>(x, ldx, dy, ld_dy, w.off, n_clouds, n_rows, channels, w.mr, w.bstats, dx, ld_dx);
  KP_LAUNCH_CHECK();
  return KPREG_OK;
}

// ---- nn.Linear backward on the tensor cores -----------------------------------------------------------------------
extern "C" int kpreg_linear_backward_workspace_bytes(int64_t m_rows, int k_dim, int n_dim, size_t* bytes) {
  if (!bytes || m_rows < 0 || k_dim < 1 || n_dim < 1) return KPREG_E_INVALID;
  *bytes = align_up(kpconv_gemm_tc_weight_bytes(n_dim, k_dim), 256) + gemm_tn_workspace_bytes(m_rows, k_dim < n_dim ? k_dim : n_dim) + 256;
  return KPREG_OK;
}

extern "C" int kpreg_linear_backward(const float* x, int ldx, const float* dy, int ld_dy, const float* weight, int64_t m_rows,
                                     int k_dim, int n_dim, float* dx, int ld_dx, float* d_weight, void* workspace,
                                     size_t workspace_bytes, void* stream_) {
  if (m_rows < 0 || k_dim < 1 || n_dim < 1 || ldx < k_dim || ld_dy < n_dim) return KPREG_E_INVALID;
  if (!x || !dy || !weight || !workspace || (dx && ld_dx < k_dim)) return KPREG_E_INVALID;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (d_weight) KP_CUDA_TRY(cudaMemsetAsync(d_weight, 0, sizeof(float) * (size_t)n_dim * (size_t)k_dim, stream));
  if (m_rows == 0) return KPREG_OK;
  size_t need = 0;
  kpreg_linear_backward_workspace_bytes(m_rows, k_dim, n_dim, &need);
  if (need > workspace_bytes) return KPREG_E_WORKSPACE;
  // both products must be addressable by the TMA paths; the caller falls back to its own kernels otherwise
  if ((dx && !gemm_tc_supported(m_rows, n_dim, k_dim, ld_dy, dy)) ||
      (d_weight && (!gemm_tn_supported(m_rows, n_dim, k_dim, ld_dy, dy) || !gemm_tn_supported(m_rows, k_dim, n_dim, ldx, x))))
    return KPREG_E_INVALID;
  ProfScope prof(KPREG_FAM_LINEAR, stream);
  float* w_split = static_cast<float*>(workspace);
  void* tn_split = static_cast<char*>(workspace) + align_up(kpconv_gemm_tc_weight_bytes(n_dim, k_dim), 256);
  if (dx) {
    // dx[M, K] = dy[M, N] W[N, K]: the forward kernel with W^T ([k = N, n = K], given as [kd, n] row-major) as its operand
    int rc = kpconv_gemm_tc_prepare_weights(weight, n_dim, k_dim, 1, w_split, stream);
    if (rc) return rc;
    rc = launch_gemm_tc(dy, ld_dy, w_split, dx, ld_dx, m_rows, n_dim, k_dim, nullptr, nullptr, nullptr, nullptr, 0, 0, 0.f, nullptr,
                        0, nullptr, 0, nullptr, 0, 0, stream);
    if (rc) return rc;
  }
  if (d_weight) {
    // d_weight[N, K] = dy^T x, reduction over the rows; the narrower operand is the one copied (pre-split) once
    if (k_dim <= n_dim) return launch_gemm_tn(dy, ld_dy, x, ldx, nullptr, d_weight, k_dim, m_rows, n_dim, k_dim, 0, tn_split, stream);
    return launch_gemm_tn(x, ldx, dy, ld_dy, nullptr, d_weight, k_dim, m_rows, k_dim, n_dim, 1, tn_split, stream);
  }
  return KPREG_OK;
}
