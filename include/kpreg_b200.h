/*
 * kpreg_b200.h — C ABI of libkpreg_b200.so: the KPConv registration hot path on B200 (sm_100a).
 *
 * Drop-in boundary.  Every entry point replaces one native/operator interface of the reference
 * (YHY138/Boosting-Fine-grained-Feature-Fusion-in-3D-Point-Cloud-Registration); citations are
 * relative to that repository.  Conventions:
 *   - every data pointer is a DEVICE pointer owned by the caller (the Python host allocates with
 *     torch); nothing is allocated, retained or freed by the library; scratch comes from a
 *     caller-provided workspace whose size the matching *_workspace_bytes() call returns;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     the host.  Data-dependent sizes (subsampled point counts, neighbour row widths) are written
 *     to device memory; the host reads them when it needs a shape;
 *   - return value 0 = success; non-zero = KPREG_E_* (the Python host raises RuntimeError, the
 *     error convention of the reference's CPython wrappers, cpp_neighbors/wrapper.cpp:75-205);
 *   - point arrays are [N,3] float32 row-major ("stacked clouds", first all points of cloud 0,
 *     then cloud 1, ...) with an int32 length per cloud, exactly the reference's batch layout
 *     (cpp_subsampling/wrapper.cpp:62-333).
 */
#ifndef KPREG_B200_H
#define KPREG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define KPREG_API __attribute__((visibility("default")))
#else
#define KPREG_API
#endif

#define KPREG_OK 0
#define KPREG_E_INVALID 1   /* bad argument (null pointer, negative size, unsupported mode) */
#define KPREG_E_WORKSPACE 2 /* workspace too small */
#define KPREG_E_CUDA 3      /* a CUDA runtime call failed; see kpreg_last_error() */
#define KPREG_E_RANGE 4     /* reported through the device status word: grid too large to index */

/* Library / build identification. */
KPREG_API int kpreg_version(void);                /* 100 * major + minor */
KPREG_API const char* kpreg_last_error(void);     /* text of the last KPREG_E_CUDA on this thread */
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
KPREG_API unsigned long long kpreg_launch_count(void);

/* Optional per-kernel-family device timing (CUDA events on the launching stream), used by bench.py for the
 * roofline figures.  kpreg_profile(1) clears the records and starts recording, kpreg_profile(0) stops;
 * kpreg_profile_read() waits for the recorded events and returns, per family, the summed device time in
 * milliseconds and the number of timed launches.  Off by default (no events are recorded). */
#define KPREG_FAM_SUBSAMPLE 0   /* whole subsample_batch call */
#define KPREG_FAM_GRID_BUILD 1  /* whole grid_build call */
#define KPREG_FAM_GRID_QUERY 2  /* k_grid_query */
#define KPREG_FAM_GATHER 3      /* k_kpconv_gather */
#define KPREG_FAM_CONTRACT 4    /* KPConv contraction GEMM */
#define KPREG_FAM_POOL 5        /* k_max_pool */
#define KPREG_FAM_KABSCH 6      /* k_kabsch */
#define KPREG_FAM_OTHER 7       /* pack rows, row sums, backward kernels */
#define KPREG_FAM_LINEAR 8      /* kpreg_linear_forward */
#define KPREG_FAM_NORM 9        /* kpreg_segment_norm_forward */
#define KPREG_N_FAMILIES 10
KPREG_API int kpreg_profile(int enable);
KPREG_API int kpreg_profile_reserve(int n_records);  /* pre-create the events of n_records timed scopes */
KPREG_API int kpreg_profile_read(double* ms /*[KPREG_N_FAMILIES]*/, unsigned long long* launches /*[KPREG_N_FAMILIES]*/);

/* ---------------------------------------------------------------------------------------------
 * Voxel-grid barycentre subsampling of a stacked batch.
 * Replaces batch_grid_subsampling()            cpp_subsampling/grid_subsampling/grid_subsampling.cpp:109-211
 * behind cpp_subsampling.subsample_batch()      cpp_subsampling/wrapper.cpp:62-333 (points-only branch,
 * the only one the repository uses: models/backbone_kpconv/finegrained_kpconv.py:371).
 *
 * out_pts   [N,3] capacity; the first out_counts[n_clouds] rows are valid, cloud after cloud, each
 *           cloud in the reference's std::unordered_map iteration order with fp32 barycentres that
 *           are bit-identical to the reference's.
 * out_counts[n_clouds+2]: per-cloud subsampled counts, then their total M, then a status word
 *           (0 or KPREG_E_RANGE).
 * max_p     < 1 = keep everything (grid_subsampling.cpp:134-135); else keep the first max_p per cloud.
 * ------------------------------------------------------------------------------------------- */
KPREG_API int kpreg_subsample_workspace_bytes(int64_t n_points, int n_clouds, size_t* bytes);
KPREG_API int kpreg_subsample_batch(const float* pts, const int32_t* lens, int64_t n_points, int n_clouds,
                          float sample_dl, int max_p, float* out_pts, int32_t* out_counts,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Radius neighbours of a stacked batch.
 * Replaces batch_nanoflann_neighbors()          cpp_neighbors/neighbors/neighbors.cpp:211-332
 * behind cpp_neighbors.batch_query()             cpp_neighbors/wrapper.cpp:58-238
 * and batch_neighbors_kpconv()'s truncation      models/backbone_kpconv/finegrained_kpconv.py:248-263.
 *
 * A support set is binned once into a cell grid (kpreg_grid_build) and can then serve any number
 * of query sets with radius <= the grid's cell size (conv, pool and upsample tables of one pyramid
 * level share their supports).  kpreg_grid_query takes the same n_supports / n_clouds the grid was
 * built with; the query batch must have the same number of clouds.
 *
 * kpreg_grid_query writes, for query i, the support indices (global = local + cloud offset) with
 * fp32 d2 = (dx*dx + dy*dy) + dz*dz < radius*radius, ascending by (d2, index), truncated to
 * `width` columns and padded with n_supports — the reference's row format.
 *   out_idx     [n_queries, width] int32 (or int64 when idx64 != 0)
 *   out_counts  optional [n_queries] int32: untruncated neighbour count per query
 *   out_stats   [2] int32, accumulated with max: {max untruncated count, status}; zero it first.
 * The reference's row width is min(out_stats[0], limit).
 *
 * Processing order.  kpreg_grid_build can return the cell-sorted permutation of its supports (out_order,
 * int32 [n_supports], may be NULL).  kpreg_grid_query, kpreg_kpconv_forward/backward and kpreg_max_pool_forward
 * accept such a permutation of their QUERY rows as `order` (may be NULL): it only changes which warp handles which
 * row — spatially adjacent queries share neighbours, which turns L2 gathers into L1 hits — never the results.
 * ------------------------------------------------------------------------------------------- */
KPREG_API int kpreg_grid_workspace_bytes(int64_t n_supports, int n_clouds, size_t* bytes);
KPREG_API int kpreg_grid_build(const float* supports, const int32_t* s_lens, int64_t n_supports, int n_clouds,
                     float cell, void* grid, size_t grid_bytes, int32_t* out_order, void* stream);
KPREG_API int kpreg_grid_query(const void* grid, int64_t n_supports, int n_clouds, const float* queries,
                     const int32_t* q_lens, int64_t n_queries, float radius, int width, int idx64,
                     const int32_t* order, void* out_idx, int32_t* out_counts, int32_t* out_stats, void* stream);
/* Re-pack [n_rows, in_width] int32 rows to [n_rows, out_width] (out_width <= in_width), int32 or int64. */
KPREG_API int kpreg_pack_rows(const int32_t* in, int64_t n_rows, int in_width, int out_width, int idx64,
                    void* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * KPConv (rigid) forward / backward.
 * Replaces KPConv.forward()   models/backbone_kpconv/finegrained_kpconv_blocks.py:265-401
 * (deformable=False branch; influence 'constant'|'linear'|'gaussian'; aggregation 'sum'|'closest';
 * including this fork's division by the number of neighbours with a positive feature sum, :396-399).
 *
 *   q_pts [n_q,3], s_pts [n_s,3] f32; idx [n_q,H] int32/int64 (pad value >= n_s = shadow);
 *   x [n_s,c_in] f32; weights [K,c_in,c_out] f32; kernel_points [K,3] f32; out [n_q,c_out] f32.
 *   influence: 0 constant, 1 linear, 2 gaussian.   aggregation: 0 sum, 1 closest.
 *   gemm: 0 = fp32 CUDA-core contraction, 1 = tcgen05 (3xTF32 split, fp32-accurate) contraction.
 * Backward returns d_x [n_s,c_in] and d_weights [K,c_in,c_out] (both overwritten).
 * ------------------------------------------------------------------------------------------- */
KPREG_API int kpreg_kpconv_workspace_bytes(int64_t n_q, int64_t n_s, int n_kpts, int c_in, int c_out,
                                 int backward, size_t* bytes);
KPREG_API int kpreg_kpconv_forward(const float* q_pts, const float* s_pts, const void* idx, int idx64,
                         const float* x, const float* weights, const float* kernel_points,
                         int64_t n_q, int64_t n_s, int n_nbrs, int n_kpts, int c_in, int c_out,
                         float kp_extent, int influence, int aggregation, int gemm, const int32_t* order,
                         float* out, void* workspace, size_t workspace_bytes, void* stream);
/* The same with the normalisation's row predicate supplied by the caller: row_pos [n_s] bytes, row_pos[j] != 0 iff the feature
 * sum of support row j is positive (finegrained_kpconv_blocks.py:396-397) — kpreg_segment_norm_forward_rowpos writes it
 * while it writes x, which saves this call's own pass over x.  row_pos == NULL: identical to kpreg_kpconv_forward. */
KPREG_API int kpreg_kpconv_forward_rowpos(const float* q_pts, const float* s_pts, const void* idx, int idx64,
                         const float* x, const float* weights, const float* kernel_points,
                         int64_t n_q, int64_t n_s, int n_nbrs, int n_kpts, int c_in, int c_out,
                         float kp_extent, int influence, int aggregation, int gemm, const int32_t* order,
                         const unsigned char* row_pos, float* out, void* workspace, size_t workspace_bytes, void* stream);
KPREG_API int kpreg_kpconv_backward(const float* q_pts, const float* s_pts, const void* idx, int idx64,
                          const float* x, const float* weights, const float* kernel_points,
                          const float* grad_out, int64_t n_q, int64_t n_s, int n_nbrs, int n_kpts,
                          int c_in, int c_out, float kp_extent, int influence, int aggregation,
                          const int32_t* order, float* d_x, float* d_weights, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Strided-block shortcut pooling.
 * Replaces max_pool()         models/backbone_kpconv/finegrained_kpconv_blocks.py:125-141
 * (max over the gathered rows of [x; 0]: a shadow index contributes the zero row).
 *   argmax [n_q,c] int32 receives the winning support row (n_s for the shadow row); may be NULL.
 * ------------------------------------------------------------------------------------------- */
KPREG_API int kpreg_max_pool_forward(const float* x, const void* idx, int idx64, int64_t n_q, int64_t n_s,
                           int n_nbrs, int channels, const int32_t* order, float* out, int32_t* argmax,
                           void* stream);
KPREG_API int kpreg_max_pool_backward(const float* grad_out, const int32_t* argmax, int64_t n_q, int64_t n_s,
                            int channels, float* d_x, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weighted Kabsch rigid-transform estimation for many correspondence sets.
 * Replaces compute_rigid_transform()       utils/se3_torch.py:131-173
 * and      fast_compute_rigid_transform()  utils/se3_torch.py:226-274 (weights <= threshold are
 * zeroed first, and — like the reference, :240-242 — written back when write_back != 0).
 *   a, b [total,3] f32; w [total] f32 or NULL (unweighted mean); set s covers rows
 *   offsets[s] .. offsets[s+1] (offsets int64 [n_sets+1]); if offsets is NULL every set has
 *   pts_per_set rows.  threshold < 0 disables thresholding.  out [n_sets,3,4] f32 = [R | t].
 * ------------------------------------------------------------------------------------------- */
KPREG_API int kpreg_kabsch(const float* a, const float* b, float* w, const int64_t* offsets, int64_t n_sets,
                 int64_t pts_per_set, float threshold, int write_back, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Encoder-block glue (the step around KPConv inside every block).
 *
 * kpreg_linear_forward: out[m, n] = act( (sum_k x[m,k] * weight[n,k]) * col_scale[n] + col_shift[n] + residual[m,n] )
 *   Replaces nn.Linear(bias=False) of UnaryBlock.mlp (models/backbone_kpconv/finegrained_kpconv_blocks.py:521-555)
 *   and of my_Bottle2neck (models/backbone_kpconv/res2net.py:84-159), with an eval-mode BatchNorm1d folded into
 *   col_scale / col_shift and the following ReLU in `act` (0 none, 1 relu, 2 leaky relu with `slope`).
 *   x [M, ldx], weight [N, K] (nn.Linear layout), out [M, ldc]; col_scale / col_shift / residual may be NULL.
 *   out2 (optional) [M, ld2] receives out + addend[M, ld_add] — the input of the next layer of res2net's chain.
 *   post_residual (optional) [M, ld_post] is added AFTER `act` and followed by `post_act`:
 *   out = post_act(act(...) + post_residual) — ResnetBottleneckBlock's `leaky_relu(res2net(x) + shortcut)`
 *   (finegrained_kpconv_blocks.py:723-725) riding on res2net's last GEMM.
 *   gemm: 1 = tcgen05 3xTF32 (falls back when the shape is not TMA-addressable), 0 = fp32 CUDA cores.
 *
 * kpreg_segment_norm_forward: out = act( (x - mean[c]) * rstd[c] + residual ), statistics per cloud c and channel
 *   Replaces BatchNormBlock's per-cloud nn.InstanceNorm1d (finegrained_kpconv_blocks.py:462-518; biased variance,
 *   eps inside the square root, no affine, no running statistics), optionally fused with the activation and the
 *   shortcut addition that follow it in the blocks.  channels, ldx, ldo, ld_res must be multiples of 4.
 *
 * kpreg_split_weights: for inference, split a weight matrix ONCE into the hi/lo TF32 operand pair of the tensor-core
 *   GEMM ([n_dim,k_dim] with transpose = 0, or KPConv's [K*c_in, c_out] with transpose = 1; `out` holds
 *   kpreg_linear_workspace_bytes(k_dim, n_dim) bytes).  Passing that buffer as `weight` / `weights` together with
 *   gemm = 2 to kpreg_linear_forward / kpreg_kpconv_forward skips the per-call split.  kpreg_gemm_supported tells
 *   whether a shape is addressable by the TMA path (K >= 4, row pitch and base 16-byte aligned, N >= 8).
 */
KPREG_API int kpreg_split_weights(const float* weight, int k_dim, int n_dim, int transpose, void* out, size_t out_bytes,
                                  void* stream);
KPREG_API int kpreg_gemm_supported(int64_t m_rows, int k_dim, int n_dim, int ldx, const void* x);
KPREG_API int kpreg_linear_workspace_bytes(int k_dim, int n_dim, size_t* bytes);
/* Output columns per CTA tile of the tensor-core GEMM for a [*, k_dim] x [k_dim, n_dim] product (32 / 64 / 128 / 256).  When
 * n_dim <= this width, a tile's rows are stored only after its whole reduction has been read and no other tile reads them:
 * `out` may then overlap columns of `x` (res2net's chained layers overwrite conv1's groups in place). */
KPREG_API int kpreg_linear_tile_cols(int k_dim, int n_dim);
KPREG_API int kpreg_linear_forward(const float* x, int ldx, const float* weight, int64_t m_rows, int k_dim, int n_dim,
                                   const float* col_scale, const float* col_shift, const float* residual, int ld_res,
                                   int act, float slope, float* out, int ldc, float* out2, int ld2, const float* addend,
                                   int ld_add, const float* post_residual, int ld_post, int post_act, int gemm,
                                   void* workspace, size_t workspace_bytes, void* stream);
/* act(([x1 | x2] Wcat^T) * col_scale + col_shift) without materialising the concatenation: the reduction index of ONE
 * tensor-core GEMM runs over x1 [M, k1] (row pitch ld1) and then over x2 [M, k2] (row pitch ld2).  w_split is the pre-split
 * weight pair (kpreg_split_weights) of Wcat [n_dim, pad32(k1) + k2] = [W1 | zero columns up to a multiple of 32 | W2].
 * With W1 = W2 = W this is (x1 + x2) W^T — one step of my_Bottle2neck's chain (models/backbone_kpconv/res2net.py:141-147)
 * reading the previous group's output and the next group of conv1's output where they lie; with W1 = conv3, W2 = the residual
 * projection, x1 = the concatenated group outputs and x2 = the unit's input it is conv3 + downsample of the same unit
 * (res2net.py:153-159) without a copy of the input behind the concatenation.  post_residual (may be NULL; row pitch ld_post) and
 * post_act as in kpreg_linear_forward: out = post_act(act(...) + post_residual).  Both bases 16-byte aligned, ld1 and
 * ld2 multiples of 4, k1 >= 4, n_dim >= 8. */
KPREG_API int kpreg_linear_pair_forward(const float* x1, int ld1, int k1, const float* x2, int ld2, int k2, const float* w_split,
                                        int64_t m_rows, int n_dim, const float* col_scale, const float* col_shift, int act,
                                        float slope, const float* post_residual, int ld_post, int post_act, float* out, int ldc,
                                        void* stream);

/* Backward of y = x W^T on the tensor cores (3xTF32): dx [M, K] = dy W (may be NULL) and d_weight [N, K] = dy^T x (may be
 * NULL; a split-k reduction over the M rows, accumulated with fp32 atomics).  Returns KPREG_E_INVALID when a shape is not
 * addressable by the TMA paths (K or N < 8, pitches / bases not 16-byte aligned): the caller then uses its own kernels. */
KPREG_API int kpreg_linear_backward_workspace_bytes(int64_t m_rows, int k_dim, int n_dim, size_t* bytes);
KPREG_API int kpreg_linear_backward(const float* x, int ldx, const float* dy, int ld_dy, const float* weight, int64_t m_rows,
                                    int k_dim, int n_dim, float* dx, int ld_dx, float* d_weight, void* workspace,
                                    size_t workspace_bytes, void* stream);
KPREG_API int kpreg_segment_norm_workspace_bytes(int n_clouds, int channels, size_t* bytes);
KPREG_API int kpreg_segment_norm_forward(const float* x, int ldx, const int32_t* lens, int n_clouds, int64_t n_rows,
                                         int channels, float eps, const float* residual, int ld_res, int act, float slope,
                                         float* out, int ldo, void* workspace, size_t workspace_bytes, void* stream);
/* The same, also writing KPConv's row predicate of the output (row_pos [n_rows] bytes: 1 iff the fp64 sum of out's row is
 * positive; NULL = none) for the KPConv that consumes `out` (kpreg_kpconv_forward_rowpos).  Only where one thread group holds
 * a whole row: kpreg_segment_norm_rowpos_supported(channels) != 0 (32 or 64 channels), KPREG_E_INVALID otherwise. */
KPREG_API int kpreg_segment_norm_rowpos_supported(int channels);
KPREG_API int kpreg_segment_norm_forward_rowpos(const float* x, int ldx, const int32_t* lens, int n_clouds, int64_t n_rows,
                                                int channels, float eps, const float* residual, int ld_res, int act, float slope,
                                                float* out, int ldo, unsigned char* row_pos, void* workspace,
                                                size_t workspace_bytes, void* stream);
/* Backward of the plain norm (no residual / activation): dx = rstd * (dy - mean(dy) - xhat * mean(dy * xhat)) per cloud and
 * channel, statistics recomputed from x.  Same workspace size as the forward. */
KPREG_API int kpreg_segment_norm_backward(const float* x, int ldx, const float* dy, int ld_dy, const int32_t* lens,
                                          int n_clouds, int64_t n_rows, int channels, float eps, float* dx, int ld_dx,
                                          void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The hierarchical chain of my_Bottle2neck in one kernel (models/backbone_kpconv/res2net.py:137-150):
 *     sp = t_0;  for i in 0 .. n_layers-1:  sp = relu(sp W_i^T + shift_i);  z_i = sp;  sp = sp + t_{i+1}
 * with t_g = t[:, g*width : (g+1)*width] (conv1's output, n_layers + 1 groups), W_i / shift_i the i-th
 * convs[i] / bns[i] pair with the eval-mode BatchNorm1d folded in.  z[:, i*width : (i+1)*width] = z_i for
 * i < n_layers, z[:, n_layers*width : (n_layers+1)*width] = t_{n_layers} (the group the reference passes through
 * unchanged, :151-152), and — when x_copy is given — z[:, (n_layers+1)*width : +c_x] = x_copy (the block input,
 * so that conv3 and the residual projection run as one GEMM over the concatenation).
 * A warp keeps 32 rows of sp in registers across all layers (mma.sync 3xTF32, weights resident in shared memory):
 * t is read once and z written once.
 *   kpreg_chain_supported: 1 if (width, n_layers) is served by this kernel (even width, ceil(width/8) in
 *       {2,4,7,8}, packed weights <= 200 KiB), else 0 — the caller then runs the layers through kpreg_linear_forward.
 *   kpreg_chain_pack: arrange weights [n_layers, width, width] (nn.Linear layout [out, in], BN scale folded in)
 *       and shifts [n_layers, width] once into the fragment order the kernel reads (kpreg_chain_pack_bytes bytes).
 *   kpreg_chain_forward: t [M, ld_t], z [M, ld_z] (ld_t, ld_z even, 8-byte aligned bases), x_copy [M, ld_x] or NULL.
 * ------------------------------------------------------------------------------------------- */
KPREG_API int kpreg_chain_supported(int width, int n_layers);
KPREG_API int kpreg_chain_pack_bytes(int width, int n_layers, size_t* bytes);
KPREG_API int kpreg_chain_pack(const float* weights, const float* shifts, int width, int n_layers, void* pack,
                               size_t pack_bytes, void* stream);
KPREG_API int kpreg_chain_forward(const float* t, int ld_t, const void* pack, int width, int n_layers, int64_t m_rows,
                                  float* z, int ld_z, const float* x_copy, int ld_x, int c_x, void* stream);

/* ---------------------------------------------------------------------------------------------
 * conv1 + bn1 + relu AND the hierarchical chain of my_Bottle2neck in one tcgen05 kernel
 * (models/backbone_kpconv/res2net.py:125-152): per 128-row tile, group by group,
 *     t_g = relu(x W1_g^T + b1_g),   y_0 = relu(t_0 Wc_0^T + bc_0),   y_g = relu((y_{g-1} + t_g) Wc_g^T + bc_g)
 * and z = [y_0 | ... | y_{G-2} | t_{G-1} | x (when copy_x)].  conv1's output and the running activation stay on the SM
 * (TMEM accumulators, fp16 hi/lo operand boxes in shared memory): x is read once, z written once.
 *   kpreg_front_supported: 1 if (width, n_groups, c_in) is served: width % 4 == 0, 2 <= n_groups <= 8, c_in % 4 == 0 and
 *       either 16 <= width <= 32 with c_in <= 32 (all weights resident in shared memory, two CTAs per SM) or
 *       32 < width <= 64 with c_in <= 64 (weights streamed per group from L2).  Else 0: the caller runs conv1 through
 *       kpreg_linear_forward and the chain through kpreg_chain_forward.
 *   kpreg_front_pack: w1 [n_groups*width, c_in] and wc [n_groups-1, width, width] (nn.Linear layout [out, in], BN scale folded
 *       in), b1 [n_groups*width], bc [n_groups-1, width] -> the operand boxes the kernel copies (kpreg_front_pack_bytes bytes).
 *   kpreg_front_forward: x [M, ld_x] (c_in columns), z [M, ld_z]; ld_x, ld_z multiples of 4, bases 16-byte aligned,
 *       ld_z >= n_groups*width (+ c_in when copy_x).  Operands are split into fp16 pairs: |value| < 65504.
 * ------------------------------------------------------------------------------------------- */
KPREG_API int kpreg_front_supported(int width, int n_groups, int c_in);
KPREG_API int kpreg_front_pack_bytes(int width, int n_groups, int c_in, size_t* bytes);
KPREG_API int kpreg_front_pack(const float* w1, const float* b1, const float* wc, const float* bc, int width, int n_groups,
                               int c_in, void* pack, size_t pack_bytes, void* stream);
KPREG_API int kpreg_front_forward(const float* x, int ld_x, int c_in, const void* pack, int width, int n_groups,
                                  int64_t m_rows, float* z, int ld_z, int copy_x, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The steps on either side of the path.
 *
 * kpreg_overlap_pool: one level of compute_overlaps()   models/backbone_kpconv/finegrained_kpconv.py:545-571
 *   out[n] = clamp(mean of level[idx[n,h]] over the entries with idx[n,h] < n_s, 0, 1); a row without a valid entry
 *   is 0/0 = NaN, as in the reference.  level [n_s] f32, idx [n_q, n_nbrs] (the pyramid's `pools` table), out [n_q].
 *
 * kpreg_sine_embed: PositionEmbeddingCoordsSine.forward()   models/transformer/position_embedding.py:29-49
 *   out[r, d*F + k] = (k even ? sin : cos)(xyz[r,d] * scale / dim_t[k]), columns >= n_dim*F zero.  dim_t [F] is the
 *   reference's `temperature ** (2 * (k // 2) / F)` evaluated by the host in fp32 (same torch expression).
 *
 * kpreg_pack_coarse: split_src_tgt() + pad_sequence(..., require_padding_mask=True) of the projected coarse features AND
 *   of their position embedding   utils/seq_manipulation.py:6-48, models/finegrained_regtr.py:149-172
 *   feats [N, ld_f] (N = sum of lens; the first n_pairs clouds are sources, the last n_pairs targets), xyz [N,3],
 *   lens int32 [2*n_pairs] on the device.  src_feats / src_pe [ns_max, n_pairs, d_model], tgt_* [nt_max, n_pairs,
 *   d_model] (zero padded), src_mask [n_pairs, ns_max] / tgt_mask [n_pairs, nt_max] bytes, 1 at padded positions.
 *   Any output pointer may be NULL.  workspace: kpreg_pack_coarse_workspace_bytes(n_pairs).
 *
 * kpreg_shuffle_gather / kpreg_remap_pairs: ShufflePoints.__call__()   data_loaders/transforms.py:95-131
 *   out_pts[i] = pts[perm[i]], out_mask[i] = mask[perm[i]] (mask may be NULL), rev[perm[i]] = i and -1 elsewhere
 *   (rev [n_in] int64, may be NULL); perm int64 [n_out] — the host's permutation, truncated to max_pts.  status
 *   (int32 on the device) receives KPREG_E_RANGE if perm holds an index outside [0, n_in).
 *   kpreg_remap_pairs maps correspondences corr [2, n_pairs] int64 through the two reverse indices into out [2,
 *   n_pairs] and sets keep[p] = 1 where both survive; the caller compacts the kept columns in order.
 * ------------------------------------------------------------------------------------------- */
KPREG_API int kpreg_overlap_pool(const float* level, const void* idx, int idx64, int64_t n_q, int64_t n_s, int n_nbrs,
                                 float* out, void* stream);
KPREG_API int kpreg_sine_embed(const float* xyz, int64_t n_rows, int n_dim, int d_model, int num_feats, float scale,
                               const float* dim_t, float* out, void* stream);
KPREG_API int kpreg_pack_coarse_workspace_bytes(int n_pairs, size_t* bytes);
KPREG_API int kpreg_pack_coarse(const float* feats, int ld_f, const float* xyz, const int32_t* lens, int n_pairs, int d_model,
                                int num_feats, float scale, const float* dim_t, int ns_max, int nt_max, float* src_feats,
                                float* tgt_feats, float* src_pe, float* tgt_pe, unsigned char* src_mask,
                                unsigned char* tgt_mask, void* workspace, size_t workspace_bytes, void* stream);
KPREG_API int kpreg_shuffle_gather(const float* pts, const unsigned char* mask, const int64_t* perm, int64_t n_out,
                                   int64_t n_in, float* out_pts, unsigned char* out_mask, int64_t* rev, int32_t* status,
                                   void* stream);
KPREG_API int kpreg_remap_pairs(const int64_t* corr, int64_t n_pairs, const int64_t* rev_src, int64_t n_src,
                                const int64_t* rev_tgt, int64_t n_tgt, int64_t* out, unsigned char* keep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KPREG_B200_H */
