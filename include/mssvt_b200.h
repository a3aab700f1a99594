/*
 * mssvt_b200.h -- C-ABI of libmssvt_b200.so: the B200 (sm_100a) implementation of the MsSVT
 * mixed-scale sparse voxel attention backbone hot path.
 *
 * Drop-in boundary.  The reference binds its native layer with pybind11 (`mssvt_ops_cuda`,
 * pcdet/ops/mssvt/src/ms_api.cpp:7-14, and four functions of `pointnet2_batch_cuda`,
 * pcdet/ops/pointnet2/pointnet2_batch/src/pointnet2_api.cpp:10-24), passing at::Tensor.  Every
 * function below replaces one of those entry points (cited per function) with:
 *   - plain device pointers and sizes, no torch types;
 *   - an explicit stream (`void *stream` = cudaStream_t; the reference always launches on the
 *     legacy default stream);
 *   - an int status instead of fprintf + exit(-1) (ms_sparse_attention_gpu.cu:110-114):
 *       0 ok, -1 invalid argument, -2 CUDA launch/runtime error (mssvt_last_cuda_error()),
 *       -3 workspace too small;
 *   - caller-allocated outputs (as in the reference, mssvt_ops.py:16-17, 36-41, 77-85), but
 *     the -1 / 0 pre-fill is done on the device by the callee, so outputs may be uninitialised.
 * All functions are asynchronous with respect to the host and re-entrant; none of them
 * synchronises, allocates device memory or keeps global state.
 *
 * Layout conventions (same as the reference): voxel / window indices are int32 rows
 * [batch, z, y, x]; hash tables are (B, H, 2) int32 [key, value] with empty = -1,
 * h(k) = k % H and linear probing; features are row-major fp32 (N, C).
 */
#ifndef MSSVT_B200_H
#define MSSVT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MSSVT_OK 0
#define MSSVT_ERR_INVALID (-1)
#define MSSVT_ERR_LAUNCH (-2)
#define MSSVT_ERR_WORKSPACE (-3)

const char *mssvt_version(void);
int mssvt_last_cuda_error(void);
long long mssvt_launch_count(void); /* kernels launched by this library since it was loaded */

/* ---- utilities ------------------------------------------------------------------------- */

int mssvt_fill_i32(int *p, long long count, int value, void *stream);

/* per-sample row counts and their exclusive prefix, from the batch column of (n, 4) indices.
 * Replaces the `.sum().item()` loops of with_bs_cnt (mssvt_backbone.py:124-130) and
 * SparseTensor.build_map_table (mssvt_utils.py:35-38).  counts (B), start (B + 1). */
int mssvt_count_samples(int num_rows, int batch_size, const int *indices, int *counts, int *start,
                        void *stream);

/* with_coords (mssvt_backbone.py:132-137): xyz (N, 3) = (idx[x,y,z] + 0.5) * voxel_size + min,
 * three separately rounded fp32 operations.  voxel_size, range_min: 3 HOST floats each. */
int mssvt_voxel_world_coords(int num_voxels, const int *v_indices, const float *voxel_size,
                             const float *range_min, float *xyz, void *stream);

/* dst[i] = sum_{j<i} src[j * stride] for i in [0, n], n = min(n_cap, *n_dev) read on the device
 * (n_dev may be NULL).  dst: n_cap + 1 ints; workspace: ceil((n_cap + 1) / 1024) + 1 ints. */
int mssvt_exclusive_scan(int n_cap, const int *n_dev, const int *src, int stride, int *dst, int *workspace,
                         void *stream);

/* ---- mssvt_ops_cuda replacements ---------------------------------------------------------- */

/* build_mapping_with_hash_wrapper (ms_sparse_attention.cpp:23-35; kernel ..._gpu.cu:66-115).
 * table: (batch_size, hash_size, 2), filled by this call. */
int mssvt_build_hash_table(int x_max, int y_max, int z_max, int num_voxels, int hash_size,
                           int batch_size, const int *v_indices, const int *v_bs_cnt, int *table,
                           void *stream);

/* Table content view (no reference counterpart; used by tests): values[i] = value stored for
 * keys[i] in sample batch_ids[i], -1 if absent (hash_table_find, ..._gpu.cu:43-64). */
int mssvt_hash_lookup(int hash_size, int num_queries, const int *batch_ids, const int *keys,
                      const int *table, int *values, void *stream);

/* window_with_hash_wrapper (ms_sparse_attention.cpp:37-59; kernel ..._gpu.cu:117-191) plus the
 * Python mask/cat loop of WindowPartition.forward (mssvt_ops.py:45-53).
 *   win_list  (list_capacity, 4) rows [b, wz, wy, wx] of all samples concatenated, numbered by
 *             first occurrence in voxel order (the reference numbers by atomicAdd arrival);
 *   table     (batch_size, hash_size, 2) window key -> per-sample row id, filled by this call;
 *   win_count (batch_size + 2): per-sample window counts (before any clamping), [B] = rows in
 *             win_list = kept windows (a sample keeps its first max_wins windows, kept windows of
 *             all samples are contiguous, at most list_capacity rows), [B+1] = windows dropped
 *             (the reference writes out of bounds in that case);
 *   workspace: mssvt_window_partition_workspace_bytes(num_voxels) bytes of scratch. */
long long mssvt_window_partition_workspace_bytes(int num_voxels);
int mssvt_window_partition(int x_wgs, int y_wgs, int z_wgs, int x_ws, int y_ws, int z_ws,
                           int num_voxels, int max_wins, int hash_size, int batch_size,
                           int list_capacity, const int *v_indices, int *win_list, int *table,
                           int *win_count, void *workspace, long long workspace_bytes,
                           void *stream);

/* The window LIST of mssvt_window_partition without the window hash table: what the fused path needs (it finds
 * voxels through the grid index, mssvt_grid_index_build).  Numbering through a dense first-voxel array over the
 * window grid: same rows in the same first-occurrence order, same win_count layout.  The reference-contract
 * table is produced by mssvt_window_partition (operator API) or, from a list, by mssvt_build_hash_table. */
long long mssvt_window_list_workspace_bytes(int x_wgs, int y_wgs, int z_wgs, int batch_size, int num_voxels);
int mssvt_window_list(int x_wgs, int y_wgs, int z_wgs, int x_ws, int y_ws, int z_ws, int num_voxels, int max_wins,
                      int batch_size, int list_capacity, const int *v_indices, int *win_list, int *win_count,
                      void *workspace, long long workspace_bytes, void *stream);

/* gather_two_window_voxels_with_hash_wrapper (ms_sparse_attention.cpp:61-120; kernel
 * ..._gpu.cu:193-381).  Same argument order as the reference wrapper; ind_* (W, max_*) padded
 * with -1, coord_* (W, max_*, 3) padded with 0, all written in full by this call. */
int mssvt_gather_two_window(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                            int max_odd, int max_even, int max_win1, int max_win2, int num_wins,
                            int hash_size, int num_odd, int num_even, int num_win1, int num_win2,
                            int *ind_odd, int *ind_even, int *ind_win1, int *ind_win2,
                            int *coord_odd, int *coord_even, int *coord_win1, int *coord_win2,
                            const int *q_odd, const int *q_even, const int *q_win1,
                            const int *q_win2, const int *win_indices, const int *table,
                            void *stream);

/* gather_one_window_voxels_with_hash_wrapper (ms_sparse_attention.cpp:122-149; kernel
 * ..._gpu.cu:383-458). */
int mssvt_gather_one_window(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                            int max_win1, int num_wins, int hash_size, int num_win1, int *ind_win1,
                            int *coord_win1, const int *q_win1, const int *win_indices,
                            const int *table, void *stream);

/* group_features_wrapper / group_features_grad_wrapper (group_features.cpp:29-68; kernels
 * group_features_gpu.cu:73-129, 15-70).  out (M, C, nsample) is written in full (zeros where
 * idx < 0); grad_features (N, C) is zeroed then accumulated. */
int mssvt_group_features(int B, int M, int C, int nsample, const float *features,
                         const int *features_batch_cnt, const int *idx, const int *idx_batch_cnt,
                         float *out, void *stream);
int mssvt_group_features_grad(int B, int M, int C, int N, int nsample, const float *grad_out,
                              const int *idx, const int *idx_batch_cnt,
                              const int *features_batch_cnt, float *grad_features, void *stream);

/* ---- pointnet2_batch_cuda replacements (the four ops the backbone calls) -------------------- */

/* farthest_point_sampling_wrapper (sampling.cpp:41-50; kernel sampling_gpu.cu:100-260), including
 * the reference's tie order.  temp (b, n) scratch is only needed for n > ~50 000 (may be NULL). */
int mssvt_fps(int b, int n, int m, const float *dataset, float *temp, int *idxs, void *stream);
int mssvt_fps_log2_block(int n); /* log2 of the reference's FPS block size, cuda_utils.h:10-14 */

/* gather_points_wrapper (sampling.cpp:13-24; kernel sampling_gpu.cu:15-51) */
int mssvt_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                        float *out, void *stream);

/* three_nn_wrapper (interpolate.cpp:17-29; kernel interpolate_gpu.cu:16-81): squared distances */
int mssvt_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                   int *idx, void *stream);

/* group_points_wrapper / group_points_grad_wrapper (group_points.cpp:18-44; kernels
 * group_points_gpu.cu:53-92, 14-50) */
int mssvt_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                       const int *idx, float *out, void *stream);
int mssvt_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                            const int *idx, float *grad_points, void *stream);

/* ---- fused entry points used by the backbone module ------------------------------------------ */

/* Voxel lookup structure of the fused path (replaces the hash table there; the op-level API above
 * keeps the reference's table): an occupancy bitmap with rank.  cells: (B * X * Y * ceil(Z/32), 2)
 * int32 {bits, base}; vals: (N) per-sample voxel index at base + popcount(bits below z).  A lookup
 * is one 8-byte load, plus one 4-byte load on a hit; no probing, no dependence on a hash size.
 * workspace: ceil(mssvt_grid_index_words / 1024) + 1 int32. */
long long mssvt_grid_index_words(int x_max, int y_max, int z_max, int batch_size);
int mssvt_grid_index_build(int x_max, int y_max, int z_max, int num_voxels, int batch_size,
                           const int *v_indices, const int *v_start, int *cells, int *vals,
                           int *workspace, void *stream);

/* Coordinate-only part of MixedScaleSparseTransformerBlock.forward (mssvt_backbone.py:213-258,
 * 264-269, 300-307) in one kernel, one warp per window, with no host synchronisation: the
 * number of windows is read from device memory (win_count_total) and bounds the work.
 * Outputs per window w < min(win_capacity, *win_count_total) (rows beyond are untouched):
 *   q_row    (cap, nq)        global feature row of each query slot (-1 pad); nq = |odd| / |even|
 *                             / max_win1 for cbs_pattern 1 / 0 / 2
 *   win1_row (cap, max_win1)  global row of each win1 voxel (-1 pad)
 *   k_row    (cap, 2K)        global row of each key slot: K from the win1 list, K from win2
 *   k_mask   (cap, 2K)        1 where the reference masks the key
 *   nn_idx   (cap, max_win1, 3) uint8, nn_w (cap, max_win1, 3): three_nn picks among the query
 *                             slots and normalised inverse-distance weights (use_interp only)
 *   covered  (num_voxels)     1 for every row the merge will overwrite (zeroed by this call)
 *   fps_idx_tap (cap, 2K), counts_tap (cap, 4): optional raw FPS picks / list lengths (NULL ok)
 *   rep_row (cap, 2K), meta (cap, 4): optional (both or neither) compact form for the tensor-core
 *                             window kernel: per scale the rows of the DISTINCT keys (unmasked slots
 *                             in order, then one entry standing for all masked slots) and
 *                             {#real queries, #win1 voxels, nrep0 | nmask0 << 8, nrep1 | nmask1 << 8}
 *   vox_slot (num_voxels)     optional: w * max_win1 + i of the win1 slot holding each voxel, -1 if none
 * voxel_size, range_min: 3 HOST floats each. */
int mssvt_block_geometry(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                         int num_odd, int num_even, int num_win1, int num_win2,
                         int max_win1, int max_win2, int key_num_sample, int cbs_pattern,
                         int use_interp, const float *voxel_size, const float *range_min,
                         const int *q_odd, const int *q_even, const int *q_win1, const int *q_win2,
                         int win_capacity, const int *win_count_total, const int *win_list,
                         const int *grid_cells, const int *grid_vals, const int *v_start,
                         int num_voxels, int *q_row,
                         int *win1_row, int *k_row, unsigned char *k_mask, unsigned char *nn_idx,
                         float *nn_w, unsigned char *covered, int *fps_idx_tap, int *counts_tap,
                         int *rep_row, int *meta, int *vox_slot, int *odd_row, int *even_row, void *stream);

/* The pattern-dependent part of the block geometry (query rows, #real queries, three-NN + weights;
 * mssvt_backbone.py:220-234, 300-307) for ANOTHER cbs_pattern, from the outputs of one mssvt_block_geometry call
 * over the same windows: src_row = the list that serves as the query set -- odd_row (cap, |odd|) for pattern 1,
 * even_row (cap, |even|) for pattern 0 (both optional outputs of mssvt_block_geometry: global rows, -1 padded) or
 * win1_row for pattern 2 -- with nq entries per window; meta_in = that call's meta; xyz = mssvt_voxel_world_coords.
 * Writes q_row (cap, nq), meta_out (cap, 4) (entry 0 = #real queries of THIS pattern, the rest copied) and, with
 * use_interp, nn_idx / nn_w (cap, max_win1, 3).  Bit-identical to a direct call with that pattern; the chessboard
 * probes and both FPS passes are not repeated (blocks that differ only in cbs_pattern share them). */
int mssvt_block_queries(int nq, int max_win1, int use_interp, int win_capacity, const int *win_count_total,
                        const int *src_row, const int *win1_row, const int *meta_in, const float *xyz, int *q_row,
                        int *meta_out, unsigned char *nn_idx, float *nn_w, void *stream);

/* q_src[q_base[w] + s] = w * nq + s for every real query slot: inverse of the compact query numbering
 * (q_base = mssvt_exclusive_scan over meta[:, 0]). */
int mssvt_query_src(int win_capacity, const int *win_count_total, int nq, const int *meta, const int *q_base,
                    int *q_src, void *stream);

/* One-window gather of the compress block without host synchronisation (same table walk as
 * mssvt_gather_one_window): k_row (cap, max_win1) global feature rows, -1 padded. */
int mssvt_window_rows(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                      int num_win1, int max_win1, const int *q_win1, int win_capacity,
                      const int *win_count_total, const int *win_list, const int *grid_cells,
                      const int *grid_vals, const int *v_start, int *k_row, void *stream);

/* nn.LayerNorm(C) over rows (mssvt_backbone.py:210, 352); num_rows_dev (may be NULL) bounds the
 * row count from device memory. */
int mssvt_layernorm(int num_rows, const int *num_rows_dev, int C, const float *x,
                    const float *gamma, const float *beta, float eps, float *y, void *stream);

/* Feature part of a two-window block between norm1 and the residual: gather + pos_proj +
 * MixedScaleAttention + interpolation + merge (mssvt_backbone.py:260-336; mssvt_utils.py:88-157).
 * `shape` is the AttnShape descriptor (mssvt_b200/csrc/block.cu; mirrored in mssvt_b200/_lib.py),
 * `params` the transposed fp32 weight pack it indexes.  Writes merged[row] for covered rows. */
int mssvt_block_attention(const void *shape, int shape_bytes, const float *params, int win_capacity,
                          const int *win_count_total, const int *win_list, const float *xn,
                          const float *xyz, const int *q_row, const int *k_row,
                          const unsigned char *k_mask, const int *win1_row,
                          const unsigned char *nn_idx, const float *nn_w, float *merged,
                          void *stream);

/* Tile plan of mssvt_block_attention_tc: packs consecutive windows of one scale into tiles of <= 128
 * distinct keys.  A function of the geometry only (meta, q_base, win_list): made once per frame and shared
 * by every block that runs over the same window lists.  Outputs (opaque, caller-allocated):
 * tiles (2, win_capacity, 2) int, tile_count (2) int, win_rec (2, win_capacity, 4) int,
 * win_ctr (win_capacity, 4) float. */
int mssvt_attention_tiles(int heads_per_group, int nq, int key_num_sample, int win_capacity,
                          const int *win_count_total, const int *win_list, const int *meta, const int *q_base,
                          const float *win_cell, const float *range_min, int *tiles, int *tile_count,
                          int *win_rec, float *win_ctr, unsigned char *tile_rows, void *stream);

/* The same step (mssvt_backbone.py:260-336 + mssvt_utils.py:88-157) as ONE tile kernel on the tcgen05 tensor
 * cores (mssvt_b200/csrc/attention_tc.cu): per tile of <= 128 rows (distinct keys + real queries of consecutive
 * windows of one head group) the positional embedding (Conv1d 6 -> C), the [K | V | Q] projection and the
 * output projection run as tcgen05.mma (TF32 operands, or split 3xTF32 with terms = 3; fp32 accumulate);
 * scores, softmax and AV of the tiny per-window matrices run on the FP32 pipe.
 * Weights packed by mssvt_pack_operand_tf32: wpos_packed = [pos_w | pos_b | 0] as a [64][8] matrix, ALWAYS
 * packed with terms = 3; wkvq0 / wkvq1 = per head group the [96][32] matrix [Wk; Wv; scale * Wq]; wp0 / wp1 = the
 * group's [32][32] output projection (packed with `terms`); bq / bkv / bp: the nn.Linear biases.
 * rep_row / meta from mssvt_block_geometry; q_base = mssvt_exclusive_scan(meta[:, 0]) (win_capacity + 1 ints),
 * q_src = mssvt_query_src, vox_slot from the geometry; tiles / tile_count / win_rec / win_ctr / tile_rows =
 * mssvt_attention_tiles; scratch: 3 * num_voxels * 64 floats (projected query rows in the last third).
 * Supported: C = 64, two groups of 32 channels with 1, 2 or 4 heads each, nq <= 32,
 * key_num_sample <= 63, cap1 <= 128; -1 otherwise. */
int mssvt_block_attention_tc(int C, int heads_per_group, int nq, int key_num_sample, int cap1, int interp,
                             int terms, float scale, const float *wpos_packed, const float *wkvq0,
                             const float *wkvq1, const float *wp0, const float *wp1, const float *bq0,
                             const float *bq1, const float *bkv0, const float *bkv1, const float *bp0,
                             const float *bp1, int win_capacity, const int *win_count_total,
                             const float *xn, const float *xyz, const int *q_row,
                             const int *rep_row, const int *meta, const int *q_base, const int *q_src,
                             const int *vox_slot, const unsigned char *nn_idx,
                             const float *nn_w, const int *tiles, const int *tile_count, const int *win_rec,
                             const float *win_ctr, const unsigned char *tile_rows, int num_voxels, float *scratch,
                             float *merged, void *stream);

/* Attention of a one-window (compress) block (mssvt_backbone.py:361-383): out (cap, C). */
int mssvt_compress_attention(const void *shape, int shape_bytes, const float *params,
                             int win_capacity, const int *win_count_total, const int *win_list,
                             const float *xn, const float *xyz, const int *k_row, float *out,
                             void *stream);

/* Tile plan of mssvt_compress_attention_tc: #real slots per window, tiles of <= 128 key tasks, window
 * centres.  A function of the window rows (coordinates) only, so it can run ahead of / concurrently with the
 * feature kernels.  Outputs (opaque, caller-allocated): tiles (win_capacity, 2) int, tile_count (1) int,
 * win_rec (win_capacity) int, win_ctr (win_capacity, 4) float. */
int mssvt_compress_tiles(int n1, int win_capacity, const int *win_count_total, const int *win_list,
                         const int *k_row, const float *win_cell, const float *range_min, int *tiles,
                         int *tile_count, int *win_rec, float *win_ctr, void *stream);

/* The same step, task-parallel with the second positional-embedding layer and the K/V projection on
 * the tcgen05 tensor cores (mssvt_b200/csrc/compress_tc.cu).  Weights: pos_w [64][6] (nn.Module layout);
 * packed by mssvt_pack_operand_tf32: wq / wp / pos2_w [64][64], wkv [128][64].  tiles .. win_ctr:
 * mssvt_compress_tiles.  scratch: 3 * win_capacity * 64 floats.
 * Supported: C = 64, one head group with 2, 4 or 8 heads, two-layer pos_proj, n1 <= 127; -1 otherwise. */
int mssvt_compress_attention_tc(int C, int heads, int n1, int terms, float scale, const float *win_cell,
                                const float *range_min, const float *pos_w, const float *pos_b,
                                const float *pos2_w, const float *pos2_b, const float *wq, const float *bq,
                                const float *wkv, const float *bkv, const float *wp, const float *bp,
                                int win_capacity, const int *win_count_total, const int *win_list,
                                const float *xn, const float *xyz, const int *k_row, const int *tiles,
                                const int *tile_count, const int *win_rec, const float *win_ctr, float *scratch,
                                float *out, void *stream);

/* residual + norm2 + linear1 / ReLU / linear2 + residual (+ out_linear)
 * (mssvt_backbone.py:337-343, 384-387).  `shape` is the FfnShape descriptor. */
int mssvt_ffn(const void *shape, int shape_bytes, const float *params, int num_rows,
              const int *num_rows_dev, const float *x, const float *merged,
              const unsigned char *covered, float *y, void *stream);

/* Packs a weight matrix w [n_rows][k] (nn.Linear layout, out x in) for the tensor-core entry points:
 * TF32 rounding + the K-major 8-row x 16-byte core-matrix layout tcgen05.mma reads from shared memory.
 * terms = 1: plain TF32, packed = n_rows * k floats.  terms = 3 ("3xTF32"): w = w_hi + w_lo with w_hi =
 * tf32(w), w_lo = tf32(w - w_hi), packed = [hi | lo] = 2 * n_rows * k floats; the *_tc entry points called with
 * terms = 3 split their activations the same way and issue A_hi W_hi + A_lo W_hi + A_hi W_lo per K step:
 * fp32-grade results (~1e-6 relative) from the TF32 tensor pipe.  n_rows % 8 == 0 and k % 8 == 0.  Done once per weight (the kernels then
 * stage it with plain asynchronous copies); arguments documented as "packed" below take this form. */
int mssvt_pack_operand_tf32(const float *w, int n_rows, int k, int terms, float *packed, void *stream);

/* The bf16 form (precision mode "bf16": tcgen05.mma.kind::f16 with bf16 operands, fp32 accumulate): packed holds
 * n_rows * k bf16 in the same K-major core-matrix layout (16-byte chunks of 8 elements).  n_rows % 8 == 0,
 * k % 16 == 0.  mssvt_ffn_tc and mssvt_block_attention_tc take these copies when called with terms = 0. */
int mssvt_pack_operand_bf16(const float *w, int n_rows, int k, void *packed, void *stream);

/* The split bf16 form (precision mode "bf16x3", terms = 2): w = w_hi + w_mid with w_hi = bf16(w), w_mid = bf16(w - w_hi);
 * packed = [hi | mid] = 2 * n_rows * k bf16 in the layout above.  The *_tc entry points called with terms = 2 split their
 * activations the same way and issue A_hi W_hi + A_mid W_hi + A_hi W_mid per K = 16 step on kind::f16: ~16 significant
 * bits per operand (results within 1e-4 of an fp32 reference, measured ~2e-5) from operand tiles of the size of one
 * TF32 tile. */
int mssvt_pack_operand_bf16x2(const float *w, int n_rows, int k, void *packed, void *stream);

/* Self-check of the tensor-map (TMA, cp.async.bulk.tensor) row movement mssvt_ffn_tc uses for its dense row tiles
 * (mssvt_b200/csrc/tma.cuh): copies a row-major (num_rows, 64) fp32 matrix src -> dst through box loads, swizzled
 * shared-memory boxes read and re-written by their owning lanes, and box stores; dst == src bit for bit, rows past
 * num_rows untouched.  (No reference counterpart: the reference moves rows with plain loads / stores.)
 * MSSVT_ERR_LAUNCH if the driver does not provide cuTensorMapEncodeTiled. */
int mssvt_tma_copy_rows(const float *src, float *dst, int num_rows, void *stream);

/* The same FFN on the tcgen05 tensor cores: TF32 operands, fp32 accumulation in TMEM, LayerNorm and
 * the residual stream in fp32 (mssvt_b200/csrc/ffn_tc.cu).  w1 [F][C] and w2 [C][F] (nn.Linear layout) packed by
 * mssvt_pack_operand_tf32.  Supported shapes: C in {32, 64}, F % 64 == 0, F + C <= 512; -1 otherwise.
 * mode 2 (C = 64): the three-NN interpolation + merge of mssvt_block_attention_tc happens on the way in --
 * `merged` is not read; vox_slot / meta / q_base / nn_idx / nn_w are the geometry maps, `projected` the projected
 * query rows (scratch + 2 * num_voxels * 64 of that call), cap1 = max_num_win1.  Otherwise those 7 are ignored.
 * xn_next (optional, with next_ln_g / next_ln_b / next_eps): also writes LayerNorm(y) with the NEXT block's
 * norm1 parameters, which saves that block its own LayerNorm pass over y. */
int mssvt_ffn_tc(int C, int F, int mode, int terms, float eps, const float *ln_g, const float *ln_b, const float *w1_packed,
                 const float *b1, const float *w2_packed, const float *b2, int num_rows, const int *num_rows_dev,
                 const float *x, const float *merged, const unsigned char *covered, float *y,
                 const float *next_ln_g, const float *next_ln_b, float next_eps, float *xn_next,
                 const int *vox_slot, const int *meta, const int *q_base, const unsigned char *nn_idx,
                 const float *nn_w, const float *projected, int cap1, void *stream);

/* SparseTensor.dense() (mssvt_utils.py:50-62): out (B, C, D, H, W), zero-filled then scattered */
int mssvt_dense_scatter(int num_rows, const int *num_rows_dev, int batch_size, int C, int D, int H,
                        int W, const float *features, const int *indices, float *out, void *stream);

int mssvt_sizeof_attn_shape(void);
int mssvt_sizeof_ffn_shape(void);

/* ---- DynamicVFE (pcdet/models/backbones_3d/vfe/dynamic_vfe.py:71-130): the producer of voxel_features /
 * voxel_coords.  Inference (eval-mode BatchNorm). */

/* words of the occupancy bitmap used by mssvt_vfe_voxelize: batch * gx * gy * ceil(gz / 32) */
long long mssvt_vfe_bitmap_words(int batch_size, int gx, int gy, int gz);

/* Dynamic voxelisation (dynamic_vfe.py:85-93, 111-116; replaces torch.unique(sorted, return_inverse)).
 * points (P, point_stride) fp32 rows [batch, x, y, z, ...].  Scratch (int): bitmap (words), counts (words),
 * base (words + 1), scan_workspace ((words + 1) / 1024 + 2).  Outputs: point_voxel (P) voxel row of every
 * point or -1 (outside the range); voxel_coords (P, 4) [b, z, y, x], rows [0, num_voxels) in ascending
 * (b, x, y, z) key order (the order of torch.unique); num_voxels = base[words] stays on the device;
 * xyz_sum (P, 4), optional: per-voxel (sum x, sum y, sum z, #points) for the cluster centre. */
int mssvt_vfe_voxelize(int num_points, const float *points, int point_stride, int batch_size, int gx, int gy,
                       int gz, const float *voxel_size, const float *range_min, int *bitmap, int *counts,
                       int *base, int *scan_workspace, int *point_voxel, int *voxel_coords, float *xyz_sum,
                       void *stream);

/* Point-feature network + per-voxel max (dynamic_vfe.py:95-108, 124-130; replaces torch_scatter.scatter_mean /
 * scatter_max and the Linear + BatchNorm1d + ReLU stack).  The caller folds each eval-mode BatchNorm1d into
 * its Linear.  One or two layers: w0 (c0, in0), b0 (c0); w1 (c1, 2 * c0), b1 (c1), or c1 = 0.
 * in0 = num_point_features + 3 (cluster centre) + 3 (voxel centre) + 1 (distance) as enabled, <= 20.
 * centre_offset = voxel_size / 2 + range_min.  scratch: voxel_capacity * c0 floats (two layers only).
 * out (voxel_capacity, c_last); rows >= num_voxels are zero. */
int mssvt_vfe_features(int num_points, const float *points, int point_stride, int num_point_features,
                       int with_cluster_center, int with_voxel_center, int with_distance, const float *voxel_size,
                       const float *range_min, const float *centre_offset, const int *point_voxel,
                       const float *xyz_sum, int voxel_capacity, const float *w0, const float *b0, int c0,
                       const float *w1, const float *b1, int c1, float *scratch, float *out, void *stream);

/* ---- training path: forward + backward of the window attention on the compact (ragged)
 * form, replacing autograd over the padded tensors of MixedScaleAttention.forward (mssvt_utils.py:100-157:
 * softmax(q k^T * scale + (-100) * key_mask) v per window and head, with the masked slots of a window all
 * holding one key).  Windows are CSR lists: queries of window w = rows [q_off[w], q_off[w + 1]) of q, keys =
 * rows [key_off[w], key_off[w + 1]) of k / v; key_mult[w] > 0: the LAST key of the window is the masked one
 * (additive -100) and stands for key_mult[w] identical slots.  q_win (num_queries) / k_win (num_keys) = window
 * of each row.  Rows have heads * head_dim channels (head_dim in {8, 16, 32}) at row strides ld* (floats,
 * multiples of 4; 16-byte aligned bases).  lse (num_queries, heads): log-sum-exp of every softmax row. */
int mssvt_ragged_attention_fwd(int heads, int head_dim, float scale, int num_queries, const int *q_win,
                               const int *key_off, const int *key_mult, const float *q, int ldq, const float *k,
                               int ldk, const float *v, int ldv, float *out, int ldo, float *lse, void *stream);

/* Backward of the above: grad_q / grad_k / grad_v from grad_out (autograd of mssvt_utils.py:123-139).  Every
 * key row belongs to one window, so dK / dV are accumulated by the thread that owns the row: no atomics,
 * deterministic.  delta (num_queries, heads) is scratch (grad_out . out per softmax row). */
int mssvt_ragged_attention_bwd(int heads, int head_dim, float scale, int num_queries, int num_keys, const int *q_win,
                               const int *k_win, const int *q_off, const int *key_off, const int *key_mult,
                               const float *q, int ldq, const float *k, int ldk, const float *v, int ldv,
                               const float *out, int ldo, const float *lse, const float *grad_out, int ldgo,
                               float *delta, float *grad_q, int ldgq, float *grad_k, int ldgk, float *grad_v,
                               int ldgv, void *stream);

/* Three-NN feature interpolation + merge back to voxels (mssvt_backbone.py:318-333; the reference blends in
 * torch): out[v] = sum_j weights[v, j] * rows[src[v, j]] over the (num_voxels, 3) maps; src < 0: a padded query
 * slot = zero row; src[v, 0] == -2: voxel not covered by any window, out[v] = x[v] (quirk Q5).  C % 4 == 0. */
int mssvt_interp_merge_fwd(int num_voxels, int C, const int *src, const float *weights, const float *rows,
                           const float *x, float *out, void *stream);

/* Its backward: grad_rows (num_rows, C) zeroed then accumulated with vector atomics, grad_x (num_voxels, C)
 * written in full (grad_out on uncovered voxels, zero elsewhere). */
int mssvt_interp_merge_bwd(int num_voxels, int C, int num_rows, const int *src, const float *weights,
                           const float *grad_out, float *grad_rows, float *grad_x, void *stream);

/* Gather + positional embedding of compact rows (mssvt_backbone.py:288-300, the one-layer pos_proj of the two-window
 * block = Conv1d(6 -> C, 1) + ReLU, on the rows that exist instead of the padded (W, C, n) tensors):
 * out[r, 0:cs] = xn[rows[r], c0:c0+cs] + relu(pos_w[c0:c0+cs] . [xyz[rows[r]] - centre[win[r]] | centre[win[r]]] + pos_b)
 * with the relative offset zeroed where masked[r] (the key that stands for the masked slots); rows[r] < 0: zero
 * features at position 0.  xn (N, C), xyz (N, 3), centre (W, 3), out (num_rows, cs), cs in {32, 64}.
 * xn == NULL: the embedding alone (first layer of the compress block's two-layer pos_proj, mssvt_backbone.py:51-54);
 * pos_w == NULL: the row gather alone. */
int mssvt_embed_rows_fwd(int num_rows, int c0, int cs, int C, const int *rows, const int *win,
                         const unsigned char *masked, const float *xn, const float *xyz, const float *centre,
                         const float *pos_w, const float *pos_b, float *out, void *stream);

/* Its backward: ACCUMULATES into grad_xn (N, C) (vector atomics), grad_w (C, 6) and grad_b (C) -- the caller zeroes
 * them once and calls this for every row set that read the same xn / pos_proj. */
int mssvt_embed_rows_bwd(int num_rows, int c0, int cs, int C, const int *rows, const int *win,
                         const unsigned char *masked, const float *xyz, const float *centre, const float *pos_w,
                         const float *pos_b, const float *grad_out, float *grad_xn, float *grad_w, float *grad_b,
                         void *stream);

/* Backward of mssvt_layernorm (nn.LayerNorm, mssvt_backbone.py:210, 340): grad_x (num_rows, C), grad_gamma / grad_beta
 * (C) zeroed then accumulated; the row statistics are recomputed from x.  C in {64, 128}. */
int mssvt_layernorm_bwd(int num_rows, int C, const float *x, const float *gamma, float eps, const float *grad_y,
                        float *grad_x, float *grad_gamma, float *grad_beta, void *stream);

/* nn.Linear over rows for the training path (the q / kv / output projections of mssvt_utils.py:108-146 over compact
 * window rows, linear1 / linear2 of the FFN, mssvt_backbone.py:340-344): y (num_rows, N) = x (num_rows, K) w^T + bias,
 * optional ReLU; w (N, K) row-major as nn.Linear stores it, bias may be NULL; K, N in {32, 64, 128}; row strides ldx / ldy
 * in floats (multiples of 4, 16-byte aligned bases).  terms = 3: split-TF32 operands on mma.sync (fp32-grade), 1: plain
 * TF32.  The input gradient is the same call on the transposed weight (dx = dy w). */
int mssvt_linear_rows_fwd(int num_rows, int K, int N, int terms, const float *x, int ldx, const float *w,
                          const float *bias, int relu, float *y, int ldy, void *stream);

/* Weight / bias gradient of the above: grad_w (N, K) = grad_y^T x, grad_b (N) = column sums of grad_y (may be NULL), one
 * pass over the rows, per-CTA partial sums reduced in a fixed order (deterministic).  workspace: at least
 * mssvt_linear_rows_wgrad_workspace_floats(K, N) floats. */
long long mssvt_linear_rows_wgrad_workspace_floats(int K, int N);
int mssvt_linear_rows_wgrad(int num_rows, int K, int N, int terms, const float *grad_y, int ldgy, const float *x,
                            int ldx, float *workspace, float *grad_w, float *grad_b, void *stream);

/* Max over the rows of every window (the max-pooled query of the compress block, mssvt_backbone.py:373, on compact
 * rows: rows [key_off[w], key_off[w + 1]) of `rows` (R, C) belong to window w).  out (num_windows, C); arg
 * (num_windows, C) int: the row that supplied each channel (first on ties, -1 for an empty window). */
int mssvt_segment_max_fwd(int num_windows, int C, const int *key_off, const float *rows, float *out, int *arg,
                          void *stream);

/* Its backward: grad_rows (num_rows, C) written in full (grad_out[k_win[r]] where row r supplied the channel, else 0). */
int mssvt_segment_max_bwd(int num_rows, int C, const int *k_win, const int *arg, const float *grad_out,
                          float *grad_rows, void *stream);

/* ---- compact window lists of the training path, built on the device (mssvt_b200/csrc/train_lists.cu) from the maps of
 * mssvt_block_geometry.  counts (cap, 4) = {#real queries, #distinct keys of head group 0, of group 1, 0} per window (key
 * counts zeroed for windows without a query); mult (2, cap) = multiplicity of each group's masked key.  The offsets are
 * mssvt_exclusive_scan over the columns of counts; the caller reads the three totals once, allocates the lists and calls
 * mssvt_ragged_lists_fill: q_rows / q_win (#queries), per group k_rows / k_win / k_masked (#keys). */
int mssvt_ragged_lists_count(int win_capacity, const int *win_count_total, const int *meta, int *counts, int *mult,
                             void *stream);
int mssvt_ragged_lists_fill(int win_capacity, const int *win_count_total, int nq, int K, const int *counts,
                            const int *mult, const int *q_off, const int *key_off0, const int *key_off1,
                            const int *q_row, const int *rep_row, int *q_rows, int *q_win, int *k_rows0, int *k_win0,
                            unsigned char *k_masked0, int *k_rows1, int *k_win1, unsigned char *k_masked1,
                            void *stream);

/* src (num_voxels, 3) / weights (num_voxels, 3) of mssvt_interp_merge_fwd from vox_slot / nn_idx / nn_w of the geometry:
 * compact query ids (q_off[window] + slot), -1 for a padded query slot, -2 for a voxel outside every window. */
int mssvt_ragged_merge_map(int num_voxels, int max_win1, const int *vox_slot, const int *meta, const int *q_off,
                           const unsigned char *nn_idx, const float *nn_w, int *src, float *weights, void *stream);

/* The same for the compress block (keys of a window = the voxels of k_row (cap, max_win1) + one pad key when slots are left,
 * quirk Q6): counts (cap), mult (cap) = number of padded slots; rows (#keys) = voxel row or -1 for the pad key, k_win. */
int mssvt_compress_lists_count(int win_capacity, const int *win_count_total, int max_win1, const int *k_row,
                               int *counts, int *mult, void *stream);
int mssvt_compress_lists_fill(int win_capacity, const int *win_count_total, int max_win1, const int *counts,
                              const int *mult, const int *key_off, const int *k_row, int *rows, int *k_win,
                              void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MSSVT_B200_H */
