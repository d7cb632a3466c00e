/*
 * b2s.h -- C ABI of libb2s.so, the B200-native (sm_100a) sparse-voxel hot path.
 *
 * This is the drop-in boundary for the two reference interfaces of the path
 * (SURVEY.md section 8(b)):
 *   #1  the MinkowskiEngine Python surface used by minsu3d
 *       (minsu3d/model/module/common.py:12-93, backbone.py:14-38,
 *        general_model.py:187-191, data/dataset/general_dataset.py:159-163)
 *   #2  the COMMON_OPS pybind module
 *       (minsu3d/common_ops/src/common_ops_api.cpp:6-29)
 *
 * Conventions
 *   - plain pointers + sizes, no torch types.  Every pointer is a DEVICE pointer
 *     unless its name starts with h_.
 *   - every entry point takes the caller's stream and never synchronises it;
 *     sizes that are only known on the device are written to a device int32
 *     (d_count ...) which the host reads back when it needs a shape.
 *   - the library never allocates: scratch comes from the caller (ws, ws_bytes);
 *     each op with scratch has a *_ws_bytes() query.
 *   - return value: 0 = ok, <0 = B2S_E_* (no exit(), cf. the reference's
 *     exit(-1) in bfs_cluster.cu:82-86 / hierarchical_aggregation.cu:198-203).
 *   - re-entrant, no global state: one process per GPU works unchanged.
 */
#ifndef B2S_H
#define B2S_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* b2s_stream_t;

#define B2S_OK 0
#define B2S_E_INVALID (-1)   /* bad argument (shape, null pointer, unsupported size) */
#define B2S_E_WORKSPACE (-2) /* workspace too small */
#define B2S_E_LAUNCH (-3)    /* cudaGetLastError() != cudaSuccess after a launch */
#define B2S_E_RANGE (-4)     /* coordinate outside the packable range */

/* library / build info: returns the compiled SM arch (100) */
int b2s_version(void);
const char* b2s_last_error_string(void);

/* ------------------------------------------------------------------------------------------------
 * T1 / V1 -- coordinate hash: insert + first-occurrence unique + inverse map.
 * Replaces ME.SparseTensor's coordinate-manager insert and ME.utils.sparse_quantize's
 * unique/inverse (general_dataset.py:159-163, general_model.py:187-189), and -- with
 * quant > 1 -- the strided coordinate map of MinkowskiConvolution(k=2,s=2) (common.py:69).
 *
 *   coords      [n,4] int32 (b,x,y,z)
 *   quant       spatial quantisation: key coords are floor(c/quant)*quant (1 = identity)
 *   table_keys  [cap] uint64, table_vals [cap] int32, cap = b2s_hash_capacity(n) (power of two)
 *   unique_idx  [n] int32 (first m valid): input row of each unique coordinate, ascending
 *   inverse     [n] int32: unique row of every input row
 *   out_coords  [n,4] int32 (first m valid): quantised coordinates of the unique rows
 *   d_count     int32[2] on device: {m, range_error_flag}
 *   d_n         optional device int32: the real number of input rows (<= n, n then being an upper bound);
 *               lets a chain of strided maps (a coordinate pyramid) be built without host round trips
 * After the call the table maps packed coordinate -> unique row.
 * ---------------------------------------------------------------------------------------------- */
int64_t b2s_hash_capacity(int64_t n);
size_t b2s_coord_unique_ws_bytes(int64_t n);
int b2s_coord_unique(const int32_t* coords, int64_t n, int32_t quant,
                     uint64_t* table_keys, int32_t* table_vals, int64_t cap,
                     int32_t* unique_idx, int32_t* inverse, int32_t* out_coords,
                     int32_t* d_count, const int32_t* d_n, void* ws, size_t ws_bytes, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * T2 -- kernel map (output-stationary neighbour table).
 * nbr[o*K + kidx] = input row whose coordinate equals out_coords[o] + offset(kidx), or -1.
 * Offsets enumerate x fastest: odd ksize: (i-(ksize-1)/2)*dil, even ksize: i*dil
 * (SURVEY.md appendix A.4).  K = ksize^3.  The table is the input map's hash table.
 * ---------------------------------------------------------------------------------------------- */
int b2s_kernel_map(const int32_t* out_coords, int64_t n_out, int32_t ksize, int32_t dil,
                   const uint64_t* table_keys, const int32_t* table_vals, int64_t cap,
                   int32_t* nbr, uint32_t* tile_mask, b2s_stream_t stream);

/* Mask-sorted tile order of a 3^3 neighbour table (K = 27) for b2s_conv_table_rows: rows are stably sorted by
 * (which faces of the stencil hold a neighbour, 27-bit neighbour mask) so that the 128-row tiles of the
 * tcgen05 convolution see few distinct offsets (the row order MinkowskiEngine defines -- first occurrence,
 * SURVEY appendix A.2 -- mixes floors, walls and edges in every tile).  Purely a schedule: results unchanged.
 *   row_perm   [n_out] int32: original row of sorted position t
 *   nbr_sorted [n_out, K] int32: nbr[row_perm[t], :]
 *   tile_mask  [ceil(n_out/128)] uint32: active-offset masks of the sorted tiles                          */
size_t b2s_tile_order_ws_bytes(int64_t n_out);
int b2s_tile_order(const int32_t* nbr, int64_t n_out, int32_t K, int32_t* row_perm, int32_t* nbr_sorted,
                   uint32_t* tile_mask, void* ws, size_t ws_bytes, b2s_stream_t stream);

/* Canonical per-offset pair lists (sorted by kidx, then by output row) from a neighbour table.
 *   pair_in/pair_out [>= number of pairs] int32, k_offsets [K+1] int32 (CSR over kidx),
 *   d_count int32[1] = total pairs.  Upper bound on pairs: n_out*K.                              */
size_t b2s_pairs_ws_bytes(int64_t n_out, int32_t K);
int b2s_pairs_from_nbr(const int32_t* nbr, int64_t n_out, int32_t K, int64_t pair_capacity,
                       int32_t* pair_in, int32_t* pair_out, int32_t* k_offsets, int32_t* d_count,
                       void* ws, size_t ws_bytes, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * T3 / T4 -- sparse convolution (MinkowskiConvolution / MinkowskiConvolutionTranspose,
 * common.py:12,31,37,40,69,77; backbone.py:14).  fp32 in / fp32 out.
 *
 * b2s_conv_table: out[o,:] = sum_k  A[nbr[o*K+k], :] @ Wk      (k ascending, -1 rows skipped)
 *     Wk = W[kk] (c_in x c_out, row-major) or its transpose when w_transposed (then W is
 *     [K, c_out, c_in] and the contraction runs over W's last axis), kk = k or K-1-k (k_reversed).
 *     nbr == NULL means K == 1 and the identity map (1x1 convolution = dense matmul).
 *     tile_mask (may be NULL): the per-128-row active-offset masks b2s_kernel_map wrote for this nbr
 *     (K <= 32); the tcgen05 path then skips empty offsets without scanning the table tile.
 *     forward: (A=in, W) ; data gradient of a stride-1 conv: (A=grad_out, w_transposed, k_reversed).
 * b2s_conv_pairs: for every pair p in [k_offsets[k], k_offsets[k+1]):
 *     out[dst[p], :] = A[src[p], :] @ Wk   (each dst row appears once; plain store)
 *     transposed-conv forward and strided-conv data gradient.
 * b2s_conv_wgrad: gW[k] (+)= sum_p A[src[p], :]^T @ G[dst[p], :]   (gW zeroed by the callee)
 * algo: 0 = auto (tcgen05 3xTF32 when c_in, c_out are multiples of 16 and c_out <= 256, else fp32 FMA),
 *       1 = fp32 FMA (SIMT), 2 = tcgen05 3xTF32 (fp32-class accuracy), 3 = tcgen05 plain TF32,
 *       4 = warp-stream mma.sync 3xTF32 kernel (csrc/conv_ws.cu; table mode, c_in, c_out in {16, 32}, K <= 32 only;
 *           an explored alternative for the narrow layers, not chosen by auto).
 * ---------------------------------------------------------------------------------------------- */
size_t b2s_conv_ws_bytes(int32_t K, int32_t c_in, int32_t c_out); /* scratch: packed weights + split partial sums */
/* Packed weights.  The tcgen05 kernels read W as SWIZZLE_64B shared-memory images split into TF32 hi / lo parts.
 * b2s_conv_pack writes BOTH orientations of W[K][c_in][c_out] -- [forward hi | forward lo | transposed hi |
 * transposed lo], b2s_conv_packed_floats() floats (0 = shape not taken by the tcgen05 path) -- in one launch; a caller
 * that keeps the result while W is unchanged (once per optimizer step) passes it as `Wp` to the products below and
 * no packing happens there.  Wp == NULL: the product packs the one image it needs into its workspace (round-1
 * behaviour: 170 extra launches per PointGroup step).                                                          */
int64_t b2s_conv_packed_floats(int32_t K, int32_t c_in, int32_t c_out);
int b2s_conv_pack(const float* W, float* Wp, int32_t K, int32_t c_in, int32_t c_out, b2s_stream_t stream);
/* The same for many layers in ONE launch: desc [n_desc, 6] int64 in device memory, row = {W, Wp, K, c_in, c_out, start},
 * start = running sum of K*c_in*c_out over the preceding rows, total = the sum over all rows.                   */
int b2s_conv_pack_multi(const int64_t* desc, int32_t n_desc, int64_t total, b2s_stream_t stream);
/* add_src (optional, [n_out, c_out]): residual added to the result (common.py:48), folded into the epilogue of the
 * tcgen05 kernel.                                                                                              */
int b2s_conv_table(const float* A, const float* W, const float* Wp, const int32_t* nbr, const uint32_t* tile_mask,
                   const float* add_src,
                   float* out, int64_t n_out, int32_t K, int32_t c_in, int32_t c_out,
                   int32_t w_transposed, int32_t k_reversed, int32_t algo,
                   void* ws, size_t ws_bytes, b2s_stream_t stream);
/* b2s_conv_table over mask-sorted tiles (b2s_tile_order below): row t of the permuted table is stored at
 * out[out_rows[t], :]; same sum, same k order, fewer all-empty (tile, offset) slabs.  tcgen05 path only
 * (c_in, c_out multiples of 16, K <= 32; algo 0/2 = 3xTF32, 3 = TF32; algo 1 is rejected).            */
int b2s_conv_table_rows(const float* A, const float* W, const float* Wp, const int32_t* nbr_sorted,
                        const uint32_t* tile_mask, const int32_t* out_rows, const float* add_src, float* out,
                        int64_t n_out, int32_t K, int32_t c_in, int32_t c_out, int32_t w_transposed,
                        int32_t k_reversed, int32_t algo, void* ws, size_t ws_bytes, b2s_stream_t stream);
int b2s_conv_pairs(const float* A, const float* W, const float* Wp, const int32_t* src, const int32_t* dst,
                   const int32_t* k_offsets, float* out, int32_t K, int32_t c_in, int32_t c_out,
                   int32_t w_transposed, int64_t max_pairs, int32_t algo,
                   void* ws, size_t ws_bytes, b2s_stream_t stream);
int b2s_conv_wgrad(const float* A, const float* G, const int32_t* src, const int32_t* dst,
                   const int32_t* k_offsets, float* gW, int32_t K, int32_t c_a, int32_t c_g,
                   int64_t max_pairs, int32_t algo, b2s_stream_t stream);
/* The same product with a caller workspace of b2s_conv_wgrad_ws_bytes(K, c_a, c_g) bytes (independent of the pair
 * count).  With channel counts that are multiples of 16 (and algo != 1) it runs the deterministic tensor-core kernel
 * (csrc/wgrad_det.cu): per-warp partial tiles in the workspace, summed in a fixed order -- no atomics, the same bits
 * on every run; other shapes fall through to b2s_conv_wgrad.                                                     */
size_t b2s_conv_wgrad_ws_bytes(int32_t K, int32_t c_a, int32_t c_g);
int b2s_conv_wgrad_ws(const float* A, const float* G, const int32_t* src, const int32_t* dst,
                      const int32_t* k_offsets, float* gW, int32_t K, int32_t c_a, int32_t c_g, int64_t max_pairs,
                      int32_t algo, void* ws, size_t ws_bytes, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * T5 -- fused BatchNorm(+ReLU) on sparse-tensor features (MinkowskiBatchNorm + MinkowskiReLU,
 * common.py:13-14,35-39).  Training statistics are biased batch statistics like
 * torch.nn.BatchNorm1d.
 * ---------------------------------------------------------------------------------------------- */
size_t b2s_bn_ws_bytes(int64_t n, int32_t c);
/* batch statistics: mean, biased variance (optional), rstd = 1/sqrt(var+eps) (optional); when
 * running_mean/var are given they are updated in place with `momentum` and the unbiased variance. */
/* counter: one int32 on the device that is ZERO between launches; a dedicated buffer per stream, not part of the
 * shared workspace.  b2s_bn_stats: the last block to finish runs the second reduction stage and resets it.
 * b2s_bn_forward / b2s_bn_backward[_add]: ONE cooperative launch each -- per-CTA column sums, a device-wide barrier on
 * this word (arrivals in the low half, departures in the high half, the last CTA to leave resets it), then every CTA
 * sums the partials in the same fixed order and normalises / differentiates the rows it has just read.  Results are
 * bit-identical to b2s_bn_stats + b2s_bn_apply; B2S_BN_FUSED=0 selects the two-kernel path.                      */
int b2s_bn_stats(const float* x, int64_t n, int32_t c, float eps, float momentum, float* running_mean,
                 float* running_var, float* mean, float* var_biased, float* rstd, int32_t* counter,
                 void* ws, size_t ws_bytes, b2s_stream_t stream);
/* b2s_bn_stats followed by b2s_bn_apply in one call and one launch (training-mode forward of BatchNorm(+ReLU)). */
int b2s_bn_forward(const float* x, int64_t n, int32_t c, float eps, float momentum,
                   float* running_mean, float* running_var, const float* gamma, const float* beta,
                   int32_t relu, float* y, float* mean, float* rstd, int32_t* counter,
                   void* ws, size_t ws_bytes, b2s_stream_t stream);
int b2s_bn_apply(const float* x, int64_t n, int32_t c, const float* mean, const float* rstd,
                 const float* gamma, const float* beta, int32_t relu, float* y, b2s_stream_t stream);
/* backward of y = relu?(gamma*(x-mean)*rstd + beta) in training mode:
 * dgamma, dbeta [C]; dx [n,C].  y is the forward output (for the ReLU mask).                     */
int b2s_bn_backward(const float* x, const float* y, const float* dy, int64_t n, int32_t c,
                    const float* mean, const float* rstd, const float* gamma, int32_t relu,
                    int32_t training, float* dx, float* dgamma, float* dbeta, int32_t* counter,
                    void* ws, size_t ws_bytes, b2s_stream_t stream);
/* the same with dx += add_src ([n,C]; the gradient arriving over a residual shortcut) in the dx pass */
int b2s_bn_backward_add(const float* x, const float* y, const float* dy, const float* add_src, int64_t n, int32_t c,
                        const float* mean, const float* rstd, const float* gamma, int32_t relu,
                        int32_t training, float* dx, float* dgamma, float* dbeta, int32_t* counter,
                        void* ws, size_t ws_bytes, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * T3 + T5 host-side fusion -- one MinkUNet residual block per call (minsu3d/model/module/common.py:21-50):
 *     out = conv3(relu(bn2(conv3(relu(bn1(x)))))) + shortcut(x),  shortcut = x or x @ Wds (1x1 convolution)
 * Same kernels, order and operands as the single-op entry points above (bit-identical results); the point is
 * one library call / one autograd node per block: the train step is bound by host launch overhead.
 * Training-mode BatchNorm only (batch statistics, running statistics updated in place).
 *   nbr/tile_mask: the block's 3^3 map (b2s_kernel_map); nbr_sorted/tile_mask_sorted/row_perm: optional
 *   b2s_tile_order schedule (all NULL = first-occurrence order).
 *   forward saves y1 = relu(bn1(x)) [n,c_in], z1 = conv(y1) [n,c_out], y2 = relu(bn2(z1)) [n,c_out],
 *   stats1 [2,c_in] / stats2 [2,c_out] = (mean, rstd); tmp [n,c_out] is scratch for the 1x1 shortcut (Wds != NULL).
 *   backward: pair lists of the same map (b2s_pairs_from_nbr) for the weight gradients; ident [n] = 0..n-1 and
 *   ident_koff = {0, n} drive the 1x1 weight gradient; dgb1 [2,c_in] / dgb2 [2,c_out] = (dgamma, dbeta);
 *   tmp_a [n,max(c_in,c_out)], tmp_b [n,c_out] and tmp_c [n,c_in] are scratch.
 *   Wp1 / Wp2 / Wpds (each may be NULL): b2s_conv_pack images of W1 / W2 / Wds.
 *   The shortcut is added in the epilogue of the second convolution (forward) and in the dx pass of the first
 *   BatchNorm's backward: no separate add kernels.
 * ---------------------------------------------------------------------------------------------- */
size_t b2s_resblock_ws_bytes(int32_t K, int32_t c_in, int32_t c_out);
int b2s_resblock_forward(const float* x, int64_t n, int32_t c_in, int32_t c_out,
                         const float* gamma1, const float* beta1, float* rmean1, float* rvar1, const float* W1,
                         const float* gamma2, const float* beta2, float* rmean2, float* rvar2, const float* W2,
                         const float* Wds, const float* Wp1, const float* Wp2, const float* Wpds,
                         float eps1, float mom1, float eps2, float mom2,
                         const int32_t* nbr, const uint32_t* tile_mask, const int32_t* nbr_sorted,
                         const uint32_t* tile_mask_sorted, const int32_t* row_perm, int32_t K,
                         float* y1, float* stats1, float* z1, float* y2, float* stats2, float* out, float* tmp,
                         int32_t* bn_counter, int32_t algo, void* ws, size_t ws_bytes, b2s_stream_t stream);
int b2s_resblock_backward(const float* gout, const float* x, const float* y1, const float* z1, const float* y2,
                          const float* stats1, const float* stats2, const float* gamma1, const float* gamma2,
                          const float* W1, const float* W2, const float* Wds, const float* Wp1, const float* Wp2,
                          const float* Wpds, int64_t n, int32_t c_in, int32_t c_out,
                          const int32_t* nbr, const uint32_t* tile_mask, const int32_t* nbr_sorted,
                          const uint32_t* tile_mask_sorted, const int32_t* row_perm, int32_t K,
                          const int32_t* pair_in, const int32_t* pair_out, const int32_t* k_offsets, int64_t max_pairs,
                          const int32_t* ident, const int32_t* ident_koff,
                          float* gx, float* gW1, float* gW2, float* gWds, float* dgb1, float* dgb2,
                          float* tmp_a, float* tmp_b, float* tmp_c,
                          int32_t* bn_counter, int32_t algo, void* ws, size_t ws_bytes, b2s_stream_t stream);

/* BN -> ReLU -> strided convolution (mode 0, common.py:67-69) or transposed convolution (mode 1, common.py:75-77)
 * of a U-Net level as one call; the strided map fine -> coarse is passed as its table nbr[n_coarse, K] and its
 * pair lists (pair_in = fine row, pair_out = coarse row).  x has n_fine rows in mode 0 and n_coarse rows in mode 1.
 * Saves y = relu(bn(x)) and stats [2,c_in] = (mean, rstd); backward needs tmp [n_x, c_in] scratch, dgb [2,c_in].
 * Wp (may be NULL): b2s_conv_pack image of W.                                                                    */
int b2s_bnconv_forward(const float* x, int64_t n_x, int32_t c_in, int32_t c_out, const float* gamma, const float* beta,
                       float* rmean, float* rvar, float eps, float mom, const float* W, const float* Wp, int32_t mode,
                       const int32_t* nbr, const uint32_t* tile_mask, const int32_t* pair_in, const int32_t* pair_out,
                       const int32_t* k_offsets, int64_t max_pairs, int64_t n_coarse, int64_t n_fine, int32_t K,
                       float* y, float* stats, float* out, int32_t* bn_counter, int32_t algo, void* ws, size_t ws_bytes,
                       b2s_stream_t stream);
int b2s_bnconv_backward(const float* gout, const float* x, const float* y, const float* stats, const float* gamma,
                        const float* W, const float* Wp, int64_t n_x, int32_t c_in, int32_t c_out, int32_t mode,
                        const int32_t* nbr,
                        const uint32_t* tile_mask, const int32_t* pair_in, const int32_t* pair_out,
                        const int32_t* k_offsets, int64_t max_pairs, int64_t n_coarse, int64_t n_fine, int32_t K,
                        float* gx, float* gW, float* dgb, float* tmp, int32_t* bn_counter, int32_t algo, void* ws,
                        size_t ws_bytes, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * V2 -- devoxelise gather and its scatter-add gradient (backbone.py:40, pointgroup.py:88).
 * idx is int64 (voxel_point_map dtype, data_module.py:65).
 * ---------------------------------------------------------------------------------------------- */
int b2s_gather_rows(const float* feat, const int64_t* idx, int64_t n, int32_t c, float* out,
                    b2s_stream_t stream);
int b2s_scatter_add_rows(const float* grad, const int64_t* idx, int64_t n, int32_t c,
                         float* gfeat /* pre-zeroed [m,c] */, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * C1 -- ballquery_batch_p (bfs_cluster.cu:15-91, functions/common_ops.py:11-47).
 * Neighbours k of p: same scene, fma(dz,dz,fma(dx,dx,dy*dy)) < r*r in fp32, ascending k,
 * at most 1000 (the lowest 1000).  Canonical CSR: start = exclusive scan of len.
 *   count: fills start_len [n,2] int32 and d_count[0] = nActive; keeps a cell grid in ws.
 *   fill : writes idx [nActive] int32 using the same ws.
 * ---------------------------------------------------------------------------------------------- */
size_t b2s_ballquery_ws_bytes(int64_t n);
int b2s_ballquery_count(const float* xyz, const uint8_t* batch_idxs, const int32_t* batch_offsets,
                        int64_t n, int32_t n_batch, float radius, int32_t* start_len,
                        int32_t* d_count, void* ws, size_t ws_bytes, b2s_stream_t stream);
int b2s_ballquery_fill(const float* xyz, const uint8_t* batch_idxs, const int32_t* batch_offsets,
                       int64_t n, int32_t n_batch, float radius, const int32_t* start_len,
                       int32_t* idx, void* ws, size_t ws_bytes, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * C2 / C3 / C4 -- BFS clustering on the GPU (bfs_cluster.cpp:28-187,
 * hierarchical_aggregation.cpp:8-97).  Bit-exact with the reference's sequential BFS:
 * cluster = set reachable from the lowest unvisited seed, clusters in seed order,
 * points inside a cluster in BFS visit order.
 *
 * b2s_cluster_label : comp[v] = lowest index that reaches v (labels == NULL: no label test).
 * b2s_cluster_select: sizes, keep mask, compaction.  mode 0: keep size >= thr_i (pg);
 *     mode 1: keep (float)size >= thr_f (sg).  mode 3: mode 1 with one threshold per class,
 *     keep (float)size >= point_num_avg[labels[seed]] (all SoftGroup classes clustered in one call:
 *     the caller stacks the per-class point sets and gives each class its own batch indices).
 *     mode 2 (HAIS): class thresholds from
 *     point_num_avg; cls: 0 = dropped, 1 = kept fragment, 2 = primary; fragments (size < high)
 *     are also listed when want_fragments.
 *     Outputs: d_count = {nCluster, sumNPoint}, cluster_offsets [nCluster+1], seeds [nCluster].
 * b2s_cluster_order : cluster_idxs [sumNPoint,2] (cluster id, point) in BFS visit order; n_active (length of
 *     nbr_idx) picks the per-cluster-CTA variant (sparse, deep graphs) or the device-wide one (dense graphs).
 * ---------------------------------------------------------------------------------------------- */
size_t b2s_cluster_ws_bytes(int64_t n);
int b2s_cluster_label(const int32_t* nbr_idx, const int32_t* start_len, const int16_t* labels,
                      int64_t n, int32_t* comp, void* ws, size_t ws_bytes, b2s_stream_t stream);
int b2s_cluster_select(const int32_t* comp, const int16_t* labels, int64_t n, int32_t mode,
                       int32_t thr_i, float thr_f, const float* point_num_avg, int32_t group,
                       int32_t* cluster_offsets, int32_t* seeds, int32_t* d_count,
                       void* ws, size_t ws_bytes, b2s_stream_t stream);
int b2s_cluster_order(const int32_t* nbr_idx, const int32_t* start_len, const int16_t* labels,
                      const int32_t* comp, int64_t n, int64_t n_active, const int32_t* cluster_offsets,
                      const int32_t* seeds, int32_t n_cluster, int32_t* cluster_idxs,
                      void* ws, size_t ws_bytes, b2s_stream_t stream);
/* HAIS: centres [nCluster,5] = (sum_x/size, sum_y/size, sum_z/size, cls, batch), sums taken
 * sequentially in BFS order in fp32 (hierarchical_aggregation.cpp:13-37,85-89).                  */
int b2s_cluster_centers(const int32_t* cluster_idxs, const int32_t* cluster_offsets,
                        int32_t n_cluster, const float* coords, const int16_t* labels,
                        const uint8_t* batch_idxs, float* centers, b2s_stream_t stream);
/* HAIS set aggregation (hierarchical_aggregation.cu:20-91): assign[f] = primary absorbing
 * fragment f or -1; then concatenation primary points + absorbed fragments (ascending f).        */
int b2s_ha_assign(const float* frag_centers, int32_t n_frag, const float* prim_centers,
                  const int32_t* prim_offsets, int32_t n_prim, const float* radius_avg,
                  int32_t* assign, b2s_stream_t stream);
size_t b2s_ha_concat_ws_bytes(int32_t n_frag, int32_t n_prim);
int b2s_ha_concat(const int32_t* frag_idxs, const int32_t* frag_offsets, int32_t n_frag,
                  const int32_t* prim_idxs, const int32_t* prim_offsets, int32_t n_prim,
                  const int32_t* assign, int32_t* out_idxs, int32_t* out_offsets,
                  void* ws, size_t ws_bytes, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * S1 / S2 / S3 -- segmented reductions (sec_mean.cu:12-85, roipool.cu:12-120).
 * ---------------------------------------------------------------------------------------------- */
int b2s_sec_mean(const float* inp, const int32_t* offsets, float* out, int32_t n_seg, int32_t c,
                 b2s_stream_t stream);
int b2s_sec_min(const float* inp, const int32_t* offsets, float* out, int32_t n_seg, int32_t c,
                b2s_stream_t stream);
int b2s_sec_max(const float* inp, const int32_t* offsets, float* out, int32_t n_seg, int32_t c,
                b2s_stream_t stream);
int b2s_roipool_fp(const float* feats, const int32_t* offsets, float* out, int32_t* maxidx,
                   int32_t n_seg, int32_t c, b2s_stream_t stream);
/* the same for long segments (proposals of ~10k points): several CTAs per segment + an in-order combine */
size_t b2s_roipool_ws_bytes(int32_t n_seg, int32_t c);
int b2s_roipool_fp_ws(const float* feats, const int32_t* offsets, float* out, int32_t* maxidx, int32_t n_seg,
                      int32_t c, int64_t n_rows, void* ws, size_t ws_bytes, b2s_stream_t stream);
int b2s_roipool_bp(float* d_feats, const int32_t* offsets, const int32_t* maxidx,
                   const float* d_out, int32_t n_seg, int32_t c, b2s_stream_t stream);
int b2s_global_avg_pool_fp(const float* feats, const int32_t* offsets, float* out, int32_t n_seg,
                           int32_t c, b2s_stream_t stream);
int b2s_global_avg_pool_bp(float* d_feats, const int32_t* offsets, const float* d_out,
                           int32_t n_seg, int32_t c, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * clusters_voxelization (minsu3d/model/general_model.py:152-193) up to and
 * including `clusters_coords.int()`, as one call: replaces three sec_* reductions, ~12 elementwise
 * torch kernels and two index_selects.  Every float operation of the reference expression sequence
 * keeps its own fp32 rounding, so the voxel coordinates are bit-identical.
 *   clusters_idx  [sumNPoint, 2] (cluster id, point index), int32 or int64 (idx_is_int64)
 *   clusters_offset [nCluster+1] int32 ; coords [N,3] float32 ; rand6 = the two torch.rand(3) draws (device)
 *   out_xyz [sumNPoint, 4] int32 = (cluster id, x, y, z) -- the `batched_xyz` handed to sparse_quantize
 *   cluster_params [nCluster, 8] float32 scratch/out: mean xyz, scale, offset xyz, unused
 * ---------------------------------------------------------------------------------------------- */
int b2s_clusters_voxelize(const void* clusters_idx, int32_t idx_is_int64, const int32_t* clusters_offset,
                          int64_t sum_npoint, int32_t n_cluster, const float* coords, float scale,
                          int32_t spatial_shape, const float* rand6, int32_t* out_xyz,
                          float* cluster_params, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * I1 / I2 -- proposal x instance IoU and mask labels (get_iou.cu:12-37,
 * cal_iou_and_masklabel.cu:14-140).  mask_scores == NULL: all proposal points count.
 * ---------------------------------------------------------------------------------------------- */
int b2s_get_iou(const int32_t* proposals_idx, const int32_t* proposals_offset,
                const int16_t* instance_labels, const int32_t* instance_pointnum,
                const float* mask_scores, float* proposals_iou, int32_t n_instance,
                int32_t n_proposal, b2s_stream_t stream);
int b2s_get_mask_label(const int32_t* proposals_idx, const int32_t* proposals_offset,
                       const int16_t* instance_labels, const int16_t* instance_cls,
                       const float* proposals_iou, int32_t n_instance, int32_t n_proposal,
                       int32_t ignored_label, float iou_thr, uint8_t* mask_label,
                       uint8_t* mask_label_mask, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * instance post-processing on the device -- PointGroup._get_pred_instances /
 * _get_nms_instances (minsu3d/model/pointgroup.py:197-265) and HAIS._get_pred_instances
 * (minsu3d/model/hais.py:210-247) without the dense [nProposal, N] masks and the CPU round trip.
 *   b2s_proposal_sort : keys_sorted [S] uint64 = sorted (point << 32 | proposal) of the pairs with valid[i] != 0
 *                       (valid == NULL: all; HAIS mask-score filter, hais.py:222-223); dropped pairs sort last
 *   b2s_proposal_npoint: npoint [nProposal] = distinct points per proposal (= mask.sum(1))
 *   b2s_proposal_iou  : remap [nProposal] = row of the proposal among the kept ones or -1;
 *                       inter [n_kept, n_kept] int32 = mask @ mask.T, iou = inter / (n_a + n_b - inter) (fp32)
 *   b2s_nms           : greedy suppression in the given order (descending score): pick [<= n], d_count = picks;
 *                       ws >= n bytes
 * ---------------------------------------------------------------------------------------------- */
size_t b2s_proposal_sort_ws_bytes(int64_t S);
int b2s_proposal_sort(const int32_t* prop_idx, const uint8_t* valid, int64_t S, uint64_t* keys_sorted, void* ws,
                      size_t ws_bytes, b2s_stream_t stream);
int b2s_proposal_npoint(const uint64_t* keys_sorted, int64_t S, int32_t n_proposal, int32_t* npoint,
                        b2s_stream_t stream);
int b2s_proposal_iou(const uint64_t* keys_sorted, int64_t S, const int32_t* remap, int32_t n_kept, int32_t* inter,
                     float* iou, b2s_stream_t stream);
int b2s_nms(const float* iou, const int32_t* order, int32_t n, float threshold, int32_t* pick, int32_t* d_count,
            void* ws, size_t ws_bytes, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * SURVEY.md 8(f) #3 -- the loader's train-split sample pipeline on the GPU (general_dataset.py:80-165,
 * util/transform.py:65-98).  Random draws are inputs (host pointers m9 / jitter3 / offset3 / range3 are read at call
 * time).  Elastic distortion and the crop test run in double precision like the reference's numpy / scipy code.
 *   b2s_aug_affine   : out_xyz = xyz @ M (float32), out_rgb = rgb + jitter (rgb may be NULL)
 *   b2s_elastic_blur : noise [3][b0][b1][b2] float32 blurred in place by six separable 3-tap box passes
 *                      (scipy.ndimage.convolve mode="constant"), tmp = scratch of the same size
 *   b2s_elastic_apply: x [n,3] double += mag * trilinear(noise, x) on the grid linspace(-(b-1) gran, (b-1) gran, b)
 *   b2s_crop_test    : out = pc + offset; valid[i] = out in [0, range); d_count[0] = number of valid points
 * ---------------------------------------------------------------------------------------------- */
int b2s_aug_affine(const float* xyz, const float* rgb, int64_t n, const float* m9_host, const float* jitter3_host,
                   float* out_xyz, float* out_rgb, b2s_stream_t stream);
int b2s_elastic_blur(float* noise, float* tmp, int32_t b0, int32_t b1, int32_t b2, b2s_stream_t stream);
int b2s_elastic_apply(double* x, const float* noise, int64_t n, int32_t b0, int32_t b1, int32_t b2, double gran,
                      double mag, b2s_stream_t stream);
int b2s_crop_test(const double* pc, int64_t n, const double* offset3_host, const double* range3_host, double* out,
                  uint8_t* valid, int32_t* d_count, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Semantic cross-entropy of the train step (general_model.py:36-40: F.cross_entropy(scores, labels, ignore_index),
 * mean over the labelled points), one fused pass per direction.  labels int16 (label_bytes 2) or int64 (8).
 * forward: loss [1], n_valid [1] (input of the backward); counter = int32 that is 0 between launches.
 * backward: gscores [n,c] = (softmax - onehot) * gout[0] / n_valid[0], 0 for ignored rows.
 * ---------------------------------------------------------------------------------------------- */
size_t b2s_cross_entropy_ws_bytes(int64_t n);
int b2s_cross_entropy_forward(const float* scores, const void* labels, int32_t label_bytes, int64_t n, int32_t c,
                              int32_t ignore_index, float* loss, float* n_valid, int32_t* counter, void* ws,
                              size_t ws_bytes, b2s_stream_t stream);
int b2s_cross_entropy_backward(const float* scores, const void* labels, int32_t label_bytes, int64_t n, int32_t c,
                               int32_t ignore_index, const float* n_valid, const float* gout, float* gscores,
                               b2s_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H */
