/*
 * p2w.h -- C ABI of libp2w.so, the sm_100a replacement for the neighbourhood-search and
 * message-passing ops on PointsToWood's inference hot path.
 *
 * The reference reaches this path through Python wrappers around the dispatcher ops of
 * torch_cluster / torch_scatter / torch_geometric (un-vendored; call sites below are in
 * /root/reference/pointstowood).  Each entry point names the interface it replaces.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - row-major contiguous arrays, FP32 coordinates [N,3], int64 CSR `ptr` arrays
 *     (ptr[b]..ptr[b+1] = rows of example/tile b; the `batch` vector must be sorted);
 *   - no allocation and no host synchronisation inside: the caller passes outputs and
 *     workspaces; work is enqueued on `stream` (a cudaStream_t) and is graph-capturable;
 *   - returns 0, or a negative P2W_E* code with the text available from p2w_last_error().
 */
#ifndef P2W_H_
#define P2W_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *p2w_stream_t; /* cudaStream_t */

#define P2W_OK 0
#define P2W_EINVAL (-1)   /* bad argument (the upstream TORCH_CHECK / AT_ASSERTM cases) */
#define P2W_ECUDA (-2)    /* CUDA runtime error at launch */
#define P2W_ENOGPU (-3)   /* no sm_100 device */

#define P2W_F32 0         /* element types of the activation arrays that may be FP32 or BF16 */
#define P2W_BF16 1

#define P2W_MAX_K 128     /* upstream knn asserts k <= 100 */

int p2w_version(void);
const char *p2w_last_error(void);
/* Number of SMs / compute capability of the current device (host query). */
int p2w_device_info(int *sm_count_host, int *cc_major_host, int *cc_minor_host);
/* Number of kernels libp2w has enqueued since it was loaded (all streams, this process). */
long long p2w_launch_count(void);

/* ---- K1: exact k nearest neighbours -------------------------------------------------
 * Replaces torch_cluster::knn (src/model.py:120 SA2/SA3 k=32; :149 knn_interpolate k=2).
 * nbr[q*k+e] = global x index of the e-th nearest source of query q inside its own tile,
 * ascending by (FP32 squared distance, index); -1 where the tile has fewer than k
 * sources.  d2 (optional, may be NULL) receives the matching squared distances (1e10 pad). */
int p2w_knn(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
            int32_t num_tiles, int64_t nx, int64_t ny, int32_t k,
            int32_t *nbr, float *d2, p2w_stream_t stream);

/* ---- K2: radius search --------------------------------------------------------------
 * Replaces torch_cluster::radius, CUDA semantics (src/model.py:118, r=0.08, max 32):
 * the first max_nbr sources in ascending index order with d2 < (float)(r*r).
 * nbr [ny,max_nbr] -1 padded, cnt [ny]. */
int p2w_radius(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
               int32_t num_tiles, int64_t nx, int64_t ny, double r, int32_t max_nbr,
               int32_t *nbr, int32_t *cnt, p2w_stream_t stream);

/* ---- K1 / K2 with a per-tile cell list ----------------------------------------------------
 * Same contracts and bit-identical results as p2w_knn / p2w_radius, but every query examines only
 * the cells of its tile's uniform grid that can hold a result (growing Chebyshev shells, exact
 * stopping bound), i.e. a few hundred candidates instead of the whole tile.  The grid is built per
 * call from the sources (bounding box + occupancy pyramid per tile, one radix sort) in `ws`
 * (p2w_grid_search_ws_bytes(nx, ny, num_tiles) bytes, 16-byte aligned).  Pays off from a few hundred
 * sources per tile; the brute-force sweep remains the better choice for tiny tiles.
 * k >= 5 runs one thread per query over a per-thread heap in shared memory, the queries binned by cell
 * of the source grid (a counting sort in `ws`, hence ny in the workspace size); k <= 4 keeps the keys in
 * registers.  p2w_grid_search_pair_evals returns (in *counter_host) the DEVICE address inside `ws` of a
 * uint64 that the last search on that workspace left behind: the number of distance evaluations it made
 * (bench.py's pair-evaluations / s figure); read it after synchronising the stream. */
size_t p2w_grid_search_ws_bytes(int64_t nx, int64_t ny, int32_t num_tiles);
int p2w_grid_search_pair_evals(const void *ws, int64_t nx, int64_t ny, int32_t num_tiles,
                               const unsigned long long **counter_host);
int p2w_knn_grid(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
                 int32_t num_tiles, int64_t nx, int64_t ny, int32_t k,
                 int32_t *nbr, float *d2, void *ws, size_t ws_bytes, p2w_stream_t stream);
/* Same with the cell size given by the caller (cell_size > 0, enlarged until the cell table fits 4 cells
 * per source, at most 1024 cells per axis): plot-wide searches such as the spatial vote, where one
 * "tile" holds millions of points and the occupancy pyramid's 64 cells per axis are too coarse. */
#define P2W_KNN_UNORDERED 1   /* flags: entry 0 of a row is the k-th (farthest) neighbour, the others follow in no
                               * particular order (same SET as the ordered table; -1 in entry 0 when the tile has
                               * fewer than k sources).  For consumers that do not need the order -- the spatial
                               * vote -- it saves the final sort.  May be ignored (an ordered table is returned
                               * with the k-th neighbour LAST) when k <= 4. */
int p2w_knn_grid_ex(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
                    int32_t num_tiles, int64_t nx, int64_t ny, int32_t k, float cell_size, int32_t flags,
                    int32_t *nbr, float *d2, void *ws, size_t ws_bytes, p2w_stream_t stream);
int p2w_radius_grid(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
                    int32_t num_tiles, int64_t nx, int64_t ny, double r, int32_t max_nbr,
                    int32_t *nbr, int32_t *cnt, void *ws, size_t ws_bytes, p2w_stream_t stream);

/* Compaction of a -1 padded [ny,k] table into upstream's [2,E] int64 edge list
 * (row 0 = query index, row 1 = source index, query-major).  edge_offset [ny+1] is a
 * caller workspace that receives the exclusive prefix sum of the per-query counts;
 * E = edge_offset[ny].  Two calls: p2w_table_count fills edge_offset, then the caller
 * reads E, allocates `edges` [2,E] and calls p2w_table_to_edges. */
int p2w_table_count(const int32_t *nbr, int64_t ny, int32_t k, int64_t *edge_offset,
                    p2w_stream_t stream);
int p2w_table_to_edges(const int32_t *nbr, int64_t ny, int32_t k, const int64_t *edge_offset,
                       int64_t num_edges, int64_t *edges, p2w_stream_t stream);

/* ---- K3: farthest point sampling ----------------------------------------------------
 * Replaces torch_cluster::fps (not called by the reference; BASELINE config 3).
 * out_ptr [num_tiles+1] (device) = cumulative ceil(ratio * n_b), computed by the caller;
 * start = first point of each tile (random_start=False); ties -> lowest index. */
int p2w_fps(const float *src, const int64_t *ptr, const int64_t *out_ptr, int32_t num_tiles,
            int64_t n, float *dist_ws /* [n] */, int64_t *out, p2w_stream_t stream);

/* ---- K4: voxel grid ids and voxel representatives ----------------------------------
 * p2w_colminmax: per-column min / max of pos [n,dim] (the start/end defaults of
 * torch_cluster::grid).  mn/mx [dim].
 * p2w_grid replaces torch_cluster::grid (src/model.py:104, src/preprocessing.py:33,58):
 *   id = sum_d (int64)((pos[d]-start[d])/size[d]) * prod_{d'<d}((int64)((end-start)/size)+1).
 * If batch != NULL an extra trailing column (float)batch[i] with size 1, start
 * batch_start and end batch_end is appended (torch_geometric.nn.voxel_grid);
 * start/end then hold dim+1 entries. */
int p2w_colminmax(const float *pos, int64_t n, int32_t dim, int32_t ld, float *mn, float *mx,
                  p2w_stream_t stream);
int p2w_grid(const float *pos, int64_t n, int32_t dim, int32_t ld, const int64_t *batch,
             const float *size, const float *start, const float *end, int64_t *ids,
             p2w_stream_t stream);

/* Voxel keys for SEVERAL reference batches in one launch ("super-batch").  Tiles are the CSR
 * tile_ptr [num_tiles+1] over points; group_ptr [num_groups+1] (device, over tiles) lists the
 * reference batches (src/predicter.py:177-180, batch_size tiles each).  Each group is voxelised on
 * its own grid, start/end = column min/max over that group's points, exactly what
 * voxel_grid(pos, size, batch) does for one batch (src/model.py:104).
 * keys[i] = (tile << spatial_bits) | spatial_id: ascending key order == the reference's ascending
 * batch-major cluster id inside every group.  group_min/group_max: [num_groups,3] workspace/outputs;
 * *overflow (device int32) is set to 1 when an id needs more than spatial_bits bits. */
int p2w_voxel_keys_grouped(const float *pos, int64_t n, int32_t ld, const int64_t *tile_ptr,
                           int32_t num_tiles, const int64_t *group_ptr, int32_t num_groups,
                           float size, int32_t spatial_bits, float *group_min, float *group_max,
                           uint64_t *keys, int32_t *overflow, p2w_stream_t stream);

/* Stable LSD radix sort of (key, value) pairs on key bits [0, key_bits).
 * ws: p2w_sort_ws_bytes(n) bytes.  Results land in keys_out / vals_out. */
size_t p2w_sort_ws_bytes(int64_t n);
int p2w_sort_pairs(const uint64_t *keys_in, const int32_t *vals_in, uint64_t *keys_out,
                   int32_t *vals_out, int64_t n, int32_t key_bits, void *ws, p2w_stream_t stream);

size_t p2w_unique_ws_bytes(int64_t n);
/* torch_geometric consecutive_cluster (src/model.py:105) on SORTED keys with their
 * original indices (ascending inside equal keys, as the stable sort leaves them):
 * perm[u] = highest original index of the u-th distinct key, inverse[orig] = u,
 * seg_start[u] = first sorted position of the u-th key and seg_start[num_unique] = n (the CSR
 * of src/preprocessing.py:59-63's per-voxel member lists); perm / inverse / seg_start may be
 * NULL.  *num_unique (device int64) = number of distinct keys.  ws: p2w_unique_ws_bytes(n). */
int p2w_unique_last(const uint64_t *sorted_keys, const int32_t *sorted_idx, int64_t n,
                    int64_t *perm, int64_t *inverse, int64_t *seg_start, int64_t *num_unique,
                    void *ws, p2w_stream_t stream);

/* ---- K5: fused PointNetConv: gather -> per-edge MLP -> max --------------------------
 * Replaces MessagePassing.propagate(aggr='max') around PointNetConv.message
 * (src/pointnet.py:108,116-132) for local_nn = Lin(C+4,H) ReLU Lin(H,Co) ReLU BN(Co):
 *   rel = pos_src[j,:3] - pos_tgt[i,:3];  m_i = max_j |rel|;
 *   msg = [x[j], rel/(m_i+1e-8), pos_src[j,3]];  out[i] = max_j BN(ReLU(W2 ReLU(W1 msg+b1)+b2))
 * over the valid entries j = nbr[i, :]; targets without neighbours get 0.
 * x [n_src,C] fp32; pos_src [n_src,4], pos_tgt [n_tgt,4] (xyz already divided by sf,
 * reflectance in column 3); w1 [H,C+4], w2 [Co,H] row-major (torch Linear layout);
 * bn_scale/bn_shift [Co] = folded eval-mode BatchNorm; out [n_tgt,Co] fp32.
 * mode: 0 = FP32 SIMT (parity mode, 1e-3), 1 = BF16 operands on tcgen05 tensor cores with
 * FP32 accumulation in TMEM (1e-2).  The E x C' edge tensor is never written to HBM. */
#define P2W_CONV_FP32 0
#define P2W_CONV_BF16_TC 1
int p2w_pointnet_conv_max(const float *x, const float *pos_src, const float *pos_tgt,
                          const int32_t *nbr, int64_t n_src, int64_t n_tgt, int32_t k,
                          int32_t c_in, int32_t hidden, int32_t c_out,
                          const float *w1, const float *b1, const float *w2, const float *b2,
                          const float *bn_scale, const float *bn_shift, float *out,
                          int32_t mode, void *ws, size_t ws_bytes, p2w_stream_t stream);
/* Extended form: feature rows in / out may be BF16 in the tensor-core mode (x_dtype, out_dtype =
 * P2W_F32 | P2W_BF16), and flags & P2W_CONV_WS_PACKED says `ws` still holds the weights re-laid-out
 * by an earlier call with the same (w1, b1, w2, b2, bn) and mode, so the packing kernels are skipped
 * (a model packs once).  tgt_index (may be NULL, int64 [n_tgt], tensor-core mode): target t sits at
 * pos_tgt[tgt_index[t]] -- pass pos_src as pos_tgt and the sampled indices instead of gathering pos[idx]. */
#define P2W_CONV_WS_PACKED 1
int p2w_pointnet_conv_max_ex(const void *x, int32_t x_dtype, const float *pos_src, const float *pos_tgt,
                             const int32_t *nbr, int64_t n_src, int64_t n_tgt, int32_t k,
                             int32_t c_in, int32_t hidden, int32_t c_out,
                             const float *w1, const float *b1, const float *w2, const float *b2,
                             const float *bn_scale, const float *bn_shift, void *out, int32_t out_dtype,
                             int32_t mode, void *ws, size_t ws_bytes, int32_t flags, const int64_t *tgt_index,
                             p2w_stream_t stream);
size_t p2w_pointnet_conv_ws_bytes(int32_t c_in, int32_t hidden, int32_t c_out, int32_t mode);

/* ---- knn_interpolate (src/model.py:149, torch_geometric.nn.knn_interpolate) ---------
 * out[q,:] = sum_e w_e x[nbr[q,e],:] / sum_e w_e,  w_e = 1/max(|pos_x[nbr]-pos_y[q]|^2, 1e-16). */
int p2w_knn_interpolate(const float *x, const float *pos_x, const float *pos_y,
                        const int32_t *nbr, int64_t ny, int32_t k, int32_t c, int32_t ld_out,
                        float *out, p2w_stream_t stream);

/* Same with FP32 or BF16 feature rows in and out (x_dtype / out_dtype = P2W_F32 | P2W_BF16); positions
 * and weights stay FP32. */
int p2w_knn_interpolate_ex(const void *x, int32_t x_dtype, const float *pos_x, const float *pos_y,
                           const int32_t *nbr, int64_t ny, int32_t k, int32_t c, int32_t ld_out,
                           void *out, int32_t out_dtype, p2w_stream_t stream);

/* FPModule.forward up to its MLP (src/model.py:149-151): out[q] = [knn_interpolate(x)[q], skip[q]], the
 * interpolation and torch.cat([x, x_skip], dim=1) in one pass over 16-byte channel groups.
 * c, c_skip and ld_out (>= c + c_skip) are multiples of 8; any mix of FP32 / BF16 rows. */
int p2w_knn_interpolate_cat(const void *x, int32_t x_dtype, const float *pos_x, const float *pos_y,
                            const int32_t *nbr, int64_t ny, int32_t k, int32_t c,
                            const void *skip, int32_t skip_dtype, int32_t c_skip, int32_t ld_out,
                            void *out, int32_t out_dtype, p2w_stream_t stream);

/* The first Linear of an FPModule applied BEFORE the interpolation (src/model.py:149-152): knn_interpolate is
 * linear with weights that sum to one, so relu([interp(x), x_skip] W^T + b) = relu(interp(x Wc^T) + (x_skip Ws^T + b)).
 * out[q, :] = act(sum_e w_e y[nbr[q,e], :] / sum_e w_e + z[q, :]) with y = x Wc^T over the COARSE rows (y_dtype) and
 * z = x_skip Ws^T + b over the fine rows (out_dtype; out may alias z); relu != 0 applies max(., 0).  c % 8 == 0. */
int p2w_knn_interpolate_add(const void *y, int32_t y_dtype, const float *pos_x, const float *pos_y,
                            const int32_t *nbr, int64_t ny, int32_t k, int32_t c, const void *z, void *out,
                            int32_t out_dtype, int32_t relu, p2w_stream_t stream);

/* ---- dense-block epilogues (src/model.py:18-85, InvertedResidualBlock in eval mode) ------------
 * What remains between two k=1 convolutions once every BatchNorm that follows a convolution is
 * folded into its weights: y = relu(x*s1 + t1), and if s2 != NULL y = relu(y*s2 + t2), per channel,
 * on [n, c] activations (c % 8 == 0; x may alias y). */
int p2w_affine_relu(const void *x, void *y, int64_t n, int32_t c, const float *s1, const float *t1,
                    const float *s2, const float *t2, int32_t dtype, p2w_stream_t stream);

/* The expand convolution of InvertedResidualBlock with its epilogue chain on tcgen05 (src/model.py:46-85, eval mode, BN folded):
 * out[n, co] = relu(relu(x[n, :] . w[co, :] + bias[co]) * a[co] + c[co]) (a == c == NULL: the first ReLU only); x [n, k] and
 * out [n, c_out] BF16 row-major, w [c_out, k] / bias / a / c FP32.  One pass instead of a library GEMM plus p2w_affine_relu over
 * [n, c_out].  k a multiple of 64 in [64, 512].  ws: p2w_dense_expand_ws_bytes(k, c_out) bytes; flags bit 0: ws still holds the
 * weights packed by an earlier call. */
size_t p2w_dense_expand_ws_bytes(int32_t k, int32_t c_out);
int p2w_dense_expand(const void *x, int64_t n, int32_t k, int32_t c_out, const float *w, const float *bias,
                     const float *a, const float *c, void *out, void *ws, size_t ws_bytes, int32_t flags,
                     p2w_stream_t stream);

/* out[r] = dot(x[r, :], w) + bias over [n, c] FP32 / BF16 rows (c % 8 == 0, c <= 1024): the 1-channel head
 * conv2 (src/model.py:243) as one streaming pass instead of a GEMM with one output column. */
int p2w_rowdot(const void *x, int32_t dtype, int64_t n, int32_t c, const float *w, float bias, float *out,
               p2w_stream_t stream);

/* out[i] = relu(a[i] + b[i]) over n FP32 / BF16 elements (n a multiple of 4 / 8; out may alias a): the shortcut
 * add + ReLU that closes InvertedResidualBlock (src/model.py:84) in one pass. */
int p2w_add_relu(const void *a, const void *b, void *out, int64_t n, int32_t dtype, p2w_stream_t stream);

/* ---- segment / scatter reductions ---------------------------------------------------
 * p2w_segment_max: torch_geometric global_max_pool (src/model.py:136) for a sorted batch
 * given as ptr: out[b,:] = max over rows ptr[b]..ptr[b+1] (0 for empty segments).
 * p2w_scatter_minmax: torch_scatter::scatter_max / scatter_min along dim 0
 * (src/pointnet.py:122, src/preprocessing.py:49): out [dim_size,c] (0 for empty slots),
 * arg [dim_size,c] int64 (n for empty slots; lowest row on ties).  is_max: 1 max, 0 min. */
int p2w_segment_max(const float *x, const int64_t *ptr, int32_t num_segments, int32_t c,
                    float *out, p2w_stream_t stream);
/* The same over FP32 / BF16 rows (dtype = P2W_F32 | P2W_BF16, c even) with an optional per-channel affine
 * applied to every element first: out[b,ch] = max_r (x[r,ch] * scale[ch] + shift[ch]) -- GlobalSAModule's
 * last BatchNorm + global_max_pool (src/model.py:134-136) in one pass.  scale / shift may be NULL. */
int p2w_segment_max_ex(const void *x, int32_t dtype, const int64_t *ptr, int32_t num_segments, int32_t c,
                       const float *scale, const float *shift, float *out, p2w_stream_t stream);
int p2w_scatter_minmax(const float *src, const int64_t *index, int64_t n, int32_t c,
                       int64_t dim_size, int32_t is_max, float *out, int64_t *arg,
                       p2w_stream_t stream);

/* ---- K7 / K8 and SA glue ------------------------------------------------------------
 * p2w_sa_prepare (src/model.py:109,122,124): pos4[i] = (pos[i]/sf[tile], refl[i]) and
 * pos_back[i] = (pos[i]/sf[tile])*sf[tile] (the reference's in-place round trip, which
 * perturbs coordinates by an ulp and feeds the next level).  tile id from ptr.
 * p2w_pack (src/predicter.py:78-94 + PyG collate): gathers rows `index` of cloud [n,ld]
 * (x,y,z,reflectance in columns 0..3) tile by tile, subtracts the per-tile mean
 * (local_shift [B,3]) and returns sf[b] = max |pos|; pos [m,3], refl [m], batch [m].
 * p2w_writeback (src/predicter.py:199-214): prob = sigmoid(nan_to_num(logit)),
 * pred = prob >= is_wood, xyz = pos + local_shift[tile]; rows (x,y,z,pred,prob) as
 * float64 [m,5] (out64) and/or compact prob [m] / pred [m] / xyz32 [m,3] (the un-shifted coordinates
 * rounded to FP32, what the spatial vote searches); every output may be NULL. */
int p2w_sa_prepare(const float *pos, int32_t ld_pos, const float *refl, const int64_t *ptr,
                   const float *sf, int32_t num_tiles, int64_t n, float *pos4, float *pos_back,
                   p2w_stream_t stream);
int p2w_pack(const float *cloud, int32_t ld, const int64_t *index, const int64_t *ptr,
             int32_t num_tiles, int64_t m, float *pos, float *refl, int64_t *batch,
             float *local_shift, float *sf, p2w_stream_t stream);
int p2w_writeback(const float *logits, const float *pos, const int64_t *ptr,
                  const float *local_shift, int32_t num_tiles, int64_t m, float is_wood,
                  double *out64, float *prob, uint8_t *pred, float *xyz32, p2w_stream_t stream);

/* ---- spatial vote (src/predicter.py:113-142, PointCloudClassifier.compute_labels) -------------
 * For every original point q and its k nearest CLASSIFIED points nbr[q, :] (from p2w_knn_grid_ex
 * over the whole plot; -1 = missing): pwood[q] = median of prob[nbr] (np.median, float64);
 * label[q] = (any_wood == 1) ? (sum of prob over pred == 1) > (sum over pred == 0)
 *                            : any(pred[nbr] > any_wood).   k <= 128. */
int p2w_spatial_vote(const int32_t *nbr, int64_t n, int32_t k, const float *prob, const uint8_t *pred,
                     float any_wood, uint8_t *label, double *pwood, p2w_stream_t stream);

/* ---- K6: tiling front end (src/preprocessing.py:18-64, 116-120) ---------------------------
 * p2w_ground_normalize (gpu_ground, :37-53): 5 m XY cells with edges x_min + 5 i (bucketize:
 * cell = number of edges strictly below the coordinate), per-cell min z, n_z = z - min.
 * cell_min is a caller workspace of nbx*nby floats; mnmx = {x_min, y_min} on the device.
 * p2w_reflectance_keys + p2w_sort_pairs(32 bits) + p2w_reflectance_normalize
 * (quantile_normalize_reflectance, :18-30): rank (stable) -> q = (rank+1)/(N+1) clamped to
 * [1e-7, 1-1e-7] -> erfinv(2q-1)*sqrt(2); then out = 2 (v-min)/(max-min) - 1.
 * p2w_assemble5: feat [n,5] = (x, y, z, reflectance_scaled, n_z), the array the reference
 * voxelises with all five columns (:52,58).
 * p2w_sampling_keys (torch.multinomial without replacement, :116-118, which draws from torch's
 * global generator in the reference): Efraimidis-Spirakis keys -log(u_i) / w_i (float64, written as
 * their order-preserving bit pattern) with w = feat[i, col] - refl_min + 1e-8 and u a counter-based hash
 * of (seed, point index) in (0,1]; ascending key order = the order in which sequential weighted
 * sampling without replacement draws the members, so the first max_pts of a tile are the sample.
 * members [m] are rows of feat; global_index (may be NULL) maps a row of feat to the point index that
 * is hashed (a rank of a sharded plot holds a subset of the rows).
 * p2w_replacement_picks (torch.randint, :120, clouds without reflectance): picks[t, s] =
 * members[seg[t] + hash(seed, voxel_id[t], s) mod (seg[t+1] - seg[t])], max_pts draws WITH
 * replacement for each of num_tiles oversized tiles.
 * p2w_ground_min / p2w_ground_apply are the two halves of p2w_ground_normalize (a sharded plot
 * min-reduces cell_min over the ranks in between). */
int p2w_ground_normalize(const float *cloud, int32_t ld, int64_t n, const float *mn_xy,
                         float cell, int32_t nbx, int32_t nby, float *cell_min, float *n_z,
                         p2w_stream_t stream);
int p2w_reflectance_keys(const float *cloud, int32_t ld, int32_t col, int64_t n, uint64_t *keys,
                         p2w_stream_t stream);
int p2w_reflectance_normalize(const int32_t *sorted_idx, int64_t n, float *v, float *mnmx_ws,
                              float *out, p2w_stream_t stream);
/* The two halves of p2w_reflectance_normalize for a plot ranked in key ranges over several GPUs: the values
 * erfinv(2q-1)*sqrt(2) of the sorted positions p -> global rank rank0 + p of n_total (v[sorted_idx[p]], and
 * their min / max in mnmx[0..1]), then -- after the ranks have min / max-reduced mnmx -- the affine map. */
int p2w_reflectance_values(const int32_t *sorted_idx, int64_t n, int64_t rank0, int64_t n_total,
                           float *v, float *mnmx, p2w_stream_t stream);
int p2w_reflectance_scale(const float *v, int64_t n, const float *mnmx, float *out, p2w_stream_t stream);
int p2w_assemble5(const float *cloud, int32_t ld, const float *refl, const float *n_z, int64_t n,
                  float *feat, p2w_stream_t stream);
int p2w_sampling_keys(const float *feat, int32_t ld, int32_t col, const int32_t *members,
                      const int32_t *global_index, int64_t m, float refl_min, uint32_t seed,
                      uint64_t *keys, p2w_stream_t stream);
int p2w_replacement_picks(const int32_t *members, const int64_t *seg, const int64_t *voxel_id,
                          int32_t num_tiles, int32_t max_pts, uint64_t seed, int32_t *picks,
                          p2w_stream_t stream);
int p2w_ground_min(const float *cloud, int32_t ld, int64_t n, const float *mn_xy, float cell,
                   int32_t nbx, int32_t nby, float *cell_min, p2w_stream_t stream);
int p2w_ground_apply(const float *cloud, int32_t ld, int64_t n, const float *mn_xy, float cell,
                     int32_t nbx, int32_t nby, const float *cell_min, float *n_z, p2w_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* P2W_H_ */
