/*
 * cpd_b200.h -- C ABI of libcpd_b200.so: the B200-native (sm_100a) implementation of
 * the CPD detection hot path (voxelizer -> sparse-3D-conv backbone -> BEV head -> NMS).
 *
 * Conventions:
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless the
 *     parameter name ends in _host; fp32 features, int32 indices, row-major;
 *   - the caller owns every buffer, including workspaces (query *_workspace_bytes);
 *     the library never allocates device memory and never synchronises the device;
 *   - every entry takes the CUDA stream to launch on and is graph-capturable;
 *   - every entry returns int32 status: 0 = ok, <0 = cpd_status; nothing ever calls
 *     exit() or throws across the boundary (the reference does: iou3d_nms.cpp:14-38);
 *   - data-dependent output sizes are returned through device counters that the host
 *     reads when it needs them (two-phase calls).
 *
 * Each entry cites the reference interface it replaces (paths relative to the
 * hailanyi/CPD tree).  INTEGRATION.md shows the Python-side binding.
 */
#ifndef CPD_B200_H_
#define CPD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CPD_API __attribute__((visibility("default")))
#else
#define CPD_API
#endif

typedef void *cpd_stream_t; /* cudaStream_t */

enum cpd_status {
    CPD_OK = 0,
    CPD_ERR_BAD_ARG = -1,
    CPD_ERR_MISALIGNED = -2,
    CPD_ERR_WORKSPACE = -3, /* workspace too small */
    CPD_ERR_UNSUPPORTED = -4,
    CPD_ERR_CUDA = -100 /* CUDA error; cpd_last_error_string() has the text */
};

#define CPD_MAX_BATCH 64

CPD_API int32_t cpd_version(void);
CPD_API const char *cpd_last_error_string(void);
/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
CPD_API int64_t cpd_launch_count(void);

/* ---------------------------------------------------------------------------------
 * Voxelizer.  Replaces spconv.utils.Point2VoxelCPU3d.point_to_voxel as called from
 * cpd/datasets/processor/data_processor.py:35-41,53-58 (VoxelGeneratorWrapper) for a
 * whole batch at once, including collate_batch's batch-index column
 * (cpd/datasets/dataset.py:262-266) and, optionally, MeanVFE
 * (cpd/models/backbones_3d/vfe/mean_vfe.py:41-50).
 *
 * points: (n_points, c) fp32, frames concatenated; frame f owns rows
 * [frame_offsets_host[f], frame_offsets_host[f+1]).  Voxel rows come out frame-major,
 * inside a frame in order of first appearance, each keeping its first max_pts points
 * in input order -- bit-identical to the sequential reference.
 * counts: (batch+1,) int32 device: voxels kept per frame, [batch] = total rows.
 * mean may be NULL.  voxels may be NULL (mean-only mode).
 * --------------------------------------------------------------------------------- */
CPD_API size_t cpd_voxelize_workspace_bytes(int64_t n_points, int32_t batch, int32_t max_pts, int64_t cap_rows);
CPD_API int32_t cpd_voxelize(const float *points, int64_t n_points, int32_t c,
                     const int64_t *frame_offsets_host, int32_t batch,
                     const float *range6_host, const float *vsize3_host,
                     int32_t max_pts, int64_t max_voxels, int64_t cap_rows,
                     float *voxels, int32_t *coords_bzyx, int32_t *num_points, float *mean,
                     int32_t *counts, void *ws, size_t ws_bytes, cpd_stream_t stream);

/* Host (CPU) entry with the same semantics for ONE frame, all pointers HOST pointers: the reference calls the
 * voxelizer inside Dataset.__getitem__, i.e. in forked DataLoader workers (data_processor.py:133-144) where CUDA must not be
 * touched.  voxels (cap, max_pts, c), coords_zyx (cap, 3), num_points (cap,), cap >= min(n, max_voxels).  Returns the number
 * of voxels (>= 0) or a cpd_status (< 0).  This is the shim's spconv.utils.Point2VoxelCPU3d; the hot path uses cpd_voxelize. */
CPD_API int64_t cpd_voxelize_cpu(const float *points, int64_t n_points, int32_t c, const float *range6_host, const float *vsize3_host,
                         int32_t max_pts, int64_t max_voxels, float *voxels, int32_t *coords_zyx, int32_t *num_points);

/* ---------------------------------------------------------------------------------
 * Rulebooks.  Replace the indice-pair generation hidden inside spconv.SubMConv3d /
 * spconv.SparseConv3d (call sites cpd/models/backbones_3d/spconv_backbone.py:17,20-21,
 * 108-115,189-190,415,451-452), cached per indice_key like upstream.
 *
 * Layout is output-stationary: nbr[(row, tap)] = input row feeding `row` through
 * tap = (kz*KH+ky)*KW+kx, or -1.  coords: (m,4) int32 [b,z,y,x].
 * --------------------------------------------------------------------------------- */
CPD_API size_t cpd_coord_hash_bytes(int64_t m);
CPD_API int32_t cpd_coord_hash_build(const int32_t *coords, int64_t m, const int32_t *shape3_host,
                             int32_t batch, void *hash, size_t hash_bytes, cpd_stream_t stream);

/* SubMConv3d: output sites == input sites; input site = site + tap - ksize/2. */
CPD_API int32_t cpd_rulebook_subm(const int32_t *coords, int64_t m, const int32_t *shape3_host, int32_t batch,
                          const int32_t *ksize3_host, const void *hash, size_t hash_bytes,
                          int32_t *nbr, cpd_stream_t stream);

/* SparseConv3d, phase 1: active output set, rows in ascending linear key
 * ((b*D+z)*H+y)*W+x.  out_shape3_host is written on return (host arithmetic);
 * n_out: int32 device counter (rows beyond cap_out are dropped and n_out still
 * reports the true count so the caller can detect overflow). */
CPD_API size_t cpd_rulebook_strided_workspace_bytes(const int32_t *shape3_host, int32_t batch,
                                            const int32_t *ksize3_host, const int32_t *stride3_host,
                                            const int32_t *pad3_host);
CPD_API int32_t cpd_rulebook_strided_outputs(const int32_t *coords, int64_t m, const int32_t *shape3_host,
                                     int32_t batch, const int32_t *ksize3_host,
                                     const int32_t *stride3_host, const int32_t *pad3_host,
                                     int32_t *out_shape3_host, int64_t cap_out, int32_t *out_coords,
                                     int32_t *n_out, void *ws, size_t ws_bytes, cpd_stream_t stream);
/* phase 2: nbr_fwd (m_out, K): input row feeding output o through tap k;
 *          nbr_bwd (m_in, K): output row fed by input i through tap k (may be NULL). */
CPD_API int32_t cpd_rulebook_strided_tables(const int32_t *in_coords, int64_t m_in, const int32_t *in_shape3_host,
                                    const void *in_hash, size_t in_hash_bytes,
                                    const int32_t *out_coords, int64_t m_out, const int32_t *out_shape3_host,
                                    const void *out_hash, size_t out_hash_bytes, int32_t batch,
                                    const int32_t *ksize3_host, const int32_t *stride3_host,
                                    const int32_t *pad3_host, int32_t *nbr_fwd, int32_t *nbr_bwd,
                                    cpd_stream_t stream);

/* ---------------------------------------------------------------------------------
 * Gather-GEMM ("implicit GEMM over a neighbour table"): the arithmetic of
 * spconv.SubMConv3d / SparseConv3d forward and input-gradient, and -- with a table
 * generated on the fly from the image geometry -- of the dense BEV convolutions.
 *
 *   y[o, :] = epilogue( sum_k  W[:, k, :] . x[nbr[o, k], :] )      nbr == -1 -> zero row
 *
 * w: (cout, K, cin) fp32 = spconv 2.x layout (cout, kz, ky, kx, cin)
 *    (cpd/models/detectors/detector3d_template.py:399-408).
 * epilogue: + bias[cout] (NULL ok); * scale[cout] + shift[cout] (NULL ok: folded
 * eval-mode BatchNorm, spconv_backbone.py:410); + residual[o,:] (NULL ok:
 * SparseBasicBlock, spconv_backbone.py:100-136); ReLU if relu != 0.
 * stats (NULL ok): (2, cout) fp32, overwritten with the per-channel sum and sum of
 * squares of the pre-affine output (training-mode BatchNorm statistics).
 * --------------------------------------------------------------------------------- */
enum cpd_gemm_algo { CPD_ALGO_AUTO = 0, CPD_ALGO_SIMT = 1, CPD_ALGO_TCGEN05 = 2 };

/* Split-row image, the operand format of the tcgen05 kernels: xs (m, 2, c) bf16, row i = [hi(c) | lo(c)],
 * hi = RN_bf16(x), lo = RN_bf16(x - hi); same byte size and row pitch (4 c bytes) as x.  c % 8 == 0.
 * cpd_gather_gemm / cpd_gather_wgrad build the images they need in their workspace unless the caller
 * passes one it already has (x_split / dy_split != NULL): forward and weight-gradient share the image
 * of x, input-gradient and weight-gradient share the image of dy. */
/* colsum (NULL ok; c a power of two <= 2048): overwritten with the per-channel column sums of x, accumulated while the
 * rows stream through -- with x = dy this is the bias gradient, for free. */
CPD_API int32_t cpd_split_rows(const float *x, int64_t m, int32_t c, void *xs, float *colsum, cpd_stream_t stream);

/* Per 128-row tile of a neighbour table (m, K <= 32): bit k of masks[tile] = some row of the tile has a
 * neighbour at tap k.  masks: ceil(m / 128) uint32.  Computed once per rulebook (it only depends on the
 * table) and passed to cpd_gather_gemm, whose tensor-core kernel then skips the k-blocks of absent taps. */
CPD_API int32_t cpd_tile_tap_masks(const int32_t *nbr, int64_t m, int32_t K, uint32_t *masks, cpd_stream_t stream);

/* Row sort key for visiting a neighbour table in tap-pattern order: keys[row] bit j = the row has a neighbour at one of the taps
 * [j * taps_per_block, (j + 1) * taps_per_block) -- i.e. the k-blocks of the tensor-core kernel the row needs
 * (taps_per_block = 64 / cin for cin < 64, else 1).  Rows sorted by key share their k-blocks, so a 128-row tile of the
 * sorted table skips (through cpd_tile_tap_masks) every k-block none of its rows uses; results go back to their rows
 * through cpd_gather_gemm's out_rows.  ceil(K / taps_per_block) <= 31. */
CPD_API int32_t cpd_tap_block_keys(const int32_t *nbr, int64_t m, int32_t K, int32_t taps_per_block, int32_t *keys, cpd_stream_t stream);

/* The sorted table in one pass: out_nbr[r, :] = nbr[perm[r], :], out_rows[r] = (int32) perm[r], tile_masks[t] (NULL ok) as
 * cpd_tile_tap_masks(out_nbr).  perm: int64 permutation of [0, m) (the argsort of cpd_tap_block_keys).  K <= 32. */
CPD_API int32_t cpd_table_permute(const int32_t *nbr, int64_t m, int32_t K, const int64_t *perm, int32_t *out_nbr,
                                  int32_t *out_rows, uint32_t *tile_masks, cpd_stream_t stream);

/* nbr (m, K) -> nbr_t (K, m): the tap-major table cpd_gather_wgrad reads (tap_major = 1).  K <= 64. */
CPD_API int32_t cpd_table_transpose(const int32_t *nbr, int64_t m, int32_t K, int32_t *nbr_t, cpd_stream_t stream);

/* x_split (NULL ok): split-row image of x (cpd_split_rows); have_x_split tells the workspace query
 * whether the call will pass one.  tile_masks (NULL ok): cpd_tile_tap_masks(nbr).
 * out_rows (NULL ok, tensor-core kernel only): a permutation of [0, m_out); row r of the table is written to
 * y[out_rows[r]] -- for tables visited in a regrouped order (strided input-gradient grouped by tap pattern). */
CPD_API int32_t cpd_gather_gemm(const float *x, const void *x_split, int64_t m_in, int32_t cin, const float *w,
                        int32_t K, int32_t cout, const int32_t *nbr, const uint32_t *tile_masks,
                        const int32_t *out_rows, int64_t m_out, const float *bias,
                        const float *scale, const float *shift, const float *residual, int32_t relu,
                        float *stats, float *y, int32_t algo, void *ws, size_t ws_bytes, cpd_stream_t stream);
CPD_API size_t cpd_gather_gemm_workspace_bytes(int64_t m_in, int64_t m_out, int32_t cin, int32_t K, int32_t cout,
                                       int32_t algo, int32_t have_x_split);

/* Weight-gradient: dw[co, k, ci] = sum_o dy[o, co] * x[nbr[o, k], ci]; dw is overwritten.
 * dbias (NULL ok): dbias[co] = sum_o dy[o, co].  nbr_tap_major != 0: nbr is the transposed
 * (K, m_out) table (coalesced per-tap scans; what the tcgen05 kernel prefers).
 * x_split / dy_split (NULL ok): split-row images of x / dy. */
CPD_API int32_t cpd_gather_wgrad(const float *x, const void *x_split, int64_t m_in, int32_t cin, const float *dy,
                         const void *dy_split, int64_t m_out, int32_t cout, const int32_t *nbr,
                         int32_t nbr_tap_major, int32_t K, float *dw, float *dbias, int32_t algo, void *ws,
                         size_t ws_bytes, cpd_stream_t stream);
CPD_API size_t cpd_gather_wgrad_workspace_bytes(int64_t m_in, int64_t m_out, int32_t cin, int32_t K, int32_t cout,
                                        int32_t have_x_split, int32_t have_dy_split);

/* Training-mode BatchNorm on a row matrix x (m, c), fused with ReLU and the residual add: replaces
 * nn.BatchNorm1d/2d + nn.ReLU (+ `out + identity`) after the convolutions
 * (cpd/models/backbones_3d/spconv_backbone.py:13-35,100-136,410; base_bev_backbone.py:31-59;
 * center_head.py:22-27,73-80).  stats (2, c) = per-channel sum and sum of squares of x as accumulated
 * by cpd_gather_gemm; mean_invstd (2, c) is written for the backward; running_* (NULL ok) follow
 * torch semantics (momentum, unbiased variance).  y = relu?((x-mean)*invstd*gamma+beta (+residual)).
 * y_split / dx_split (NULL ok, c % 8 == 0): also write the split-row image (cpd_split_rows format) of y / dx in the
 * same pass -- the operand the next / previous convolution's tensor-core kernels gather from. */
CPD_API int32_t cpd_bn_train_fwd(const float *x, int64_t m, int32_t c, const float *stats, const float *gamma,
                         const float *beta, const float *residual, int32_t relu, float eps, float momentum,
                         float *running_mean, float *running_var, float *mean_invstd, float *y,
                         void *y_split, cpd_stream_t stream);
/* dz = dy*(y>0) if relu; dx = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)); dresidual (NULL ok) = dz;
 * dgamma_dbeta (2, c): row 0 = dbeta = sum dz, row 1 = dgamma = sum dz*xhat. */
CPD_API int32_t cpd_bn_train_bwd(const float *x, const float *y, const float *dy, int64_t m, int32_t c,
                         const float *mean_invstd, const float *gamma, int32_t relu, float *dx,
                         void *dx_split, float *dresidual, float *dgamma_dbeta, cpd_stream_t stream);

/* (cout, K, cin) -> (cin, K, cout), optionally reversing the tap order (the operand of
 * the input-gradient: SubM reuses its own table with flipped taps). */
CPD_API int32_t cpd_weight_transpose(const float *w, int32_t cout, int32_t K, int32_t cin, int32_t flip_taps,
                             float *wt, cpd_stream_t stream);

/* Neighbour table of a dense 2-D convolution over an NHWC image batch (n, h, w, c):
 * rows are pixels ((n*ho + y)*wo + x), taps (ky*kw + kx); replaces cuDNN's implicit
 * addressing for nn.Conv2d in cpd/models/backbones_2d/base_bev_backbone.py:31-47 and
 * cpd/models/dense_heads/center_head.py:11-45,73-80.  transposed != 0 builds the table
 * of the input-gradient / of nn.ConvTranspose2d (base_bev_backbone.py:48-59). */
CPD_API int32_t cpd_conv2d_table(int32_t n, int32_t h, int32_t w, int32_t kh, int32_t kw, int32_t stride,
                         int32_t pad, int32_t transposed, int32_t ho, int32_t wo, int32_t *nbr,
                         cpd_stream_t stream);

/* ---------------------------------------------------------------------------------
 * Dense 2-D convolutions over NHWC image batches, stride 1: nn.Conv2d / nn.ConvTranspose2d of
 * cpd/models/backbones_2d/base_bev_backbone.py:31-59 and cpd/models/dense_heads/center_head.py:11-45,73-80
 * (cuDNN in the reference).  Same arithmetic and epilogue as cpd_gather_gemm (bf16x3 on tcgen05), but the A operand
 * of tap (ky, kx) is the output tile's 16 x 8 pixel patch shifted by the tap: one tiled TMA load
 * (cp.async.bulk.tensor.4d over the split-row image seen as [n][h][w][2 cin] bf16) per (tap, 64-channel block, hi / lo),
 * with the zero padding and the image border filled by the TMA unit -- no neighbour table, no gather.
 * x_split: split-row image (cpd_split_rows) of x (n*h*w, cin); wgt: (cout, kh*kw, cin); y: (n*ho*wo, cout),
 * ho = h + 2 pad - kh + 1.  Needs cin % 64 == 0 and cout in {16, 32, 64, 128, 256 j} (cpd_conv2d_supported); the other
 * dense shapes of the model (stride 2, 1-3 output channels) run through cpd_conv2d_table + cpd_gather_gemm.
 * --------------------------------------------------------------------------------- */
CPD_API int32_t cpd_conv2d_supported(int32_t cin, int32_t kh, int32_t kw, int32_t cout);
CPD_API size_t cpd_conv2d_workspace_bytes(int32_t cin, int32_t kh, int32_t kw, int32_t cout);
CPD_API int32_t cpd_conv2d_fwd(const void *x_split, int32_t n, int32_t h, int32_t w, int32_t cin, const float *wgt, int32_t kh,
                       int32_t kw, int32_t pad, int32_t cout, const float *bias, const float *scale, const float *shift,
                       const float *residual, int32_t relu, float *stats, float *y, void *ws, size_t ws_bytes, cpd_stream_t stream);
/* input-gradient dx (n*h*w, cin) from the split-row image of dy (n*ho*wo, cout): the same kernel with the flipped,
 * transposed weights (built in the workspace) and padding k - 1 - pad.  Needs cout % 64 == 0, cin in {16, ..., 256 j}. */
CPD_API size_t cpd_conv2d_dgrad_workspace_bytes(int32_t cin, int32_t kh, int32_t kw, int32_t cout);
CPD_API int32_t cpd_conv2d_dgrad(const void *dy_split, int32_t n, int32_t h, int32_t w, int32_t cin, const float *wgt, int32_t kh,
                         int32_t kw, int32_t pad, int32_t cout, float *dx, void *ws, size_t ws_bytes, cpd_stream_t stream);
/* weight gradient dw (cout, kh*kw, cin), overwritten: the row-stationary tcgen05 kernel of cpd_gather_wgrad over a pixel
 * table generated into the workspace (callers that keep the table use cpd_gather_wgrad directly). */
CPD_API size_t cpd_conv2d_wgrad_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t kh, int32_t kw, int32_t pad);
CPD_API int32_t cpd_conv2d_wgrad(const void *x_split, const void *dy_split, int32_t n, int32_t h, int32_t w, int32_t cin, int32_t kh,
                         int32_t kw, int32_t pad, int32_t cout, float *dw, void *ws, size_t ws_bytes, cpd_stream_t stream);
/* nn.ConvTranspose2d with kernel == stride == s (base_bev_backbone.py:48-59): s*s 1x1 GEMMs whose epilogues write the
 * pixel-shuffled (n, h*s, w*s, cout) map directly.  wgt: torch layout (cin, cout, s, s).  stats accumulate over the whole map. */
CPD_API size_t cpd_convt2d_workspace_bytes(int32_t cin, int32_t s, int32_t cout);
CPD_API int32_t cpd_convt2d_fwd(const void *x_split, int32_t n, int32_t h, int32_t w, int32_t cin, const float *wgt, int32_t s,
                        int32_t cout, const float *scale, const float *shift, int32_t relu, float *stats, float *y, void *ws,
                        size_t ws_bytes, cpd_stream_t stream);

/* SparseConvTensor.dense() fused with HeightCompression's view
 * (cpd/models/backbones_2d/map_to_bev/height_compression.py:136-138): scatter (m, c)
 * rows into a zeroed image.  channels_last == 0: out is (B, C, D, H, W) as upstream;
 * channels_last != 0: out is (B, H, W, C*D) with channel index c*D + z, i.e. the NHWC
 * form of the (B, C*D, H, W) BEV map.  The _bwd entry gathers the gradient back. */
CPD_API int32_t cpd_sparse_to_dense(const float *feat, const int32_t *coords, int64_t m, int32_t c, int32_t batch,
                            const int32_t *shape3_host, int32_t channels_last, float *out, cpd_stream_t stream);
CPD_API int32_t cpd_sparse_to_dense_bwd(const float *dout, const int32_t *coords, int64_t m, int32_t c, int32_t batch,
                                const int32_t *shape3_host, int32_t channels_last, float *dfeat,
                                cpd_stream_t stream);

/* ---------------------------------------------------------------------------------
 * RoI grid pooling primitives of VoxelRCNN[Proto]Head.  Replace
 * cpd/ops/pointnet2/pointnet2_stack/src/voxel_query_gpu.cu:10-89 (voxel_query_wrapper) and
 * group_points_gpu.cu:15-125 (group_points[_grad]_wrapper), as driven by
 * cpd/models/roi_heads/voxel_rcnn_head.py:186-273 and pointnet2_stack/voxel_query_utils.py:12-110.
 * The voxel -> row lookup probes the level's coordinate hash (cpd_coord_hash_build) instead of the dense
 * (B, Z, Y, X) int32 map the reference scatters per scale per call (cpd/utils/spconv_utils.py:4-21); passing that dense
 * map instead of the hash keeps the reference's own Python running on the shim.
 * new_xyz (m, 3) query points; new_coords (m, 4) int32 [b, z, y, x] their cells at this scale; xyz (n, 3) voxel centres;
 * idx (m, nsample) int32 GLOBAL rows (the reference returns the same and rebases per batch in Python); empty (m,) uint8.
 * Scan order, radius test and fill rule are the reference's: results are bit-identical.
 * --------------------------------------------------------------------------------- */
CPD_API int32_t cpd_voxel_query(const float *new_xyz, const int32_t *new_coords, int64_t m, const float *xyz, const void *hash,
                        size_t hash_bytes, const int32_t *dense_map, const int32_t *shape3_host, int32_t batch,
                        const int32_t *range3_host, float radius, int32_t nsample, int32_t *idx, uint8_t *empty,
                        cpd_stream_t stream);
/* out (m, c, nsample) = features[idx[m, s], c]; the _bwd entry overwrites grad_features (n, c) with the scattered sum. */
CPD_API int32_t cpd_group_points(const float *features, const int32_t *idx, int64_t m, int32_t c, int32_t nsample, float *out,
                         cpd_stream_t stream);
CPD_API int32_t cpd_group_points_bwd(const float *grad_out, const int32_t *idx, int64_t m, int32_t c, int32_t nsample, int64_t n,
                             float *grad_features, cpd_stream_t stream);

/* ---------------------------------------------------------------------------------
 * iou3d_nms.  Replace cpd/ops/iou3d_nms/src/iou3d_nms.h:9-12 (boxes_overlap_bev_gpu,
 * boxes_iou_bev_gpu, nms_gpu, nms_normal_gpu).  boxes: (n,7) fp32
 * [x,y,z,dx,dy,dz,heading].  NMS expects rows sorted by descending score, like the
 * reference (iou3d_nms_utils.py:111-115); keep (n,) int64 and n_keep (1,) int32 are
 * DEVICE buffers -- the reference's blocking D2H mask copy + host scan
 * (iou3d_nms.cpp:103-132) is replaced by an on-device scan.
 * --------------------------------------------------------------------------------- */
CPD_API int32_t cpd_overlap_bev(const float *a, int32_t na, const float *b, int32_t nb, float *out, cpd_stream_t stream);
CPD_API int32_t cpd_iou_bev(const float *a, int32_t na, const float *b, int32_t nb, float *out, cpd_stream_t stream);
CPD_API size_t cpd_nms_workspace_bytes(int32_t n);
CPD_API int32_t cpd_nms_rotated(const float *boxes, int32_t n, float thresh, int64_t *keep, int32_t *n_keep,
                        void *ws, size_t ws_bytes, cpd_stream_t stream);
CPD_API int32_t cpd_nms_normal(const float *boxes, int32_t n, float thresh, int64_t *keep, int32_t *n_keep,
                       void *ws, size_t ws_bytes, cpd_stream_t stream);
/* the suppression bit-matrix alone (n x ceil(n/64) uint64, upper-triangle tiles; tiles
 * below the diagonal are zero -- the reference computes them but never reads them,
 * iou3d_nms.cpp:128) -- exposed for parity tests against the reference kernel. */
CPD_API int32_t cpd_nms_mask(const float *boxes, int32_t n, float thresh, int32_t rotated, uint64_t *mask,
                     cpd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CPD_B200_H_ */
