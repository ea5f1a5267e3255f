/* bnv_b200.h -- C ABI of libbnv_b200.so: BNV-Fusion's per-frame dense hot path on B200 (sm_100a).
 *
 * The reference (likojack/bnv_fusion, pure Python) has no FFI of its own; this is the boundary a
 * maintainer binds with ctypes to replace the third-party native code the
 * path runs on today (tinycudann, torch_scatter, open3d.core.HashMap, torch.unique, grid_sample,
 * kornia).  Each entry point cites the reference interface it replaces (paths relative to the
 * reference repo root).
 *
 * Conventions
 *   - plain pointers and sizes only; every `dev` pointer is device memory owned by the caller
 *     (e.g. torch.Tensor.data_ptr()), every `host` pointer is ordinary host memory.
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream) and never synchronise the host unless stated ("host sync").
 *   - return 0 on success, a negative BNV_E* code otherwise; bnv_last_error() gives the message
 *     (thread-local).  Device-side faults that cannot be reported synchronously (capacity
 *     overflow, key outside the grid) are latched in the map and returned by bnv_map_status().
 *   - float maths follows the op order PyTorch-CUDA executes for the reference (SURVEY.md §8a
 *     rules A1-A8 / D1-D7), in particular `tensor / python_scalar` == x * (1.0f / (float)s).
 */
#ifndef BNV_B200_H
#define BNV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNV_ABI_VERSION 1

#define BNV_OK 0
#define BNV_E_ARG (-1)      /* bad argument */
#define BNV_E_CUDA (-2)     /* CUDA runtime error (message has the cudaError_t string) */
#define BNV_E_CAPACITY (-3) /* a fixed-capacity buffer would overflow */
#define BNV_E_RANGE (-4)    /* voxel key outside the grid */
#define BNV_E_ALLOC (-5)    /* out of memory */
#define BNV_E_UNSUPPORTED (-6)

/* MLP arithmetic selector for bnv_mlp_create / the encode and decode calls. */
#define BNV_MLP_FP32 0 /* fp32 CUDA-core path: the exact-parity mode (fp32 oracle, ~1e-6) */
#define BNV_MLP_TC16 1 /* tcgen05 tensor-core path: fp16 operands, fp32 accumulators in TMEM */

typedef struct bnv_map bnv_map_t; /* sparse voxel map  == reference SparseVolume */
typedef struct bnv_mlp bnv_mlp_t; /* packed weights of one tiny MLP == tcnn.NetworkWithInputEncoding */

/* Geometry of the voxel grid (reference: SparseVolume.__init__, src/models/sparse_volume.py:485-497,
 * voxel_utils.get_world_range, src/utils/voxel_utils.py:83-88).  bmin/bmax are the float32-cast
 * world bounds, voxel_size the Python double the reference passes around. */
typedef struct bnv_geom {
  float bmin[3];
  float bmax[3];
  double voxel_size;
  int32_t n_xyz[3];
} bnv_geom_t;

int bnv_abi_version(void);
const char* bnv_last_error(void);

/* ---- tiny MLPs ------------------------------------------------------------------------------
 * Replaces tcnn.NetworkWithInputEncoding(Identity -> FullyFusedMLP, 64 neurons, 3 hidden layers,
 * ReLU, no bias) constructed at src/utils/pointnet_utils.py:274-279 (encoder, 6 -> 8) and
 * src/models/fusion/modules.py:171-176 (decoder, 17 -> 1).  `params_host` is the flat float32
 * `model.params` tensor of the checkpoint (10240 / 11264 floats): [W0 | W1 | W2 | W3], each block
 * row-major [out, in], input padded with ones to a multiple of 16 (tiny-cuda-nn semantics). */
int bnv_mlp_create(bnv_mlp_t** out, const float* params_host, int64_t n_params, int n_in, int n_out,
                   int device);
int bnv_mlp_destroy(bnv_mlp_t* mlp);
/* Plain batched forward (tcnnPointNetEncoder.forward, pointnet_utils.py:283-294;
 * tcnnNeRFModel.geo_forward, modules.py:249-253).  x_dev [n, n_in] fp32 -> y_dev [n, n_out] fp32. */
int bnv_mlp_forward(const bnv_mlp_t* mlp, const float* x_dev, int64_t n, float* y_dev, int mode,
                    void* stream);

/* ---- sparse voxel map -----------------------------------------------------------------------
 * Replaces SparseVolume over open3d.core.HashMap (src/models/sparse_volume.py:484-600).  Keys are
 * voxel coordinates (x,y,z) inside the grid; the reference's int32 flat id
 * (voxel_utils.flatten, src/utils/voxel_utils.py:62-65) is < 2^31 by construction, so the map uses
 * the flat id itself as a collision-free index into an HBM-resident slot table.
 * `capacity` = rows of the value pool (active voxels), `max_points` = most points one frame may
 * carry (sizes the per-frame accumulators: 8 rows per point). */
int bnv_map_create(bnv_map_t** out, const bnv_geom_t* geom, int n_feats, int64_t capacity,
                   int64_t max_points, int device);
int bnv_map_destroy(bnv_map_t* map);
/* Drop all voxels (SparseVolume.reset, sparse_volume.py:587-600). */
int bnv_map_reset(bnv_map_t* map, void* stream);
/* Number of active voxels and latched device-side status.  Host sync on `stream`.  bnv_map_size also fails (after
 * writing the count) when a fault is latched: a map that dropped voxels is never read silently. */
int bnv_map_size(bnv_map_t* map, int64_t* n_active_host, void* stream);
int bnv_map_status(bnv_map_t* map, void* stream);
/* Tile ownership for the multi-GPU shard: 3-D checkerboard of 2^brick_log2 bricks, a voxel belongs to
 * rank ((x >> b) + (y >> b) + (z >> b)) % world.
 * Encode calls drop (point, corner) rows whose voxel another rank owns.  world = 1 disables. */
int bnv_map_set_shard(bnv_map_t* map, int rank, int world, int brick_log2);

/* Boundary (halo) exchange of the tile shard (no reference equivalent: the reference is single-GPU).  A query's
 * 8 corners are floor/ceil voxels and meshlize samples id +- 0.5, so a rank also needs the one-voxel shell around each
 * of its bricks.  Halo copies are only ever READ (decode, query), never integrated into, so the exchange is decoupled
 * from the frame loop: with bnv_map_halo_enable, bnv_fuse_frame / bnv_fuse_points remember -- once per voxel -- every
 * voxel they integrate that lies on the outer shell of its brick; an exchange EPOCH (every K frames, and before the map
 * is read) turns that list into records with the voxels' current values
 *     struct { int32 flat_id; float weight; float feat[8]; }   (40 bytes)
 * in buf_dev = [int32 count, int32 pad[9], records...] (bnv_map_halo_pack, which also re-arms the list).  The caller
 * all-gathers the ranks' buffers (ONE NCCL all-gather per epoch) and hands the result to bnv_map_insert_halo, which
 * upserts the records of the other ranks that touch one of this rank's bricks.  capacity_records = 0 disables. */
#define BNV_HALO_RECORD_BYTES 40
int bnv_map_halo_enable(bnv_map_t* map, int64_t capacity_records);
int bnv_map_halo_pack(bnv_map_t* map, void* buf_dev, int64_t capacity_records, void* stream);
int bnv_map_insert_halo(bnv_map_t* map, const void* gathered_dev, int world, int64_t capacity_records,
                        void* stream);

/* Peer-memory variant of the exchange (csrc/bnv_p2p.cu): the sender routes each boundary record straight into the
 * inbox of the ranks that need it (stores over NVLink into cudaIpc-mapped memory), the receiver upserts its inbox on a
 * side stream once every peer's epoch sequence number has arrived -- no collective, no all-to-all traffic.  Usage, one
 * process per GPU: bnv_map_set_shard, bnv_map_halo_enable, bnv_exchange_create, bnv_exchange_handle (64 bytes),
 * all-gather the handles by any transport, bnv_exchange_connect; then bnv_exchange_push on every rank at the same
 * points of the frame stream (pack + route + upsert of one epoch); bnv_exchange_join before reading the map.  Only the
 * pack kernel and one event record / wait touch the fusing stream; routing and upsert run on a side stream. */
typedef struct bnv_exchange bnv_exchange_t;
int bnv_exchange_create(bnv_exchange_t** out, bnv_map_t* map, int64_t capacity_records_per_peer);
int bnv_exchange_handle(bnv_exchange_t* ex, void* handle64_out_host);
int bnv_exchange_connect(bnv_exchange_t* ex, const void* handles_host /* [world][64], rank order */);
int bnv_exchange_push(bnv_exchange_t* ex, void* stream);
int bnv_exchange_join(bnv_exchange_t* ex, void* stream);
int bnv_exchange_destroy(bnv_exchange_t* ex);

/* SparseVolume.query (sparse_volume.py:661-695): coords_dev [n,3] int64 -> feats [n,F], weights [n],
 * num_hits [n] (zeros for misses), found [n] uint8 (nullable). */
int bnv_map_query(bnv_map_t* map, const int64_t* coords_dev, int64_t n, float* feats_dev,
                  float* weights_dev, float* hits_dev, uint8_t* found_dev, void* stream);
/* SparseVolume.insert (sparse_volume.py:561-585): upsert -- insert absent keys, overwrite all three
 * values of present ones.  Duplicate keys inside one call: one of them wins (as with o3c). */
int bnv_map_insert(bnv_map_t* map, const int64_t* coords_dev, const float* feats_dev,
                   const float* weights_dev, const float* hits_dev, int64_t n, void* stream);
/* SparseVolume.to_tensor (sparse_volume.py:525-559): copy the first n active rows (n from
 * bnv_map_size) to dense tensors.  Row r of the export is the map's slot r; the decode calls use
 * that identity instead of the reference's second hash map (tensor_indexer). */
int bnv_map_export(bnv_map_t* map, int64_t n, int64_t* coords_dev, float* feats_dev,
                   float* weights_dev, float* hits_dev, void* stream);
/* SparseVolume.count_optim (sparse_volume.py:602-622): weights_rows[row(key)] += 1 for every key
 * found among the first n_rows slots (non-accumulating for duplicate keys, like index_put). */
int bnv_map_count_optim(bnv_map_t* map, const float* nbr_coords_dev, int64_t n, float* weights_rows_dev,
                        int64_t n_rows, void* stream);

/* The same for the 8 floor/ceil corners of n query points (render_with_rays, src/utils/render_utils.py:494-496:
 * `coords = get_neighbors((pts - min_coords) / voxel_size); volume.count_optim(coords)`), without the [8 n, 3] tensor.
 * coords_dev [n,3] fp32: voxel units if is_coords, else world. */
int bnv_map_count_optim_queries(bnv_map_t* map, const float* coords_dev, int64_t n, int is_coords,
                                float* weights_rows_dev, int64_t n_rows, void* stream);

/* ---- per-frame local fusion -----------------------------------------------------------------
 * bnv_backproject: the dataset-side arithmetic of FusionInferenceAbstractDataset.__getitem__
 * (src/datasets/fusion_inference_dataset.py:52-74; load_depth, src/utils/common.py:86-120;
 * depth2xyz, src/utils/geometry.py:150-171; kornia depth_to_normals) in float64 on the device,
 * rounded to float32 like run_e2e.py:247-249.  depth_mm_dev [H,W] uint16 millimetres; K_host [9]
 * and T_wc_host [16] row-major float32.  Writes the masked pixels' [x,y,z,nx,ny,nz] in row-major
 * pixel order to pts6_dev [H*W,6] and their number to n_valid_dev. */
int bnv_backproject(bnv_map_t* map, const uint16_t* depth_mm_dev, int H, int W, const float* K_host,
                    const float* T_wc_host, double max_depth, float* pts6_dev, int32_t* n_valid_dev,
                    void* stream);

/* LitFusionPointNet.encode_pointcloud(..., return_dense=False)
 * (src/models/fusion/local_point_fusion.py:81-151): bound mask, 8-neighbour expansion, encoder MLP
 * per (point, corner) row, per-voxel mean, count >= min_pts filter, ascending flat-id order.
 * Outputs sized for `out_capacity` voxels: feats [cap,F] f32, counts [cap] i64, flat_ids [cap] i64,
 * coords [cap,3] i64; stats_dev int64[2] = {M (voxels written), M_t (touched voxels)};
 * navg_dev float = mean points per touched voxel.  M == 0 and M_t == 0 <=> the reference returns
 * five Nones. */
int bnv_encode_points(bnv_map_t* map, const float* pts6_dev, int64_t n_points, const bnv_mlp_t* enc,
                      int min_pts, int mode, float* feats_dev, int64_t* counts_dev,
                      int64_t* flat_ids_dev, int64_t* coords_dev, int64_t out_capacity,
                      int64_t* stats_dev, float* navg_dev, void* stream);

/* LitFusionPointNet._integrate / _update (local_point_fusion.py:647-673): weight = clip(count/32,1),
 * running weighted mean with the stored voxel, upsert (num_hits written back unchanged). */
int bnv_integrate(bnv_map_t* map, const int64_t* coords_dev, const float* feats_dev,
                  const int64_t* counts_dev, int64_t n, void* stream);

/* The whole of NeuralMap.integrate's local-fusion half (src/run_e2e.py:78-98) for one depth frame
 * in two kernels and no intermediate tensors: back-project -> encode -> integrate.  Leaves the map
 * in the same state as bnv_backproject + bnv_encode_points + bnv_integrate.  frame_stats_dev
 * (nullable) int64[4] = {valid pixels, rows, touched voxels, voxels integrated}; navg_dev nullable. */
int bnv_fuse_frame(bnv_map_t* map, const uint16_t* depth_mm_dev, int H, int W, const float* K_host,
                   const float* T_wc_host, double max_depth, const bnv_mlp_t* enc, int min_pts,
                   int mode, int64_t* frame_stats_dev, float* navg_dev, void* stream);
/* bnv_fuse_frame with HOST buffers, the shape of the reference's per-frame call (src/run_e2e.py:246-252:
 * the frame arrives in host memory, `.cuda()` per frame): depth_mm_host [H,W] uint16 (pinned memory for a
 * truly asynchronous copy) -> H2D into a map-owned staging buffer -> fuse -> D2H of the four frame
 * statistics into frame_stats_host (nullable, pinned), all enqueued on `stream`; no host sync inside.  The
 * caller synchronises the stream before reading frame_stats_host or reusing depth_mm_host.
 * next_depth_mm_host (nullable) is a prefetch hint, the role the reference's DataLoader workers play
 * (run_e2e.py:217-223): its H2D copy is started on the map's copy stream into the second staging buffer and
 * overlaps this frame's kernels; the next call that passes the same pointer as depth_mm_host skips its copy.
 * The hinted buffer must stay valid and unchanged until that next call's work has completed. */
int bnv_fuse_frame_host(bnv_map_t* map, const uint16_t* depth_mm_host, int H, int W, const float* K_host,
                        const float* T_wc_host, double max_depth, const bnv_mlp_t* enc, int min_pts,
                        int mode, int64_t* frame_stats_host, const uint16_t* next_depth_mm_host, void* stream);
/* ---- frame batches ----------------------------------------------------------------------------
 * The reference's driver loop (src/run_e2e.py:229-252: `for frame in loader: neural_map.integrate(frame)`) fuses one
 * frame per iteration.  Back-projection, encoding and the per-voxel means of a frame do not depend on the map, only
 * the final running average (_update, local_point_fusion.py:647-651) does, and it only has to be applied per voxel
 * in frame order.  bnv_fuse_frames therefore fuses n_frames depth frames in ONE pass of the three kernels -- the map
 * ends up exactly as after n_frames bnv_fuse_frame calls in the same order (bit-identical in BNV_MLP_FP32 mode) --
 * which amortises the kernels' fixed latencies over the batch (offline scenes, or a live stream that tolerates
 * n_frames of delay).
 *
 * bnv_map_set_frame_batch lays the per-frame table out for batches of up to n_frames (<= 7) frames: 8 table words
 * per grid cell instead of 1 (8.6 GB instead of 1.1 GB for a 512^3 grid); n_frames == 0 restores the single-frame
 * layout.  The map must have been created with max_points >= n_frames * H * W.  Synchronises the device. */
int bnv_map_set_frame_batch(bnv_map_t* map, int n_frames);
/* depth_mm_dev: HOST array of n_frames device pointers ([H,W] uint16 each); K_host [n_frames,3,3], T_wc_host
 * [n_frames,4,4] row-major fp32.  batch_stats_dev (nullable) int64[4] = the four bnv_fuse_frame statistics summed
 * over the batch; navg_dev nullable. */
int bnv_fuse_frames(bnv_map_t* map, const uint16_t* const* depth_mm_dev, int n_frames, int H, int W,
                    const float* K_host, const float* T_wc_host, double max_depth, const bnv_mlp_t* enc,
                    int min_pts, int mode, int64_t* batch_stats_dev, float* navg_dev, void* stream);
/* bnv_fuse_frames with HOST frames (depth_mm_host: n_frames host pointers, pinned for asynchronous copies) and the
 * prefetch hint of bnv_fuse_frame_host for the n_next frames of the next call (next_depth_mm_host nullable). */
int bnv_fuse_frames_host(bnv_map_t* map, const uint16_t* const* depth_mm_host, int n_frames, int H, int W,
                         const float* K_host, const float* T_wc_host, double max_depth, const bnv_mlp_t* enc,
                         int min_pts, int mode, int64_t* batch_stats_host,
                         const uint16_t* const* next_depth_mm_host, int n_next, void* stream);

/* Same, starting from world-space points (frame['input_pts'], [n,6] fp32). */
int bnv_fuse_points(bnv_map_t* map, const float* pts6_dev, int64_t n_points, const bnv_mlp_t* enc,
                    int min_pts, int mode, int64_t* frame_stats_dev, float* navg_dev, void* stream);

/* ---- SDF decode -----------------------------------------------------------------------------
 * SparseVolume.decode_pts (src/models/sparse_volume.py:768-833) with fusion/utils.get_neighbors
 * (src/models/fusion/utils.py:98-167), positional_encoding (src/models/fusion/modules.py:81-123),
 * tcnnNeRFModel.geo_forward (modules.py:249-253), _query_tensor (sparse_volume.py:625-659) and the
 * nearest-neighbour grid_sample of the TSDF prior (sparse_volume.py:819-832) fused into one kernel.
 * coords_dev [Q,3] fp32 (voxel units if is_coords, else world); feats_rows/weights_rows are the
 * exported (possibly optimised) tensors of bnv_map_export with n_rows rows; tsdf_delta_dev nullable
 * [Tx,Ty,Tz] fp32.  out_sdf_dev [Q]; out_mask_dev [Q] uint8 nullable. */
int bnv_decode_sdf(bnv_map_t* map, const float* coords_dev, int64_t n_queries, int is_coords,
                   const float* feats_rows_dev, const float* weights_rows_dev, int64_t n_rows,
                   const bnv_mlp_t* dec, int min_pts, int mode, const float* tsdf_delta_dev,
                   const int32_t* tsdf_dims_host, float* out_sdf_dev, uint8_t* out_mask_dev,
                   void* stream);
/* Backward of bnv_decode_sdf w.r.t. the exported features -- what torch autograd computes through
 * SparseVolume.decode_pts in NeuralMap.optimize (src/run_e2e.py:111-156, volume.features is the only leaf):
 * grad_feats_rows_dev [n_rows, F] += d(sum_q grad_out[q] * sdf[q]) / d(feats_rows).  The caller zero-fills
 * grad_feats_rows_dev.  fp32 CUDA cores. */
int bnv_decode_sdf_backward(bnv_map_t* map, const float* coords_dev, int64_t n_queries, int is_coords,
                            const float* feats_rows_dev, const float* weights_rows_dev, int64_t n_rows,
                            const bnv_mlp_t* dec, int min_pts, const float* grad_out_dev,
                            float* grad_feats_rows_dev, void* stream);

/* ---- global optimisation step ---------------------------------------------------
 * calculate_loss (src/utils/render_utils.py:559-594) around the decode: with bnv_map_count_optim_queries,
 * bnv_decode_sdf and bnv_decode_sdf_backward, one inner step of NeuralMap.optimize (src/run_e2e.py:111-156) is five
 * launches instead of ~100 small PyTorch kernels.
 * bnv_ray_samples: get_camera_params (:426-458) + hierarchical_sampling (:190-233) -- uv_dev [n,2] pixel coordinates,
 *   gt_pts_dev [n,3] world, K_host [9], T_wc_host [16]; t_fine_dev [n,n_fine] / t_coarse_dev [n,n_coarse] are the
 *   uniform draws of stratified_sampling (:91); pts_dev [n, n_fine + n_coarse, 3] world points (fine samples first:
 *   the reference sorts each ray's samples by distance, which the loss, a sum over samples, does not see).
 * bnv_ray_sdf_loss: compute_sdf_loss (:510-557) -- pred_sdf_dev [n,S] decoded at pts_dev; nbr_pts_dev [n,n_nbr,3],
 *   nbr_mask_dev [n,n_nbr], ray_mask_dev [n] as the dataset delivers them; n_valid_dev = sum(mask) + 1e-4 (float32
 *   device scalar, :576); loss_dev double device scalar (overwritten); grad_pred_dev [n,S] = d loss / d pred_sdf. */
int bnv_ray_samples(const float* uv_dev, const float* gt_pts_dev, int64_t n_rays, const float* K_host,
                    const float* T_wc_host, const float* t_fine_dev, int n_fine, const float* t_coarse_dev, int n_coarse,
                    double truncated_dist, float* pts_dev, void* stream);
int bnv_ray_sdf_loss(const float* pts_dev, const float* pred_sdf_dev, int64_t n_rays, int n_samples,
                     const float* gt_pts_dev, const float* T_wc_host, const float* nbr_pts_dev, const float* nbr_mask_dev,
                     int n_nbr, const float* ray_mask_dev, const float* n_valid_dev, double truncated_dist, double* loss_dev,
                     float* grad_pred_dev, void* stream);

/* The sampling half of SparseVolume.meshlize (sparse_volume.py:697-738) fused with the decode:
 * for active voxels [first, first+count) evaluate the 27 samples id + {-0.5,0,0.5}^3.
 * out_sdf_dev [count,27] ('ij' meshgrid order). */
int bnv_decode_voxel_blocks(bnv_map_t* map, int64_t first, int64_t count, const float* feats_rows_dev,
                            const float* weights_rows_dev, int64_t n_rows, const bnv_mlp_t* dec,
                            int min_pts, int mode, const float* tsdf_delta_dev,
                            const int32_t* tsdf_dims_host, float* out_sdf_dev, void* stream);

/* ---- mesh extraction ------------------------------------------------------
 * The second half of SparseVolume.meshlize (src/models/sparse_volume.py:738-766): marching cubes at level 0 over the
 * 3x3x3 sample block of every active voxel (sample spacing 0.5 voxel; a block is meshed iff max > 0 and min < 0, :740),
 * vertices at (coord - 0.5 + 0.5 * index) * voxel_size + min_coords (:755,762), per-block meshes concatenated in voxel
 * order (:758-761).  Replaces one skimage.measure.marching_cubes call per voxel on the CPU (with a D2H per 500 voxels).
 * sdf_blocks_dev [n,27] is the output of bnv_decode_voxel_blocks; coords_dev [n,3] int64 the exported voxel coordinates.
 *   bnv_mesh_count: tri_offsets_dev int32[n+1] <- exclusive scan of the triangles per voxel (offsets[n] = total).
 *   bnv_mesh_emit : verts_dev float32 [3T,3] (triangle t = rows 3t..3t+2, outward = towards sdf > 0), keys_dev int64 [3T]
 *                   (id of the half-voxel lattice edge the vertex lies on: equal ids <=> the same vertex, which is what
 *                   welds the per-block meshes without a distance threshold), tri_voxel_dev int32 [T] nullable. */
int bnv_mesh_count(const float* sdf_blocks_dev, int64_t n_voxels, int32_t* tri_offsets_dev, void* stream);
int bnv_mesh_emit(const float* sdf_blocks_dev, const int64_t* coords_dev, int64_t n_voxels, const int32_t* tri_offsets_dev,
                  double voxel_size, const float* min_coords_host, const int32_t* n_xyz_host, int64_t tri_capacity,
                  float* verts_dev, int64_t* keys_dev, int32_t* tri_voxel_dev, void* stream);

/* ---- coarse TSDF prior -------------------------------------------------
 * Replaces third_parties/fusion.py TSDFVolume (constructor :22-167, integrate :208-294 -- CPU mode
 * semantics, the reference's PyCUDA kernel :68-141 re-uploads both images per launch -- get_volume
 * :296-300) and NeuralMap.prepare_tsdf_volume (src/run_e2e.py:169-186).  The volume lives on the device.
 * vol_bnds_host = {x0,x1,y0,y1,z0,z1} metres (fusion.py's (3,2) array, row-major). */
typedef struct bnv_tsdf bnv_tsdf_t;
int bnv_tsdf_create(bnv_tsdf_t** out, const double* vol_bnds_host, double voxel_size, int device);
int bnv_tsdf_destroy(bnv_tsdf_t* tsdf);
int bnv_tsdf_dims(const bnv_tsdf_t* tsdf, int32_t* dims_host);
/* integrate(color_im, depth_im, cam_intr, cam_pose, obs_weight): rgb_dev [H,W,3] float32 0..255 (nullable:
 * skip the colour average); depth_dev [H,W] float32 metres, or uint16 millimetres when depth_is_u16_mm;
 * K_host [9]; Tinv_host [12] = rows 0..2 of inv(cam_pose) in float32 (np.linalg.inv, fusion.py:254). */
int bnv_tsdf_integrate(bnv_tsdf_t* tsdf, const float* rgb_dev, const void* depth_dev, int depth_is_u16_mm, int H,
                       int W, const float* K_host, const float* Tinv_host, double obs_weight, void* stream);
/* get_volume(): device pointers of the resident [Tx,Ty,Tz] float32 volumes (any may be NULL). */
int bnv_tsdf_volume(bnv_tsdf_t* tsdf, float** tsdf_dev, float** color_dev, float** weight_dev);
/* Snapshot of one volume into caller memory: which = 0 tsdf, 1 colour, 2 weight; out_dev [Tx,Ty,Tz] float32. */
int bnv_tsdf_copy(bnv_tsdf_t* tsdf, int which, float* out_dev, void* stream);
/* prepare_tsdf_volume: out = clip(tsdf * (voxel_size * 5), +-truncated_dist) * sdf_delta_weight, the array
 * bnv_decode_sdf takes as tsdf_delta_dev.  out_dev == NULL writes an internal buffer (see bnv_tsdf_volume). */
int bnv_tsdf_prior(bnv_tsdf_t* tsdf, double truncated_dist, double sdf_delta_weight, float* out_dev, void* stream);

/* Per-kernel device timing for bench.py's roofline: when enabled, bnv_fuse_frame / bnv_fuse_points
 * record CUDA events on `stream` around the encode kernels and the finalize kernel.
 * bnv_map_get_timing waits for the last recorded call (host sync) and returns both durations. */
int bnv_map_set_timing(bnv_map_t* map, int enable);
int bnv_map_get_timing(bnv_map_t* map, float* encode_ms_host, float* finalize_ms_host);
/* Same, per kernel of the frame: ms3_host = {prepass (back-projection + claims + compaction), encoder MLP kernel,
 * finalize}; encode_ms above is the sum of the first two. */
int bnv_map_get_timing_stages(bnv_map_t* map, float* ms3_host);

/* Number of kernels this library launched since load (bench.py's gpu_launches evidence). */
int64_t bnv_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BNV_B200_H */
