// Shared declarations of libbnv_b200: device-side map layout, error plumbing, small helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bnv_b200.h"

namespace bnv {

constexpr int kFeat = 8;            // feature_vector_size (configs/model/fusion_pointnet_model.yaml:4)
constexpr int kWidth = 64;          // n_neurons (src/models/tcnn_config.json:28)
constexpr int32_t kEmpty = -1;      // slot-table sentinel
constexpr double kFixScale = 1073741824.0;  // 2^30: fixed-point scale of the per-frame sums
constexpr int kMaxBatch = 7;        // frames per bnv_fuse_frames call (8 table words per grid cell, one is the lock)

// latched device-side status bits
constexpr int kErrCapacity = 1;     // value pool or frame scratch full
constexpr int kErrRange = 2;        // key outside the grid
constexpr int kErrExchange = 4;     // peer-memory halo exchange: a peer's frame never arrived (bnv_p2p.cu)

// floor (0) / ceil (1) flavour of corner k per axis, get_neighbors' order (src/models/fusion/utils.py:98-167):
// k: (f,f,f) (c,f,f) (f,c,f) (f,f,c) (c,c,f) (c,f,c) (f,c,c) (c,c,c)
__host__ __device__ constexpr int corner_sx(int k) { return (0xB2 >> k) & 1; }
__host__ __device__ constexpr int corner_sy(int k) { return (0xD4 >> k) & 1; }
__host__ __device__ constexpr int corner_sz(int k) { return (0xE8 >> k) & 1; }

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);

#define BNV_CUDA(expr)                                  \
  do {                                                  \
    cudaError_t _e = (expr);                            \
    if (_e != cudaSuccess) return bnv::cuda_fail(_e, #expr); \
  } while (0)

#define BNV_LAUNCH_CHECK(name)                               \
  do {                                                       \
    bnv::count_launch();                                     \
    cudaError_t _e = cudaGetLastError();                     \
    if (_e != cudaSuccess) return bnv::cuda_fail(_e, name);  \
  } while (0)

// Geometry in the form the kernels consume.
struct GeomDev {
  float bmin[3];
  float lo[3];          // bmin + (float)vs   (rule A1 bounds, fp32)
  float hi[3];          // bmax - (float)vs
  float vs;             // (float)voxel_size
  float inv_vs;         // 1.0f / (float)voxel_size  (PyTorch-CUDA true-div fast path, rule A2)
  int32_t n[3];
  int32_t nyz;          // n[1]*n[2]
  int64_t n_vox;
  // tile shard (multi-GPU): owner(x,y,z) = ((x >> b) + (y >> b) + (z >> b)) % world  (3-D brick checkerboard)
  int32_t rank, world, brick_log2;
};

// Device-side view of the voxel map.
struct MapDev {
  GeomDev g;
  // persistent map: flat id -> slot (identity-hashed table), dense SoA value pool in slot order
  int32_t* table;       // [n_vox]
  int32_t* keys;        // [cap] flat id of slot
  float* feats;         // [cap, 8]
  float* weights;       // [cap]
  float* hits;          // [cap]
  int32_t cap;
  // per-frame scratch.  ftable[flat] is ONE 64-bit word per grid cell: low 32 bits = rows (point, corner) that
  // fell into the voxel this frame -- the exact integer count of scatter_mean, and the claim at the same time: the
  // row whose atomicAdd returns count 0 touched the voxel first -- high 32 bits = the voxel's dense scratch row,
  // allocated by that first row from ctr[1].  Scratch rows are therefore contiguous in first-touch order
  // (fkeys / fsum rows [0, n_touched)), finalize streams them and zeroes the table entries it visits.
  //
  // Frame batches (bnv_fuse_frames): every grid cell owns 2^fshift consecutive words -- one per frame of the batch plus
  // the LAST one as the cell's finalize lock -- so the entries of a voxel that consecutive frames keep touching share
  // a cache line.  fshift == 0 (default): one word per cell, single frames only.
  unsigned long long* ftable;  // [n_vox << fshift]
  int32_t fshift;
  unsigned long long* ftable_dummy;   // [1024] sink of the prepass' no-op atomics (lanes without a run to count)
  int32_t* fkeys;       // [fcap] flat id of scratch row (a (frame, voxel) pair in a batch)
  long long* fsum;      // [fcap, 8] 2^30 fixed-point int64 sums (exact-parity mode: order-independent => deterministic);
                        // the tensor-core mode uses the same buffer as float [fcap, 8]
  float* prec;          // [max_points, 8] compacted in-bounds point records of the frame: voxel-space xyz, normal, pad
  int32_t fcap;         // min(8 * max_points, n_vox * max(1, frames per batch))
  // device counters: [0] n_slots, [1] n_touched, [2] status bits, [3] finalize block counter, [4] n point records,
  // [5] n dirty shell voxels
  int32_t* ctr;
  // tile shard (nullable): voxels on the shell of their brick that were integrated since the last boundary exchange
  // are remembered once each (flag per pool slot + list of slots, counter ctr[5]); bnv_map_halo_pack turns the list
  // into records [int32 count, pad[9], records of 10 x 4 B] with the voxels' CURRENT values
  int32_t* dirty_flag;  // [cap]
  int32_t* dirty_list;  // [dirty_cap]
  int32_t dirty_cap;
};

// table word of (grid cell, frame of the batch)
__host__ __device__ inline unsigned long long* ft_entry(const MapDev& m, int32_t flat, int frame) {
  return m.ftable + (((size_t)flat << m.fshift) + (size_t)frame);
}

__host__ __device__ inline int owner_of(const GeomDev& g, int x, int y, int z) {
  return ((x >> g.brick_log2) + (y >> g.brick_log2) + (z >> g.brick_log2)) % g.world;
}
__host__ __device__ inline bool owns(const GeomDev& g, int x, int y, int z) {
  return g.world <= 1 || owner_of(g, x, y, z) == g.rank;
}
// voxel on the outer shell of its brick: some rank other than its owner may need it as a corner
__host__ __device__ inline bool on_brick_shell(const GeomDev& g, int x, int y, int z) {
  const int m = (1 << g.brick_log2) - 1;
  const int a = x & m, b = y & m, c = z & m;
  return a == 0 || a == m || b == 0 || b == m || c == 0 || c == m;
}
// does `rank` own a brick that touches voxel (x,y,z) (26-neighbourhood) ?
__host__ __device__ inline bool rank_touches(const GeomDev& g, int x, int y, int z, int rank) {
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dz = -1; dz <= 1; ++dz) {
        const int u = x + dx, v = y + dy, w = z + dz;
        if (u < 0 || v < 0 || w < 0 || u >= g.n[0] || v >= g.n[1] || w >= g.n[2]) continue;
        if (owner_of(g, u, v, w) == rank) return true;
      }
  return false;
}

}  // namespace bnv

// The opaque handles of the C ABI.
struct bnv_map {
  bnv::MapDev d;
  int device;
  int batch_cap;        // frames per bnv_fuse_frames call the per-frame table is laid out for (0: single frames only)
  unsigned int batch_seq;   // sequence number of the last batch (the finalize lock value; never 0)
  // scratch for the sorted encode_points path (CUB temp storage + index arrays)
  void* cub_tmp;
  size_t cub_tmp_bytes;
  int32_t* sort_keys_in;
  int32_t* sort_keys_out;
  int32_t* sort_vals_in;
  int32_t* sort_vals_out;
  int32_t* flags;
  int32_t* scan;
  // dense back-projection staging for bnv_backproject
  // bnv_fuse_frame_host: two device staging buffers [max_points] uint16, a copy stream for the prefetch hint
  uint16_t* depth_stage[2];
  cudaStream_t copy_stream;
  cudaEvent_t stage_ready[2], stage_free[2];
  const void* prefetched[bnv::kMaxBatch];  // host frames whose copy into depth_stage[stage_next] is in flight / done
  int n_prefetched;
  size_t prefetched_bytes;   // per frame
  int stage_next;
  int64_t* user_stats;   // device int64[4] frame statistics of bnv_fuse_frame_host
  float* bp_pts;
  int32_t* bp_flags;
  int32_t* bp_scan;
  int64_t max_points;
  int64_t* stats;       // device int64[8] frame statistics accumulators
  void* dec_pack;       // [cap] x 16 B: fp16x8 copy of the exported features for the tensor-core decode gather
  void* gtable;         // [(cap + 1) x 27] float: G[V][l] of the factored meshlize decode
  size_t gtable_bytes;
  int timing;           // bnv_map_set_timing
  cudaEvent_t ev[4];    // before the prepass / after the MLP kernel / after finalize / after the prepass
};

struct bnv_mlp {
  int n_in, n_out, in_pad, out_pad, device;
  int64_t n_params;
  float* w32;           // device fp32 k-major image for the CUDA-core kernels (bnv_mlp_simt.cuh)
  float* wraw;          // device fp32 copy of params as given (row-major [out,in] blocks): backward pass
  void* w16;            // device fp16 image in the UMMA canonical shared-memory layout
  size_t w16_bytes;
};
