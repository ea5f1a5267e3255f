// Per-frame local fusion: back-projection, 8-neighbour expansion + encoder MLP + per-voxel
// scatter-mean, running-average integration into the voxel map.
//
// Reference call chain being replaced (paths relative to the reference repo):
//   src/datasets/fusion_inference_dataset.py:52-74  (depth -> world points + normals, CPU fp64)
//   src/models/fusion/local_point_fusion.py:81-165  (encode_pointcloud, get_relative_xyz)
//   src/models/fusion/modules.py:178-247            (get_neighbors)
//   src/utils/pointnet_utils.py:283-294             (tcnn encoder forward)
//   src/utils/voxel_utils.py:62-80                  (flatten / unflatten)
//   torch.unique + torch_scatter.scatter_mean       (local_point_fusion.py:118-126)
//   src/models/fusion/local_point_fusion.py:647-673 (_update / _integrate)
//
// Design: no sort and no per-row temporaries; three stream-ordered kernels per frame.
//   1. frame_prepass_kernel (full occupancy, latency-bound work): back-projection in float64, bound mask (A1),
//      voxel coordinates (A2), and for every (point, corner) row ONE 64-bit atomicAdd on the per-frame table
//      ftable[flat]: it counts the row (scatter_mean's exact integer count) and, when it returns 0, makes the row
//      the voxel's first toucher, which allocates the voxel's dense scratch row.  Runs of neighbouring pixels that
//      fall into the same voxel are merged by warp shuffle before the atomic (one add of the run length).  The
//      in-bounds points are compacted into 32-byte records (voxel-space xyz + normal).
//   2. the MLP kernel over the compacted records -- tcgen05 chain (bnv_tc_chain.cu) or fp32 CUDA cores
//      (encode_rows_simt_kernel below): 8 corner rows per point, features reduced into the dense scratch rows.
//   3. finalize_fused_kernel: streams the dense scratch rows (mean, count filter, running average into the
//      persistent map) and re-arms the table entries it visited.
// Exact-parity mode accumulates 2^30 fixed-point int64 sums: integer addition is associative, so the per-voxel
// mean is bit-reproducible regardless of the order in which warps arrive (the reference's fp16 atomics are not).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "bnv_common.cuh"
#include "bnv_decode_common.cuh"
#include "bnv_finalize.cuh"
#include "bnv_frame.cuh"
#include "bnv_mlp_simt.cuh"

namespace bnv {

// dense back-projection for bnv_backproject: all pixels + validity flags
__global__ void backproject_kernel(const uint16_t* __restrict__ depth, Camera cam, float* __restrict__ pts,
                                   int32_t* __restrict__ flags) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cam.H * cam.W) return;
  float p[6];
  const bool ok = backproject_pixel(depth, cam, idx % cam.W, idx / cam.W, p);
  flags[idx] = ok ? 1 : 0;
  if (ok) {
#pragma unroll
    for (int j = 0; j < 6; ++j) pts[(size_t)idx * 6 + j] = p[j];
  }
}

__global__ void compact_pts_kernel(const float* __restrict__ pts, const int32_t* __restrict__ flags,
                                   const int32_t* __restrict__ scan, int n, float* __restrict__ out,
                                   int32_t* __restrict__ n_valid) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  if (flags[idx]) {
    const int o = scan[idx];
#pragma unroll
    for (int j = 0; j < 6; ++j) out[(size_t)o * 6 + j] = pts[(size_t)idx * 6 + j];
  }
  if (idx == n - 1) *n_valid = scan[idx] + flags[idx];
}

// ------------------------------------------------------------------------------------------- //
// 1. prepass: back-project, bound mask, claim + count every (point, corner) row, compact the points
// ------------------------------------------------------------------------------------------- //
constexpr int kPreThreads = kTileW * kTileH;   // 256: one 32 x 8 pixel tile (a warp = 32 pixels of one image row)

// NP (key, count) pairs per thread: independent 32-bit atomics on the LOW word of the table entries (the count).  The
// first NP - 1 pairs exist for every lane of the warp: unconditional, straight-line code, so that all round trips to
// L2 are in flight before the first result is read; only the last one is predicated.  (Padding with no-op atomics on
// a dummy word, as this kernel first did, doubled the atomic traffic -- profiles/r2e.)
// Returns the mask of pairs whose voxel this thread touched first (count was 0).
// A pair = (flat voxel id, run length | frame << 8).
__device__ __forceinline__ int pair_frame(const int2& pr) { return pr.y >> 8; }
__device__ __forceinline__ unsigned int pair_run(const int2& pr) { return (unsigned int)(pr.y & 0xff); }

template <int NP>
__device__ __forceinline__ uint32_t claim_pairs(const MapDev& m, const int2 (&pr)[8]) {
  unsigned int old[NP];
#pragma unroll
  for (int j = 0; j < NP - 1; ++j)
    old[j] = atomicAdd(reinterpret_cast<unsigned int*>(ft_entry(m, pr[j].x, pair_frame(pr[j]))), pair_run(pr[j]));
  old[NP - 1] = 1u;
  if (pr[NP - 1].x >= 0)
    old[NP - 1] = atomicAdd(reinterpret_cast<unsigned int*>(ft_entry(m, pr[NP - 1].x, pair_frame(pr[NP - 1]))), pair_run(pr[NP - 1]));
  uint32_t win = 0;
#pragma unroll
  for (int j = 0; j < NP; ++j)
    if (old[j] == 0u) win |= 1u << j;
  return win;
}

// Warp-level claim + count of 32 points (one per lane; `own` == 0: no point): groups of lanes with the same 2 x 2 x 2
// corner block -> (key, group size) pairs per owned corner, compacted into the warp's shared-memory list, then claimed
// with dense independent atomics.  Neighbouring pixels mostly fall into the same voxel: a group is counted with ONE
// atomicAdd of its size per corner.  Two points share the voxel of one corner exactly when they share all eight (same
// floor voxel, same ceil - floor pattern), so the groups are found once per point (two match.any), not once per corner.  (Issued in place
// under `if (head)`, the compiler sinks each result test into its branch: eight serialised L2 round trips per thread,
// 55 % of the prepass' stall samples in profiles/r2a.)
// Returns the mask of this lane's pairs (pr) whose voxel it touched first.
__device__ __forceinline__ uint32_t claim_points(const MapDev& m, int2* __restrict__ runs, int lane, int fx, int fy, int fz, int ex,
                                                 int ey, int ez, uint32_t own, int frame, bool sharded, bool no_claims,
                                                 int2 (&pr)[8]) {
  const GeomDev& g = m.g;
  const int32_t key0 = own ? fx * g.nyz + fy * g.n[2] + fz : -1 - lane;   // rule A5 (int32); negatives never merge
  const int32_t pat = ex | (ey << 1) | (ez << 2) | (frame << 3);
  // lanes with the same corner block, wherever they sit in the warp (an 8 x 4 pixel patch in the depth path)
  const uint32_t group = __match_any_sync(0xffffffffu, key0) & __match_any_sync(0xffffffffu, pat);
  const bool lead = own != 0 && lane == __ffs(group) - 1;               // this lane writes its group's pairs
  int n_runs, at;                                                        // pairs of the warp (uniform) / before this lane
  if (!sharded) {
    const uint32_t real = __ballot_sync(0xffffffffu, lead);
    at = 8 * __popc(real & ((1u << lane) - 1u));
    n_runs = 8 * __popc(real);
  } else {
    const int mine = lead ? __popc(own) : 0;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    at = incl - mine;
    n_runs = __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lead) {
    const int run = __popc(group) | (frame << 8);
    const int dx = ex * g.nyz, dy = ey * g.n[2];
    int2* out = runs + at;
#pragma unroll
    for (int k = 0; k < 8; ++k)                                          // rule A3 corner order (modules.py:178-247)
      if ((own >> k) & 1u)
        *out++ = make_int2(key0 + (corner_sx(k) ? dx : 0) + (corner_sy(k) ? dy : 0) + (corner_sz(k) ? ez : 0), run);
  }
  __syncwarp();
  // lane j takes the warp's pairs j, j + 32, ...
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = j * 32 + lane;
    pr[j] = i < n_runs ? runs[i] : make_int2(-1, 0);
  }
  __syncwarp();
  uint32_t win = 0;
  switch (no_claims ? 0 : (n_runs + 31) >> 5) {                          // warp-uniform
    case 1: win = claim_pairs<1>(m, pr); break;
    case 2: win = claim_pairs<2>(m, pr); break;
    case 3: win = claim_pairs<3>(m, pr); break;
    case 4: win = claim_pairs<4>(m, pr); break;
    case 5: win = claim_pairs<5>(m, pr); break;
    case 6: win = claim_pairs<6>(m, pr); break;
    case 7: win = claim_pairs<7>(m, pr); break;
    case 8: win = claim_pairs<8>(m, pr); break;
    default: break;
  }
  return win;
}

// first touchers publish their dense scratch rows (first row = `row`)
__device__ __forceinline__ void publish_rows(const MapDev& m, const int2 (&pr)[8], uint32_t win, int row) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if ((win >> j) & 1u) {
      // the dense row in the entry's high word (visible to the kernels that follow) and its key
      reinterpret_cast<int32_t*>(ft_entry(m, pr[j].x, pair_frame(pr[j])))[1] = row;
      m.fkeys[row] = pr[j].x;
      // finalize will look this voxel up in the persistent table two kernels from now: pull the line into L2
      asm volatile("prefetch.global.L2 [%0];" ::"l"(m.table + pr[j].x));
      ++row;
    }
}

// `frame` = position of the frame in its batch (0 for a single frame): selects the frame's word of every table entry
// and tags the point records.
template <bool FROM_DEPTH>
__device__ __forceinline__ void prepass_body(const MapDev& m, const EncSrc& src, int frame, long long* __restrict__ stats, int dbg) {
  // dbg (tools/prepass_ablation.py only; 0 in production): 1 no claim atomics, 2 no global counter atomics,
  // 4 no back-projection (every pixel invalid), 8 no record / key stores, 16 no depth staging
  __shared__ FrameTile tile;
  __shared__ int2 s_runs[kPreThreads / 32][256];       // per warp: (voxel key, run length) of up to 8 x 32 runs
  __shared__ int s_new[kPreThreads / 32], s_keep[kPreThreads / 32], s_stat[3], s_base[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const GeomDev& g = m.g;
  if (tid < 3) s_stat[tid] = 0;
  float p[6];
  bool valid = false;
  int32_t pix = 0;
  if (FROM_DEPTH) {
    const int tiles_x = (src.cam.W + kTileW - 1) / kTileW;
    const int u0 = (blockIdx.x % tiles_x) * kTileW, v0 = (blockIdx.x / tiles_x) * kTileH;
    if (!(dbg & 16)) stage_frame_tile(tile, src.depth, src.cam, u0, v0);     // reads the depth image only
    __syncthreads();
    grid_dependency_wait();                                                  // the previous frame's finalize
    // a warp covers an 8 x 4 pixel patch of the tile (not 32 pixels of one image row): the patch is about as wide as it
    // is high in voxels, so more of its pixels share their corner block and are counted together (claim_points)
    const int tx = (warp & 3) * 8 + (lane & 7), ty = (warp >> 2) * 4 + (lane >> 3);
    const int u = u0 + tx, v = v0 + ty;
    pix = v * src.cam.W + u;
    bool foreign = false;
    const bool in_img = u < src.cam.W && v < src.cam.H && !(dbg & 4);
    if (g.world > 1) {
      // Tile shard: every rank sees every pixel, but only ~1/world of them land in its bricks.  A float32 estimate
      // of the point (error ~1e-4 voxel) is enough to tell when the whole 2 x 2 x 2 corner block, padded by a full
      // voxel, lies inside ONE brick of another rank: the float64 back-projection is skipped for those pixels, and a
      // tile without any other pixel (a 32 x 8 pixel tile is smaller than a brick: most tiles of most ranks) is done
      // after counting its valid pixels for the frame statistics.
      const float zf = in_img ? (float)tile.z[ty + 1][tx + 1] : 0.f;
      if (zf > 0.f) {
        const float xf = (float)tile.ax[tx + 1] * zf, yf = (float)tile.ay[ty + 1] * zf;
        int lo[3], hi[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float w = src.cam.T[r * 4] * xf + src.cam.T[r * 4 + 1] * yf + src.cam.T[r * 4 + 2] * zf + src.cam.T[r * 4 + 3];
          const float cv = (w - g.bmin[r]) * g.inv_vs;
          lo[r] = (int)floorf(cv - 1.05f) >> g.brick_log2;
          hi[r] = (int)floorf(cv + 2.05f) >> g.brick_log2;
        }
        foreign = lo[0] == hi[0] && lo[1] == hi[1] && lo[2] == hi[2] && (lo[0] + lo[1] + lo[2]) % g.world != g.rank &&
                  lo[0] >= 0 && lo[1] >= 0 && lo[2] >= 0;
      }
      if (__syncthreads_count(zf > 0.f && !foreign) == 0) {      // block-uniform
        const int n_valid = __syncthreads_count(zf > 0.f);
        if (tid == 0 && n_valid && !(dbg & 2))
          atomicAdd(reinterpret_cast<unsigned long long*>(stats + 0), (unsigned long long)n_valid);
        return;
      }
    }
    if (in_img) {
      if (foreign) valid = true;                 // a valid pixel (frame statistic) that contributes no row here
      else valid = backproject_tile_pixel(tile, src.cam, tx, ty, u, v, p);
    }
    if (foreign) {                               // park it outside the bounds: rule A1 drops it below
#pragma unroll
      for (int j = 0; j < 6; ++j) p[j] = 3.0e38f;
    }
  } else {
    const int64_t idx = (int64_t)blockIdx.x * kPreThreads + tid;
    pix = (int32_t)idx;
    if (idx < src.n_points) {
      valid = true;
#pragma unroll
      for (int j = 0; j < 6; ++j) p[j] = __ldg(src.pts6 + idx * 6 + j);
    }
    __syncthreads();
    grid_dependency_wait();
  }
  // rule A1 (local_point_fusion.py:94-100): strict bounds one voxel inside the volume
  bool inb = valid;
#pragma unroll
  for (int a = 0; a < 3; ++a) inb = inb && (p[a] < g.hi[a]) && (p[a] > g.lo[a]);
  float c[3], fl[3], ce[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    c[a] = inb ? __fmul_rn(__fsub_rn(p[a], g.bmin[a]), g.inv_vs) : 0.f;   // rule A2
    fl[a] = floorf(c[a]);
    ce[a] = ceilf(c[a]);
  }
  const bool sharded = g.world > 1;                                      // block-uniform
  const int fx = (int)fl[0], fy = (int)fl[1], fz = (int)fl[2];
  const int ex = (int)ce[0] - fx, ey = (int)ce[1] - fy, ez = (int)ce[2] - fz;          // 0 or 1 each
  uint32_t own = inb ? 0xffu : 0u;
  if (sharded && inb) {
    own = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)                                          // rule A3 corner order (modules.py:178-247)
      own |= (owner_of(g, fx + (corner_sx(k) ? ex : 0), fy + (corner_sy(k) ? ey : 0), fz + (corner_sz(k) ? ez : 0)) == g.rank ? 1u : 0u) << k;
  }
  // ---- claim + count.  Tried and removed (round 2): merging the tile's pairs per voxel in a
  // shared-memory hash table first (~6 x fewer global atomics: 148 instead of 153 us on 7 frames, 31 instead of 27 us
  // on one); leaving the claims to a second kernel over the compacted point records (195 instead of 156 us, 36 instead
  // of 27 us: the claims cost the same ~100 G atomics/s there, plus the records' second trip through L2) ----
  int2 pr[8];
  const uint32_t win = claim_points(m, s_runs[warp], lane, fx, fy, fz, ex, ey, ez, own, frame, sharded, (dbg & 1) != 0, pr);
  // ---- first touchers allocate dense scratch rows; in-bounds points get a record slot ---------------------
  const int n_new = __popc(win);
  int incl = n_new;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const bool keep = own != 0;
  const uint32_t kb = __ballot_sync(0xffffffffu, keep);
  const uint32_t vb = __ballot_sync(0xffffffffu, valid), ib = __ballot_sync(0xffffffffu, inb);
  int rows = __popc(own);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);
  if (lane == 31) s_new[warp] = incl;
  if (lane == 0) {
    s_keep[warp] = __popc(kb);
    if (vb) atomicAdd(&s_stat[0], __popc(vb));
    if (rows) atomicAdd(&s_stat[1], rows);
    if (ib) atomicAdd(&s_stat[2], __popc(ib));
  }
  __syncthreads();
  if (lane == 0 && warp < 3) {       // three warps: the two allocations and the statistics go out in parallel
    if (warp == 0) {
      int t = 0;
      for (int w = 0; w < kPreThreads / 32; ++w) t += s_new[w];
      s_base[0] = (t && !(dbg & 2)) ? atomicAdd(&m.ctr[1], t) : 0;
    } else if (warp == 1) {
      int t = 0;
      for (int w = 0; w < kPreThreads / 32; ++w) t += s_keep[w];
      s_base[1] = (t && !(dbg & 2)) ? atomicAdd(&m.ctr[4], t) : (dbg & 2) ? (int)blockIdx.x * kPreThreads : 0;
    } else if (!(dbg & 2)) {
      if (s_stat[0]) atomicAdd(reinterpret_cast<unsigned long long*>(stats + 0), (unsigned long long)s_stat[0]);
      if (s_stat[1]) atomicAdd(reinterpret_cast<unsigned long long*>(stats + 1), (unsigned long long)s_stat[1]);
      if (s_stat[2]) atomicAdd(reinterpret_cast<unsigned long long*>(stats + 4), (unsigned long long)s_stat[2]);
    }
  }
  __syncthreads();
  int row = s_base[0] + incl - n_new, rec = s_base[1] + __popc(kb & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) {
    row += s_new[w];
    rec += s_keep[w];
  }
  publish_rows(m, pr, win, row);
  if (keep && !(dbg & 8)) {
    float4* r4 = reinterpret_cast<float4*>(m.prec + (size_t)rec * 8);
    r4[0] = make_float4(c[0], c[1], c[2], p[3]);
    r4[1] = make_float4(p[4], p[5], __int_as_float((int)own),
                        __int_as_float((int)(((uint32_t)pix & 0xffffffu) | ((uint32_t)frame << 24))));
  }
}

template <bool FROM_DEPTH>
__global__ void __launch_bounds__(kPreThreads, 5) frame_prepass_kernel(MapDev m, EncSrc src, long long* __restrict__ stats, int dbg) {
  prepass_body<FROM_DEPTH>(m, src, 0, stats, dbg);
}

// a batch of frames in one launch: blockIdx.y = frame, blockIdx.x = pixel tile
__global__ void __launch_bounds__(kPreThreads, 5) frame_prepass_batch_kernel(MapDev m, FrameBatch fb, long long* __restrict__ stats, int dbg) {
  const int frame = blockIdx.y;
  EncSrc src;
  src.depth = fb.depth[frame];
  src.cam = fb.cam[frame];
  src.pts6 = nullptr;
  src.n_points = 0;
  prepass_body<true>(m, src, frame, stats, dbg);
}

// ------------------------------------------------------------------------------------------- //
// 2. exact-parity mode: encoder MLP on the CUDA cores, one thread per point record, 8 corner rows each
// ------------------------------------------------------------------------------------------- //
using EncMlp = SimtMlp<6, 8>;
constexpr int kEncThreads = 256;
constexpr size_t kEncSmem = (size_t)(EncMlp::kFloats + 64 * kEncThreads) * sizeof(float);

__global__ void __launch_bounds__(kEncThreads) encode_rows_simt_kernel(MapDev m, const float* __restrict__ gW) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* sH = smem + EncMlp::kFloats + threadIdx.x;
  load_weights(sW, gW, EncMlp::kFloats);
  grid_dependency_wait();                               // the prepass
  const int n_rec = m.ctr[4];
  const GeomDev& g = m.g;
  for (int64_t idx = (int64_t)blockIdx.x * kEncThreads + threadIdx.x; idx < n_rec; idx += (int64_t)gridDim.x * kEncThreads) {
    const float4* r4 = reinterpret_cast<const float4*>(m.prec + (size_t)idx * 8);
    const float4 a = __ldg(r4), b = __ldg(r4 + 1);
    const float c[3] = {a.x, a.y, a.z}, nrm[3] = {a.w, b.x, b.y};
    const uint32_t own = (uint32_t)__float_as_int(b.z);
    const int frame = record_frame(b.w);
    float fl[3], ce[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      fl[ax] = floorf(c[ax]);
      ce[ax] = ceilf(c[ax]);
    }
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      if (!((own >> k) & 1u)) continue;
      float nb[3];
      corner_of(k, fl, ce, nb);
      float x[6], y[8];
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        const float rel_n = __fsub_rn(c[ax], nb[ax]);              // rule A4
        const float rel = __fmul_rn(rel_n, g.vs);
        x[ax] = __fmul_rn(rel, g.inv_vs);
        x[3 + ax] = nrm[ax];
      }
      EncMlp::run(sW, sH, kEncThreads, x, y);
      add_row_fixed(m, scratch_row_of(m, (int)nb[0] * g.nyz + (int)nb[1] * g.n[2] + (int)nb[2], frame), y);
    }
  }
}

// 3. finalize of the fused path: for every voxel touched this frame (dense scratch rows [0, n_touched)) -> mean,
// count filter, running average into the persistent map; clears the scratch.  ONE THREAD PER VOXEL: the kernel is a
// chain of dependent scattered reads (key -> table entries -> map slot -> old features), i.e. latency x concurrency
// bound, so every thread keeps its own voxel's chain in flight (an 8-lanes-per-voxel layout had 8x fewer chains in
// flight and ran 23 us on 1.3e5 voxels); the 32-byte scratch and feature rows move as two 16-byte accesses.
__global__ void __launch_bounds__(256) finalize_fused_kernel(MapDev m, int min_pts, bool f32acc, long long* __restrict__ stats,
                                                             long long* __restrict__ user_stats,
                                                             float* __restrict__ user_navg) {
  grid_dependency_wait();                               // the encoder MLP kernel
  const int n_touched = m.ctr[1];
  const int integrated = finalize_rows(m, min_pts, f32acc, n_touched, (int64_t)blockIdx.x * blockDim.x + threadIdx.x,
                                       (int64_t)gridDim.x * blockDim.x);
  finalize_publish(m, integrated, n_touched, stats, user_stats, user_navg);
}

// the same for a batch of frames: the first of a voxel's (frame, voxel) scratch rows to arrive makes its thread do all
// the voxel's frames
template <bool F32>
__global__ void __launch_bounds__(256, 2) finalize_batch_kernel(MapDev m, int min_pts, unsigned int seq, long long* __restrict__ stats,
                                                                long long* __restrict__ user_stats, float* __restrict__ user_navg) {
  grid_dependency_wait();                               // the encoder MLP kernel
  const int n_touched = m.ctr[1];
  const int integrated = finalize_batch_rows<F32>(m, min_pts, seq, n_touched);
  finalize_publish(m, integrated, n_touched, stats, user_stats, user_navg);
}

// ---- sorted path (encode_pointcloud's return values) ----------------------------------------
__global__ void sort_prep_kernel(MapDev m, int n, int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  keys[t] = m.fkeys[t];
  vals[t] = t;
}

__global__ void sort_flag_kernel(MapDev m, int n, const int32_t* __restrict__ keys, int min_pts,
                                 int32_t* __restrict__ flags) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  flags[t] = ft_count(*ft_entry(m, keys[t], 0)) >= min_pts ? 1 : 0;
}

__global__ void sort_emit_kernel(MapDev m, int n, const int32_t* __restrict__ keys, const int32_t* __restrict__ rows,
                                 const int32_t* __restrict__ flags, const int32_t* __restrict__ scan, bool f32acc,
                                 int64_t out_cap, float* __restrict__ feats, int64_t* __restrict__ counts,
                                 int64_t* __restrict__ flat_ids, int64_t* __restrict__ coords) {
  const int64_t tt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int t = (int)(tt >> 3);
  const int j = (int)(tt & 7);
  if (t >= n) return;
  const int32_t row = rows[t];
  const int32_t key = keys[t];
  const int32_t cnt = ft_count(*ft_entry(m, key, 0));
  const float mean = scratch_mean(m, row, j, cnt, f32acc);
  scratch_clear(m, row, j, f32acc);
  if (flags[t]) {
    const int64_t o = scan[t];
    if (o < out_cap) {
      feats[o * kFeat + j] = mean;
      if (j == 0) {
        counts[o] = cnt;
        flat_ids[o] = key;
        const int32_t x = key / m.g.nyz;                          // unflatten, voxel_utils.py:68-80
        const int32_t r = key - x * m.g.nyz;
        const int32_t y = r / m.g.n[2];
        coords[o * 3 + 0] = x;
        coords[o * 3 + 1] = y;
        coords[o * 3 + 2] = r - y * m.g.n[2];
      }
    } else if (j == 0) {
      atomicOr(&m.ctr[2], kErrCapacity);
    }
  }
  __syncwarp(0xFFu << ((threadIdx.x & 31) & ~7));
  if (j == 0) *ft_entry(m, key, 0) = 0ull;
}

__global__ void sort_publish_kernel(MapDev m, int n, const int32_t* __restrict__ flags, const int32_t* __restrict__ scan,
                                    long long* __restrict__ stats, int64_t* __restrict__ user_stats,
                                    float* __restrict__ user_navg) {
  const long long M = n > 0 ? (long long)scan[n - 1] + flags[n - 1] : 0;
  user_stats[0] = M;
  user_stats[1] = n;
  if (user_navg) *user_navg = n > 0 ? (float)((double)stats[1] / (double)n) : 0.f;
  stats[0] = stats[1] = stats[3] = stats[4] = 0;
  m.ctr[1] = 0;
  m.ctr[4] = 0;
}

// _integrate (local_point_fusion.py:653-673) on explicit arrays: 8 lanes per voxel
__global__ void integrate_kernel(MapDev m, const int64_t* __restrict__ coords, const float* __restrict__ feats,
                                 const int64_t* __restrict__ counts, int64_t n) {
  const int64_t tt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = tt >> 3;
  const int j = (int)(tt & 7);
  const bool act = i < n;
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
  int32_t slot = -1;
  if (act && j == 0) {
    const long long x = coords[i * 3], y = coords[i * 3 + 1], z = coords[i * 3 + 2];
    if (x < 0 || y < 0 || z < 0 || x >= m.g.n[0] || y >= m.g.n[1] || z >= m.g.n[2]) {
      atomicOr(&m.ctr[2], kErrRange);
    } else {
      const int32_t flat = (int32_t)x * m.g.nyz + (int32_t)y * m.g.n[2] + (int32_t)z;
      slot = m.table[flat];
      if (slot == kEmpty) {
        const int32_t tag = -(int32_t)(i % 0x3fffffff) - 2;
        const int32_t old = atomicCAS(&m.table[flat], kEmpty, tag);
        if (old == kEmpty) {
          slot = atomicAdd(&m.ctr[0], 1);
          if (slot < m.cap) {
            m.keys[slot] = flat;
            m.weights[slot] = 0.f;
            m.hits[slot] = 0.f;
            atomicExch(&m.table[flat], slot);
            slot |= 0x40000000;
          } else {
            atomicOr(&m.ctr[2], kErrCapacity);
            atomicExch(&m.table[flat], kEmpty);
            slot = -1;
          }
        } else {
          slot = old;   // negative tag: duplicate key in this call, the claimant integrates it
        }
      }
    }
  }
  slot = __shfl_sync(gmask, slot, (threadIdx.x & 31) & ~7);
  if (!act || slot < 0) return;
  const bool fresh = slot & 0x40000000;
  slot &= 0x3fffffff;
  const float w_new = fminf(__fmul_rn((float)counts[i], 0.03125f), 1.0f);
  const float w_old = fresh ? 0.f : m.weights[slot];
  const float f_old = fresh ? 0.f : m.feats[(size_t)slot * kFeat + j];
  const float w = __fadd_rn(w_old, w_new);
  m.feats[(size_t)slot * kFeat + j] = fuse_feat(f_old, w_old, feats[i * kFeat + j], w_new, w);
  __syncwarp(gmask);
  if (j == 0) m.weights[slot] = w;
}

static int make_camera(Camera& cam, int H, int W, const float* K, const float* T, double max_depth) {
  if (!K || !T || H <= 0 || W <= 0) { set_error("camera: bad arguments"); return BNV_E_ARG; }
  cam.fx = K[0]; cam.fy = K[4]; cam.cx = K[2]; cam.cy = K[5];
  for (int i = 0; i < 12; ++i) cam.T[i] = T[i];
  cam.max_depth = max_depth;
  cam.H = H; cam.W = W;
  return BNV_OK;
}

}  // namespace bnv

using namespace bnv;

// defined in bnv_mlp.cu / bnv_tc.cu
const float* bnv_internal_simt_weights(const bnv_mlp_t* mlp);
int bnv_internal_encode_tc(bnv_map_t* map, int64_t max_records, const bnv_mlp_t* enc, cudaStream_t s);

namespace bnv {
// launch with programmatic stream serialization (see grid_dependency_wait in bnv_frame.cuh)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static int g_prepass_debug = 0;      // profiling ablations only (bnv_debug_prepass, not part of the ABI header)

static int check_encoder(const bnv_mlp_t* enc, int mode) {
  if (!enc || enc->n_in != 6 || enc->n_out != 8) { set_error("encode: encoder MLP must be 6 -> 8"); return BNV_E_ARG; }
  if (mode != BNV_MLP_FP32 && mode != BNV_MLP_TC16) { set_error("encode: unknown MLP mode %d", mode); return BNV_E_ARG; }
  return BNV_OK;
}

// kernel 2 of the frame: the encoder MLP over the point records the prepass wrote (at most `max_records`)
static int launch_encode_rows(bnv_map_t* map, int64_t max_records, const bnv_mlp_t* enc, int mode, cudaStream_t s) {
  if (map->timing) BNV_CUDA(cudaEventRecord(map->ev[3], s));
  if (mode == BNV_MLP_TC16) return bnv_internal_encode_tc(map, max_records, enc, s);
  static bool attr[64] = {false};            // cudaFuncSetAttribute is per device
  if (!attr[map->device & 63]) {
    BNV_CUDA(cudaFuncSetAttribute(encode_rows_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEncSmem));
    attr[map->device & 63] = true;
  }
  const int64_t blocks = (max_records + kEncThreads - 1) / kEncThreads;
  BNV_CUDA(launch_pdl(encode_rows_simt_kernel, dim3((unsigned)(blocks < 148 * 2 ? blocks : 148 * 2)), dim3(kEncThreads),
                      kEncSmem, s, map->d, bnv_internal_simt_weights(enc)));
  BNV_LAUNCH_CHECK("encode_rows_simt_kernel");
  return BNV_OK;
}

// kernels 1 + 2 of the frame: prepass over `n_threads` pixels / points, then the encoder MLP over the records
static int launch_encode(bnv_map_t* map, const EncSrc& src, bool from_depth, int64_t n_threads,
                         const bnv_mlp_t* enc, int mode, cudaStream_t s) {
  int rc = check_encoder(enc, mode);
  if (rc) return rc;
  if (n_threads > map->max_points) {
    set_error("encode: %lld points exceed the map's max_points %lld", (long long)n_threads, (long long)map->max_points);
    return BNV_E_CAPACITY;
  }
  if (n_threads == 0) return BNV_OK;
  const dim3 grid(from_depth ? (unsigned)(((src.cam.W + kTileW - 1) / kTileW) * ((src.cam.H + kTileH - 1) / kTileH))
                             : (unsigned)((n_threads + kPreThreads - 1) / kPreThreads));
  BNV_CUDA(launch_pdl(from_depth ? frame_prepass_kernel<true> : frame_prepass_kernel<false>, grid, dim3(kPreThreads), 0, s, map->d,
                      src, (long long*)map->stats, g_prepass_debug));
  BNV_LAUNCH_CHECK("frame_prepass_kernel");
  return launch_encode_rows(map, n_threads, enc, mode, s);
}

static int launch_finalize(bnv_map_t* map, int min_pts, int mode, int64_t* frame_stats, float* navg, cudaStream_t s) {
  BNV_CUDA(launch_pdl(finalize_fused_kernel, dim3(148 * 4), dim3(256), 0, s, map->d, min_pts, mode == BNV_MLP_TC16,
                      (long long*)map->stats, (long long*)frame_stats, navg));
  BNV_LAUNCH_CHECK("finalize_fused_kernel");
  return BNV_OK;
}
}  // namespace bnv

extern "C" {

// profiling tool hook (tools/prepass_ablation.py); deliberately not declared in include/bnv_b200.h
int bnv_debug_prepass(int flags) { g_prepass_debug = flags; return BNV_OK; }

int bnv_backproject(bnv_map_t* map, const uint16_t* depth, int H, int W, const float* K, const float* T,
                    double max_depth, float* pts6, int32_t* n_valid, void* stream) {
  if (!map || !depth || !pts6 || !n_valid) { set_error("bnv_backproject: null argument"); return BNV_E_ARG; }
  Camera cam;
  int rc = make_camera(cam, H, W, K, T, max_depth);
  if (rc) return rc;
  const int n = H * W;
  if (n > map->max_points) { set_error("bnv_backproject: %d pixels exceed max_points %lld", n, (long long)map->max_points); return BNV_E_CAPACITY; }
  cudaStream_t s = (cudaStream_t)stream;
  backproject_kernel<<<(n + 127) / 128, 128, 0, s>>>(depth, cam, map->bp_pts, map->bp_flags);
  BNV_LAUNCH_CHECK("backproject_kernel");
  size_t tmp = map->cub_tmp_bytes;
  BNV_CUDA(cub::DeviceScan::ExclusiveSum(map->cub_tmp, tmp, map->bp_flags, map->bp_scan, n, s));
  count_launch(2);
  compact_pts_kernel<<<(n + 255) / 256, 256, 0, s>>>(map->bp_pts, map->bp_flags, map->bp_scan, n, pts6, n_valid);
  BNV_LAUNCH_CHECK("compact_pts_kernel");
  return BNV_OK;
}

int bnv_fuse_frame(bnv_map_t* map, const uint16_t* depth, int H, int W, const float* K, const float* T,
                   double max_depth, const bnv_mlp_t* enc, int min_pts, int mode, int64_t* frame_stats,
                   float* navg, void* stream) {
  if (!map || !depth) { set_error("bnv_fuse_frame: null argument"); return BNV_E_ARG; }
  EncSrc src{};
  int rc = make_camera(src.cam, H, W, K, T, max_depth);
  if (rc) return rc;
  src.depth = depth;
  cudaStream_t s = (cudaStream_t)stream;
  if (map->timing) BNV_CUDA(cudaEventRecord(map->ev[0], s));
  rc = launch_encode(map, src, true, (int64_t)H * W, enc, mode, s);
  if (rc) return rc;
  if (map->timing) BNV_CUDA(cudaEventRecord(map->ev[1], s));
  rc = launch_finalize(map, min_pts, mode, frame_stats, navg, s);
  if (map->timing) BNV_CUDA(cudaEventRecord(map->ev[2], s));
  return rc;
}

// host frames -> the map's staging buffer `cur` (frames back to back), unless the copy stream already brought exactly
// these frames there (prefetch hint of the previous call)
static int stage_frames(bnv_map_t* map, const uint16_t* const* depth_host, int n, size_t bytes, cudaStream_t s) {
  if (!map->copy_stream) {                       // first use: copy stream + buffer events
    BNV_CUDA(cudaStreamCreateWithFlags(&map->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      BNV_CUDA(cudaEventCreateWithFlags(&map->stage_ready[i], cudaEventDisableTiming));
      BNV_CUDA(cudaEventCreateWithFlags(&map->stage_free[i], cudaEventDisableTiming));
      BNV_CUDA(cudaEventRecord(map->stage_free[i], s));
    }
  }
  const int cur = map->stage_next;
  // a prefetch hinted at the previous call targets depth_stage[cur]: whether or not it is what is passed now,
  // nothing on `s` may read or overwrite that buffer before the copy stream's write has landed
  if (map->n_prefetched) BNV_CUDA(cudaStreamWaitEvent(s, map->stage_ready[cur], 0));
  bool hit = map->n_prefetched == n && map->prefetched_bytes == bytes;
  for (int i = 0; hit && i < n; ++i) hit = map->prefetched[i] == depth_host[i];
  if (!hit)
    for (int i = 0; i < n; ++i)
      BNV_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(map->depth_stage[cur]) + (size_t)i * bytes, depth_host[i], bytes,
                               cudaMemcpyHostToDevice, s));
  return BNV_OK;
}

// after the frames' kernels are enqueued: release the staging buffer, start the hinted copy into the other one
static int stage_next_frames(bnv_map_t* map, const uint16_t* const* next_host, int n, size_t bytes, cudaStream_t s) {
  const int cur = map->stage_next;
  BNV_CUDA(cudaEventRecord(map->stage_free[cur], s));                  // this staging buffer's readers are done
  map->n_prefetched = 0;
  map->stage_next = cur ^ 1;
  if (next_host && n > 0) {
    const int nxt = cur ^ 1;
    BNV_CUDA(cudaStreamWaitEvent(map->copy_stream, map->stage_free[nxt], 0));
    for (int i = 0; i < n; ++i) {
      BNV_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(map->depth_stage[nxt]) + (size_t)i * bytes, next_host[i], bytes,
                               cudaMemcpyHostToDevice, map->copy_stream));
      map->prefetched[i] = next_host[i];
    }
    BNV_CUDA(cudaEventRecord(map->stage_ready[nxt], map->copy_stream));
    map->n_prefetched = n;
    map->prefetched_bytes = bytes;
  }
  return BNV_OK;
}

int bnv_fuse_frame_host(bnv_map_t* map, const uint16_t* depth_host, int H, int W, const float* K, const float* T,
                        double max_depth, const bnv_mlp_t* enc, int min_pts, int mode, int64_t* frame_stats_host,
                        const uint16_t* next_depth_host, void* stream) {
  if (!map || !depth_host || H <= 0 || W <= 0) { set_error("bnv_fuse_frame_host: bad argument"); return BNV_E_ARG; }
  if ((int64_t)H * W > map->max_points) {
    set_error("bnv_fuse_frame_host: %d x %d pixels exceed the map's max_points %lld", H, W, (long long)map->max_points);
    return BNV_E_CAPACITY;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = (size_t)H * W * 2;
  int rc = stage_frames(map, &depth_host, 1, bytes, s);
  if (rc) return rc;
  rc = bnv_fuse_frame(map, map->depth_stage[map->stage_next], H, W, K, T, max_depth, enc, min_pts, mode,
                      frame_stats_host ? map->user_stats : nullptr, nullptr, stream);
  if (rc) return rc;
  if (frame_stats_host)
    BNV_CUDA(cudaMemcpyAsync(frame_stats_host, map->user_stats, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  return stage_next_frames(map, &next_depth_host, next_depth_host ? 1 : 0, bytes, s);
}

// ---- frame batches ---------------------------------------------------------------------------------------------
static int check_batch(bnv_map_t* map, int n_frames, int H, int W, const char* who) {
  if (n_frames < 1 || H <= 0 || W <= 0) { set_error("%s: bad argument", who); return BNV_E_ARG; }
  if (n_frames > map->batch_cap) {
    set_error("%s: %d frames, but the map is laid out for batches of %d (bnv_map_set_frame_batch)", who, n_frames, map->batch_cap);
    return BNV_E_CAPACITY;
  }
  if ((int64_t)H * W > (1 << 24) || (int64_t)n_frames * H * W > map->max_points) {
    set_error("%s: %d frames of %d x %d pixels exceed the map's max_points %lld", who, n_frames, H, W, (long long)map->max_points);
    return BNV_E_CAPACITY;
  }
  return BNV_OK;
}

int bnv_fuse_frames(bnv_map_t* map, const uint16_t* const* depth, int n_frames, int H, int W, const float* K, const float* T,
                    double max_depth, const bnv_mlp_t* enc, int min_pts, int mode, int64_t* batch_stats, float* navg,
                    void* stream) {
  if (!map || !depth) { set_error("bnv_fuse_frames: null argument"); return BNV_E_ARG; }
  int rc = check_batch(map, n_frames, H, W, "bnv_fuse_frames");
  if (rc) return rc;
  rc = check_encoder(enc, mode);
  if (rc) return rc;
  FrameBatch fb{};
  for (int i = 0; i < n_frames; ++i) {
    if (!depth[i]) { set_error("bnv_fuse_frames: null frame %d", i); return BNV_E_ARG; }
    rc = make_camera(fb.cam[i], H, W, K ? K + 9 * i : nullptr, T ? T + 16 * i : nullptr, max_depth);
    if (rc) return rc;
    fb.depth[i] = depth[i];
  }
  cudaStream_t s = (cudaStream_t)stream;
  // the batch's sequence number = the value its finalize swaps into the cells' lock words.  0 is the cleared word; on
  // the wrap after 2^32 batches the table (all zero between batches except for old lock words) is cleared once
  if (++map->batch_seq == 0u) {
    BNV_CUDA(cudaMemsetAsync(map->d.ftable, 0, ((size_t)map->d.g.n_vox << map->d.fshift) * 8, s));
    map->batch_seq = 1u;
  }
  if (map->timing) BNV_CUDA(cudaEventRecord(map->ev[0], s));
  const unsigned tiles = (unsigned)(((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH));
  BNV_CUDA(launch_pdl(frame_prepass_batch_kernel, dim3(tiles, (unsigned)n_frames), dim3(kPreThreads), 0, s, map->d, fb,
                      (long long*)map->stats, g_prepass_debug));
  BNV_LAUNCH_CHECK("frame_prepass_batch_kernel");
  rc = launch_encode_rows(map, (int64_t)n_frames * H * W, enc, mode, s);
  if (rc) return rc;
  if (map->timing) BNV_CUDA(cudaEventRecord(map->ev[1], s));
  {
    const dim3 grid(148 * 2), block(256);
    auto kernel = mode == BNV_MLP_TC16 ? finalize_batch_kernel<true> : finalize_batch_kernel<false>;
    BNV_CUDA(launch_pdl(kernel, grid, block, 0, s, map->d, min_pts, map->batch_seq, (long long*)map->stats,
                        (long long*)batch_stats, navg));
  }
  BNV_LAUNCH_CHECK("finalize_batch_kernel");
  if (map->timing) BNV_CUDA(cudaEventRecord(map->ev[2], s));
  return BNV_OK;
}

int bnv_fuse_frames_host(bnv_map_t* map, const uint16_t* const* depth_host, int n_frames, int H, int W, const float* K,
                         const float* T, double max_depth, const bnv_mlp_t* enc, int min_pts, int mode,
                         int64_t* batch_stats_host, const uint16_t* const* next_depth_host, int n_next, void* stream) {
  if (!map || !depth_host) { set_error("bnv_fuse_frames_host: null argument"); return BNV_E_ARG; }
  int rc = check_batch(map, n_frames, H, W, "bnv_fuse_frames_host");
  if (rc) return rc;
  if (n_next < 0 || n_next > map->batch_cap || (int64_t)n_next * H * W > map->max_points) {
    set_error("bnv_fuse_frames_host: bad prefetch hint (%d frames)", n_next);
    return BNV_E_ARG;
  }
  for (int i = 0; i < n_frames; ++i)
    if (!depth_host[i]) { set_error("bnv_fuse_frames_host: null frame %d", i); return BNV_E_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = (size_t)H * W * 2;
  rc = stage_frames(map, depth_host, n_frames, bytes, s);
  if (rc) return rc;
  const uint16_t* dev[kMaxBatch];
  for (int i = 0; i < n_frames; ++i) dev[i] = map->depth_stage[map->stage_next] + (size_t)i * H * W;
  rc = bnv_fuse_frames(map, dev, n_frames, H, W, K, T, max_depth, enc, min_pts, mode,
                       batch_stats_host ? map->user_stats : nullptr, nullptr, stream);
  if (rc) return rc;
  if (batch_stats_host)
    BNV_CUDA(cudaMemcpyAsync(batch_stats_host, map->user_stats, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  return stage_next_frames(map, next_depth_host, next_depth_host ? n_next : 0, bytes, s);
}

int bnv_fuse_points(bnv_map_t* map, const float* pts6, int64_t n_points, const bnv_mlp_t* enc, int min_pts,
                    int mode, int64_t* frame_stats, float* navg, void* stream) {
  if (!map || n_points < 0 || (n_points > 0 && !pts6)) { set_error("bnv_fuse_points: bad argument"); return BNV_E_ARG; }
  EncSrc src{};
  src.pts6 = pts6;
  src.n_points = n_points;
  cudaStream_t s = (cudaStream_t)stream;
  if (map->timing) BNV_CUDA(cudaEventRecord(map->ev[0], s));
  int rc = launch_encode(map, src, false, n_points, enc, mode, s);
  if (rc) return rc;
  if (map->timing) BNV_CUDA(cudaEventRecord(map->ev[1], s));
  rc = launch_finalize(map, min_pts, mode, frame_stats, navg, s);
  if (map->timing) BNV_CUDA(cudaEventRecord(map->ev[2], s));
  return rc;
}

int bnv_encode_points(bnv_map_t* map, const float* pts6, int64_t n_points, const bnv_mlp_t* enc, int min_pts,
                      int mode, float* feats, int64_t* counts, int64_t* flat_ids, int64_t* coords,
                      int64_t out_capacity, int64_t* stats_dev, float* navg_dev, void* stream) {
  if (!map || n_points < 0 || !stats_dev || (n_points > 0 && !pts6) ||
      (out_capacity > 0 && (!feats || !counts || !flat_ids || !coords))) {
    set_error("bnv_encode_points: bad argument");
    return BNV_E_ARG;
  }
  EncSrc src{};
  src.pts6 = pts6;
  src.n_points = n_points;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_encode(map, src, false, n_points, enc, mode, s);
  if (rc) return rc;
  // The reference synchronises here as well (torch.unique's output size, `int(v) for v in n_xyz`,
  // local_point_fusion.py:90,118): the sorted path reads the touched-voxel count to size the sort.
  int32_t c[4];
  BNV_CUDA(cudaMemcpyAsync(c, map->d.ctr, sizeof(c), cudaMemcpyDeviceToHost, s));
  BNV_CUDA(cudaStreamSynchronize(s));
  const int n = c[1];
  if (n > 0) {
    const unsigned nb = (unsigned)((n + 255) / 256);
    sort_prep_kernel<<<nb, 256, 0, s>>>(map->d, n, map->sort_keys_in, map->sort_vals_in);
    BNV_LAUNCH_CHECK("sort_prep_kernel");
    size_t tmp = map->cub_tmp_bytes;
    int end_bit = 1;
    while (end_bit < 31 && (1ll << end_bit) < map->d.g.n_vox) ++end_bit;
    BNV_CUDA(cub::DeviceRadixSort::SortPairs(map->cub_tmp, tmp, map->sort_keys_in, map->sort_keys_out,
                                             map->sort_vals_in, map->sort_vals_out, n, 0, end_bit, s));
    count_launch(4);
    sort_flag_kernel<<<nb, 256, 0, s>>>(map->d, n, map->sort_keys_out, min_pts, map->flags);
    BNV_LAUNCH_CHECK("sort_flag_kernel");
    tmp = map->cub_tmp_bytes;
    BNV_CUDA(cub::DeviceScan::ExclusiveSum(map->cub_tmp, tmp, map->flags, map->scan, n, s));
    count_launch(2);
    sort_emit_kernel<<<(unsigned)(((int64_t)n * 8 + 255) / 256), 256, 0, s>>>(
        map->d, n, map->sort_keys_out, map->sort_vals_out, map->flags, map->scan, mode == BNV_MLP_TC16, out_capacity, feats, counts,
        flat_ids, coords);
    BNV_LAUNCH_CHECK("sort_emit_kernel");
  }
  sort_publish_kernel<<<1, 1, 0, s>>>(map->d, n, map->flags, map->scan, (long long*)map->stats, stats_dev, navg_dev);
  BNV_LAUNCH_CHECK("sort_publish_kernel");
  if (n > 0) {
    // flags are reused by count_optim as a zero-initialised mark array
    BNV_CUDA(cudaMemsetAsync(map->flags, 0, (size_t)n * 4, s));
  }
  return BNV_OK;
}

int bnv_integrate(bnv_map_t* map, const int64_t* coords, const float* feats, const int64_t* counts, int64_t n,
                  void* stream) {
  if (!map || n < 0 || (n > 0 && (!coords || !feats || !counts))) { set_error("bnv_integrate: bad argument"); return BNV_E_ARG; }
  if (n == 0) return BNV_OK;
  integrate_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(map->d, coords, feats, counts, n);
  BNV_LAUNCH_CHECK("integrate_kernel");
  return BNV_OK;
}

}  // extern "C"
