// Sparse voxel map (reference: SparseVolume over open3d.core.HashMap,
// src/models/sparse_volume.py:484-695) + error plumbing + MLP weight handles.
//
// B200-first layout: the reference's voxel key is an int32 flat id < 2^31 by construction
// (voxel_utils.flatten, src/utils/voxel_utils.py:62-65), so instead of a probing hash table built
// for 8-24 GB GPUs the map keeps an identity-hashed slot table `table[flat] -> slot` resident in
// HBM (4 B per grid cell: 537 MB for the 512^3 grid of the headline workload, 8.6 GB worst case --
// <5 % of a B200's 180 GB) and a dense SoA value pool in slot (= activation) order.  A lookup is one
// 4-byte read, no probing and no CAS loop; to_tensor() is a contiguous copy of the pool prefix.
#include <stdarg.h>

#include <atomic>
#include <vector>

#include "bnv_common.cuh"

namespace bnv {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return BNV_E_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------- //
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // 16-byte stores over the bulk (pointers from cudaMalloc are 256-byte aligned)
  int4* p4 = reinterpret_cast<int4*>(p);
  const int64_t n4 = n / 4;
  const int4 v4 = make_int4(v, v, v, v);
  for (int64_t j = i; j < n4; j += stride) p4[j] = v4;
  for (int64_t j = n4 * 4 + i; j < n; j += stride) p[j] = v;
}

static int fill_i32(int32_t* p, int64_t n, int32_t v, cudaStream_t s) {
  if (n == 0) return BNV_OK;
  fill_i32_kernel<<<148 * 8, 256, 0, s>>>(p, n, v);
  BNV_LAUNCH_CHECK("fill_i32_kernel");
  return BNV_OK;
}

__device__ __forceinline__ bool key_to_flat(const GeomDev& g, long long x, long long y, long long z,
                                            int32_t& flat) {
  if (x < 0 || y < 0 || z < 0 || x >= g.n[0] || y >= g.n[1] || z >= g.n[2]) return false;
  flat = (int32_t)x * g.nyz + (int32_t)y * g.n[2] + (int32_t)z;
  return true;
}

// SparseVolume.query (sparse_volume.py:661-695): 8 lanes per key, one feature each.
__global__ void map_query_kernel(MapDev m, const int64_t* __restrict__ coords, int64_t n,
                                 float* __restrict__ feats, float* __restrict__ weights,
                                 float* __restrict__ hits, uint8_t* __restrict__ found) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t >> 3;
  const int j = (int)(t & 7);
  if (i >= n) return;
  int32_t flat;
  int32_t slot = kEmpty;
  if (key_to_flat(m.g, coords[i * 3], coords[i * 3 + 1], coords[i * 3 + 2], flat)) slot = m.table[flat];
  const bool hit = slot >= 0;
  feats[i * kFeat + j] = hit ? m.feats[(int64_t)slot * kFeat + j] : 0.f;
  if (j == 0) {
    weights[i] = hit ? m.weights[slot] : 0.f;
    hits[i] = hit ? m.hits[slot] : 0.f;
    if (found) found[i] = hit ? 1 : 0;
  }
}

// SparseVolume.insert (sparse_volume.py:561-585): upsert.  A key absent from the table is claimed
// with a unique negative tag (-(i+2)); the claimant allocates the slot.  Another row of the same
// call that carries the same key sees the tag and yields (one of the duplicates wins, as in o3c).
__global__ void map_insert_kernel(MapDev m, const int64_t* __restrict__ coords,
                                  const float* __restrict__ feats, const float* __restrict__ weights,
                                  const float* __restrict__ hits, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t flat;
  if (!key_to_flat(m.g, coords[i * 3], coords[i * 3 + 1], coords[i * 3 + 2], flat)) {
    atomicOr(&m.ctr[2], kErrRange);
    return;
  }
  int32_t slot = m.table[flat];
  if (slot == kEmpty) {
    const int32_t tag = -(int32_t)(i % 0x3fffffff) - 2;
    const int32_t old = atomicCAS(&m.table[flat], kEmpty, tag);
    if (old == kEmpty) {
      slot = atomicAdd(&m.ctr[0], 1);
      if (slot >= m.cap) {
        atomicOr(&m.ctr[2], kErrCapacity);
        atomicExch(&m.table[flat], kEmpty);
        return;
      }
      m.keys[slot] = flat;
      atomicExch(&m.table[flat], slot);
    } else {
      slot = old;
    }
  }
  if (slot < 0) return;  // duplicate key inside this call: the claimant's values stand
  const float4* f = reinterpret_cast<const float4*>(feats + i * kFeat);
  float4* o = reinterpret_cast<float4*>(m.feats + (int64_t)slot * kFeat);
  o[0] = f[0];
  o[1] = f[1];
  m.weights[slot] = weights[i];
  m.hits[slot] = hits[i];
}

__global__ void map_export_kernel(MapDev m, int64_t n, int64_t* __restrict__ coords,
                                  float* __restrict__ feats, float* __restrict__ weights,
                                  float* __restrict__ hits) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t >> 3;
  const int j = (int)(t & 7);
  if (i >= n) return;
  feats[i * kFeat + j] = m.feats[i * kFeat + j];
  if (j == 0) {
    weights[i] = m.weights[i];
    hits[i] = m.hits[i];
    const int32_t flat = m.keys[i];
    const int32_t x = flat / m.g.nyz;
    const int32_t r = flat - x * m.g.nyz;
    const int32_t y = r / m.g.n[2];
    coords[i * 3 + 0] = x;
    coords[i * 3 + 1] = y;
    coords[i * 3 + 2] = r - y * m.g.n[2];
  }
}

// count_optim (sparse_volume.py:602-622): weights[rows(keys)] += 1, once per distinct row.
__global__ void count_optim_mark_kernel(MapDev m, const float* __restrict__ nbr, int64_t n,
                                        int32_t* __restrict__ flags, int64_t n_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t flat;
  if (!key_to_flat(m.g, (long long)nbr[i * 3], (long long)nbr[i * 3 + 1], (long long)nbr[i * 3 + 2], flat))
    return;
  const int32_t slot = m.table[flat];
  if (slot >= 0 && slot < n_rows) flags[slot] = 1;
}

// the same for the 8 floor/ceil corners of query points (render_with_rays, src/utils/render_utils.py:494-496:
// coords = (pts - min_coords) / voxel_size; get_neighbors; count_optim) without materialising the [8 Q, 3] corner tensor
__global__ void count_optim_mark_queries_kernel(MapDev m, const float* __restrict__ q, int64_t n, int is_coords,
                                                int32_t* __restrict__ flags, int64_t n_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float v = q[i * 3 + a];
    const float c = is_coords ? v : __fmul_rn(__fsub_rn(v, m.g.bmin[a]), m.g.inv_vs);
    lo[a] = (long long)floorf(c);
    hi[a] = (long long)ceilf(c);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int32_t flat;
    if (!key_to_flat(m.g, (k & 1) ? hi[0] : lo[0], (k & 2) ? hi[1] : lo[1], (k & 4) ? hi[2] : lo[2], flat)) continue;
    const int32_t slot = m.table[flat];
    if (slot >= 0 && slot < n_rows) flags[slot] = 1;
  }
}

__global__ void count_optim_apply_kernel(int32_t* __restrict__ flags, float* __restrict__ w, int64_t n_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  if (flags[i]) {
    w[i] += 1.f;
    flags[i] = 0;
  }
}

// boundary exchange, sender side: one record {flat id, weight, feat[8]} with the CURRENT values of every shell voxel
// integrated since the last exchange (the dirty list holds each such slot once), flags cleared on the way
__global__ void __launch_bounds__(256) halo_pack_kernel(MapDev m, int32_t* __restrict__ buf, int cap) {
  const int n = min(min(m.ctr[5], m.dirty_cap), cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int32_t slot = m.dirty_list[i];
    m.dirty_flag[slot] = 0;
    const float4* f4 = reinterpret_cast<const float4*>(m.feats + (size_t)slot * kFeat);
    const float4 a = f4[0], b = f4[1];
    int32_t* rec = buf + 10 + (size_t)i * 10;                   // 40-byte records: 8-byte aligned
    reinterpret_cast<int2*>(rec)[0] = make_int2(m.keys[slot], __float_as_int(m.weights[slot]));
    reinterpret_cast<float2*>(rec)[1] = make_float2(a.x, a.y);
    reinterpret_cast<float2*>(rec)[2] = make_float2(a.z, a.w);
    reinterpret_cast<float2*>(rec)[3] = make_float2(b.x, b.y);
    reinterpret_cast<float2*>(rec)[4] = make_float2(b.z, b.w);
  }
}
__global__ void halo_pack_reset_kernel(MapDev m, int32_t* __restrict__ buf, int cap) {
  if (threadIdx.x == 0) {
    const int n = m.ctr[5];
    if (n > m.dirty_cap || n > cap) atomicOr(&m.ctr[2], kErrCapacity);    // records were dropped
    buf[0] = min(min(n, m.dirty_cap), cap);
    m.ctr[5] = 0;
  }
}

// upsert the halo records of the other ranks that this rank needs: 8 lanes per record
__global__ void insert_halo_kernel(MapDev m, const int32_t* __restrict__ gathered, int world, int64_t cap) {
  const int64_t stride_words = 10 + cap * 10;
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
  const int lane8 = threadIdx.x & 7;
  for (int r = 0; r < world; ++r) {
    if (r == m.g.rank) continue;
    const int32_t* buf = gathered + r * stride_words;
    const int n = min(buf[0], (int)cap);
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < n;
         i += ((int64_t)gridDim.x * blockDim.x) >> 3) {
      const int32_t* rec = buf + 10 + i * 10;
      const int32_t flat = rec[0];
      const int x = flat / m.g.nyz, rr = flat - x * m.g.nyz, y = rr / m.g.n[2], z = rr - y * m.g.n[2];
      const bool need = rank_touches(m.g, x, y, z, m.g.rank);
      int32_t slot = -1;
      if (need && lane8 == 0) {
        slot = m.table[flat];
        if (slot < 0) {                      // keys are unique within one rank's buffer and across ranks
          slot = atomicAdd(&m.ctr[0], 1);
          if (slot < m.cap) {
            m.table[flat] = slot;
            m.keys[slot] = flat;
            m.hits[slot] = 0.f;
          } else {
            atomicOr(&m.ctr[2], kErrCapacity);
            slot = -1;
          }
        }
      }
      slot = __shfl_sync(gmask, slot, (threadIdx.x & 31) & ~7);
      if (slot >= 0) {
        m.feats[(size_t)slot * kFeat + lane8] = reinterpret_cast<const float*>(rec)[2 + lane8];
        if (lane8 == 0) m.weights[slot] = reinterpret_cast<const float*>(rec)[1];
      }
    }
  }
}

}  // namespace bnv

using namespace bnv;

extern "C" {

int bnv_map_halo_enable(bnv_map_t* m, int64_t capacity_records) {
  if (!m || capacity_records < 0 || capacity_records > 0x7fffffff) { set_error("bnv_map_halo_enable: bad argument"); return BNV_E_ARG; }
  BNV_CUDA(cudaSetDevice(m->device));
  BNV_CUDA(cudaDeviceSynchronize());
  if (m->d.dirty_flag) cudaFree(m->d.dirty_flag);
  if (m->d.dirty_list) cudaFree(m->d.dirty_list);
  m->d.dirty_flag = m->d.dirty_list = nullptr;
  m->d.dirty_cap = 0;
  if (capacity_records == 0) return BNV_OK;
  BNV_CUDA(cudaMalloc((void**)&m->d.dirty_flag, (size_t)m->d.cap * 4));
  BNV_CUDA(cudaMalloc((void**)&m->d.dirty_list, (size_t)capacity_records * 4));
  BNV_CUDA(cudaMemset(m->d.dirty_flag, 0, (size_t)m->d.cap * 4));
  BNV_CUDA(cudaMemset(m->d.ctr + 5, 0, 4));
  m->d.dirty_cap = (int32_t)capacity_records;
  return BNV_OK;
}

int bnv_map_halo_pack(bnv_map_t* m, void* buf, int64_t capacity_records, void* stream) {
  if (!m || !buf || capacity_records <= 0 || !m->d.dirty_list) { set_error("bnv_map_halo_pack: bad argument (bnv_map_halo_enable first)"); return BNV_E_ARG; }
  halo_pack_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(m->d, (int32_t*)buf, (int)(capacity_records < 0x7fffffff ? capacity_records : 0x7fffffff));
  BNV_LAUNCH_CHECK("halo_pack_kernel");
  halo_pack_reset_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(m->d, (int32_t*)buf, (int)(capacity_records < 0x7fffffff ? capacity_records : 0x7fffffff));
  BNV_LAUNCH_CHECK("halo_pack_reset_kernel");
  return BNV_OK;
}

int bnv_map_insert_halo(bnv_map_t* m, const void* gathered, int world, int64_t cap, void* stream) {
  if (!m || !gathered || world != m->d.g.world || cap <= 0) { set_error("bnv_map_insert_halo: bad argument"); return BNV_E_ARG; }
  insert_halo_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(m->d, (const int32_t*)gathered, world, cap);
  BNV_LAUNCH_CHECK("insert_halo_kernel");
  return BNV_OK;
}

int bnv_abi_version(void) { return BNV_ABI_VERSION; }
const char* bnv_last_error(void) { return g_err; }
int64_t bnv_launch_count(void) { return (int64_t)g_launches.load(); }

int bnv_map_create(bnv_map_t** out, const bnv_geom_t* geom, int n_feats, int64_t capacity,
                   int64_t max_points, int device) {
  if (!out || !geom) { set_error("bnv_map_create: null argument"); return BNV_E_ARG; }
  if (n_feats != kFeat) { set_error("bnv_map_create: n_feats must be %d (got %d)", kFeat, n_feats); return BNV_E_UNSUPPORTED; }
  const int64_t nx = geom->n_xyz[0], ny = geom->n_xyz[1], nz = geom->n_xyz[2];
  if (nx <= 0 || ny <= 0 || nz <= 0 || nx * ny * nz >= (1ll << 31)) {
    set_error("bnv_map_create: grid %lld x %lld x %lld does not fit the reference's int32 flat id",
              (long long)nx, (long long)ny, (long long)nz);
    return BNV_E_ARG;
  }
  if (capacity <= 0 || capacity >= (1ll << 31) || max_points <= 0 || max_points * 8 >= (1ll << 31)) {
    set_error("bnv_map_create: bad capacity %lld / max_points %lld", (long long)capacity, (long long)max_points);
    return BNV_E_ARG;
  }
  BNV_CUDA(cudaSetDevice(device));
  bnv_map* m = new bnv_map();
  memset(m, 0, sizeof(*m));
  m->device = device;
  m->max_points = max_points;
  GeomDev& g = m->d.g;
  const float vs = (float)geom->voxel_size;
  for (int a = 0; a < 3; ++a) {
    g.bmin[a] = geom->bmin[a];
    g.lo[a] = geom->bmin[a] + vs;     // fp32, like `bound_min[i] + voxel_size` on a float tensor
    g.hi[a] = geom->bmax[a] - vs;
    g.n[a] = geom->n_xyz[a];
  }
  g.vs = vs;
  g.inv_vs = 1.0f / vs;
  g.nyz = (int32_t)(ny * nz);
  g.n_vox = nx * ny * nz;
  g.rank = 0; g.world = 1; g.brick_log2 = 4;
  MapDev& d = m->d;
  d.cap = (int32_t)capacity;
  d.fcap = (int32_t)(max_points * 8 < g.n_vox ? max_points * 8 : g.n_vox);   // a frame cannot touch more voxels
  const int64_t aux = d.cap > d.fcap ? d.cap : d.fcap;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  alloc((void**)&d.table, g.n_vox * 4);
  alloc((void**)&d.ftable, g.n_vox * 8);
  alloc((void**)&d.ftable_dummy, 1024 * 8);
  alloc((void**)&d.keys, (size_t)d.cap * 4);
  alloc((void**)&d.feats, (size_t)d.cap * kFeat * 4);
  alloc((void**)&d.weights, (size_t)d.cap * 4);
  alloc((void**)&d.hits, (size_t)d.cap * 4);
  alloc((void**)&d.fkeys, (size_t)d.fcap * 4);
  alloc((void**)&d.fsum, (size_t)d.fcap * kFeat * 8);
  alloc((void**)&d.prec, (size_t)max_points * 8 * 4);
  alloc((void**)&d.ctr, 16 * 4);
  alloc((void**)&m->sort_keys_in, (size_t)d.fcap * 4);
  alloc((void**)&m->sort_keys_out, (size_t)d.fcap * 4);
  alloc((void**)&m->sort_vals_in, (size_t)d.fcap * 4);
  alloc((void**)&m->sort_vals_out, (size_t)d.fcap * 4);
  alloc((void**)&m->flags, (size_t)aux * 4);
  alloc((void**)&m->scan, (size_t)aux * 4);
  alloc((void**)&m->depth_stage[0], (size_t)max_points * 2);
  alloc((void**)&m->depth_stage[1], (size_t)max_points * 2);
  alloc((void**)&m->user_stats, 4 * 8);
  alloc((void**)&m->bp_pts, (size_t)max_points * 6 * 4);
  alloc((void**)&m->bp_flags, (size_t)max_points * 4);
  alloc((void**)&m->bp_scan, (size_t)max_points * 4);
  alloc((void**)&m->stats, 8 * 8);
  alloc((void**)&m->dec_pack, (size_t)d.cap * 16);
  m->gtable_bytes = ((size_t)d.cap + 1) * 27 * sizeof(float);
  alloc((void**)&m->gtable, m->gtable_bytes);
  m->cub_tmp_bytes = (size_t)d.fcap * 16 + (1 << 20);
  alloc((void**)&m->cub_tmp, m->cub_tmp_bytes);
  if (e != cudaSuccess) {
    bnv_map_destroy(m);
    set_error("bnv_map_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    return BNV_E_ALLOC;
  }
  *out = m;
  int rc = bnv_map_reset(m, nullptr);
  if (rc != BNV_OK) return rc;
  BNV_CUDA(cudaMemsetAsync(d.fsum, 0, (size_t)d.fcap * kFeat * 8, 0));
  BNV_CUDA(cudaMemsetAsync(m->flags, 0, (size_t)aux * 4, 0));
  BNV_CUDA(cudaMemsetAsync(m->stats, 0, 64, 0));
  BNV_CUDA(cudaMemsetAsync(d.ftable, 0, (size_t)g.n_vox * 8, 0));
  BNV_CUDA(cudaStreamSynchronize(0));
  return BNV_OK;
}

int bnv_map_set_frame_batch(bnv_map_t* m, int n_frames) {
  if (!m || n_frames < 0 || n_frames > kMaxBatch) {
    set_error("bnv_map_set_frame_batch: frames per batch must be in [0, %d]", kMaxBatch);
    return BNV_E_ARG;
  }
  BNV_CUDA(cudaSetDevice(m->device));
  BNV_CUDA(cudaDeviceSynchronize());             // no frame may be in flight while the per-frame table is replaced
  MapDev& d = m->d;
  const int shift = n_frames == 0 ? 0 : 3;       // 8 table words per grid cell: kMaxBatch frames + the finalize lock
  const int64_t pairs = d.g.n_vox * (n_frames > 1 ? n_frames : 1);       // (frame, voxel) pairs a call can touch
  const int64_t fcap = m->max_points * 8 < pairs ? m->max_points * 8 : pairs;
  if (shift != d.fshift) {
    BNV_CUDA(cudaFree(d.ftable));
    d.ftable = nullptr;
    d.fshift = 0;
    m->batch_cap = 0;
    const size_t bytes = ((size_t)d.g.n_vox << shift) * 8;
    cudaError_t e = cudaMalloc((void**)&d.ftable, bytes);
    if (e != cudaSuccess) {                      // fall back to the single-frame layout so that the map stays usable
      cudaGetLastError();
      BNV_CUDA(cudaMalloc((void**)&d.ftable, (size_t)d.g.n_vox * 8));
      BNV_CUDA(cudaMemset(d.ftable, 0, (size_t)d.g.n_vox * 8));
      set_error("bnv_map_set_frame_batch: no memory for %zu bytes of per-frame table", bytes);
      return BNV_E_ALLOC;
    }
    BNV_CUDA(cudaMemset(d.ftable, 0, bytes));
    d.fshift = shift;
    m->batch_seq = 0;
  }
  if (fcap > d.fcap) {                           // scratch rows for every (frame, voxel) pair of a batch
    int32_t* fkeys = nullptr;
    long long* fsum = nullptr;
    cudaError_t e = cudaMalloc((void**)&fkeys, (size_t)fcap * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&fsum, (size_t)fcap * kFeat * 8);
    if (e != cudaSuccess) {
      cudaGetLastError();
      if (fkeys) cudaFree(fkeys);
      set_error("bnv_map_set_frame_batch: no memory for %lld scratch rows", (long long)fcap);
      return BNV_E_ALLOC;
    }
    BNV_CUDA(cudaMemset(fsum, 0, (size_t)fcap * kFeat * 8));
    cudaFree(d.fkeys);
    cudaFree(d.fsum);
    d.fkeys = fkeys;
    d.fsum = fsum;
    d.fcap = (int32_t)fcap;
  }
  m->batch_cap = n_frames;
  return BNV_OK;
}

int bnv_map_destroy(bnv_map_t* m) {
  if (!m) return BNV_OK;
  cudaSetDevice(m->device);
  MapDev& d = m->d;
  void* ptrs[] = {d.table, d.ftable, d.keys, d.feats, d.weights, d.hits, d.fkeys, d.fsum, d.prec, d.ftable_dummy, d.dirty_flag, d.dirty_list,
                  d.ctr, m->sort_keys_in, m->sort_keys_out, m->sort_vals_in,
                  m->sort_vals_out, m->flags, m->scan, m->depth_stage[0], m->depth_stage[1], m->user_stats, m->bp_pts, m->bp_flags, m->bp_scan, m->stats, m->dec_pack, m->gtable,
                  m->cub_tmp};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int i = 0; i < 4; ++i) if (m->ev[i]) cudaEventDestroy(m->ev[i]);
  for (int i = 0; i < 2; ++i) {
    if (m->stage_ready[i]) cudaEventDestroy(m->stage_ready[i]);
    if (m->stage_free[i]) cudaEventDestroy(m->stage_free[i]);
  }
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  delete m;
  return BNV_OK;
}

int bnv_map_reset(bnv_map_t* m, void* stream) {
  if (!m) { set_error("bnv_map_reset: null map"); return BNV_E_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  BNV_CUDA(cudaSetDevice(m->device));
  int rc = fill_i32(m->d.table, m->d.g.n_vox, kEmpty, s);
  if (rc != BNV_OK) return rc;
  BNV_CUDA(cudaMemsetAsync(m->d.ctr, 0, 64, s));
  if (m->d.dirty_flag) BNV_CUDA(cudaMemsetAsync(m->d.dirty_flag, 0, (size_t)m->d.cap * 4, s));
  return BNV_OK;
}

// latched device-side faults -> error code + message
static int status_rc(const bnv_map_t* m, int32_t bits) {
  if (bits & kErrCapacity) { set_error("voxel map capacity exceeded (pool capacity %d voxels, boundary-exchange capacity %d records): voxels were dropped", m->d.cap, m->d.dirty_cap); return BNV_E_CAPACITY; }
  if (bits & kErrRange) { set_error("voxel key outside the %d x %d x %d grid", m->d.g.n[0], m->d.g.n[1], m->d.g.n[2]); return BNV_E_RANGE; }
  if (bits & kErrExchange) { set_error("peer-memory halo exchange timed out waiting for another rank"); return BNV_E_CUDA; }
  return BNV_OK;
}

int bnv_map_size(bnv_map_t* m, int64_t* n_active_host, void* stream) {
  if (!m || !n_active_host) { set_error("bnv_map_size: null argument"); return BNV_E_ARG; }
  int32_t c[4];
  BNV_CUDA(cudaMemcpyAsync(c, m->d.ctr, sizeof(c), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  BNV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  *n_active_host = c[0] < m->d.cap ? c[0] : m->d.cap;
  // every host-side reader of the map (len / to_tensor / save / load) comes through here: a latched fault means
  // voxels were dropped, so the size call fails loudly instead of handing out a truncated map
  return status_rc(m, c[2]);
}

int bnv_map_status(bnv_map_t* m, void* stream) {
  if (!m) { set_error("bnv_map_status: null map"); return BNV_E_ARG; }
  int32_t c[4];
  BNV_CUDA(cudaMemcpyAsync(c, m->d.ctr, sizeof(c), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  BNV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return status_rc(m, c[2]);
}

int bnv_map_set_shard(bnv_map_t* m, int rank, int world, int brick_log2) {
  if (!m || world < 1 || rank < 0 || rank >= world || brick_log2 < 0 || brick_log2 > 20) {
    set_error("bnv_map_set_shard: bad arguments rank=%d world=%d brick_log2=%d", rank, world, brick_log2);
    return BNV_E_ARG;
  }
  m->d.g.rank = rank; m->d.g.world = world; m->d.g.brick_log2 = brick_log2;
  return BNV_OK;
}

int bnv_map_set_timing(bnv_map_t* m, int enable) {
  if (!m) { set_error("bnv_map_set_timing: null map"); return BNV_E_ARG; }
  BNV_CUDA(cudaSetDevice(m->device));
  if (enable && !m->ev[0])
    for (int i = 0; i < 4; ++i) BNV_CUDA(cudaEventCreate(&m->ev[i]));
  m->timing = enable ? 1 : 0;
  return BNV_OK;
}

int bnv_map_get_timing(bnv_map_t* m, float* enc_ms, float* fin_ms) {
  if (!m || !m->ev[0] || !enc_ms || !fin_ms) { set_error("bnv_map_get_timing: timing was never enabled"); return BNV_E_ARG; }
  BNV_CUDA(cudaEventSynchronize(m->ev[2]));
  BNV_CUDA(cudaEventElapsedTime(enc_ms, m->ev[0], m->ev[1]));
  BNV_CUDA(cudaEventElapsedTime(fin_ms, m->ev[1], m->ev[2]));
  return BNV_OK;
}

int bnv_map_get_timing_stages(bnv_map_t* m, float* ms3) {
  if (!m || !m->ev[0] || !ms3) { set_error("bnv_map_get_timing_stages: timing was never enabled"); return BNV_E_ARG; }
  BNV_CUDA(cudaEventSynchronize(m->ev[2]));
  BNV_CUDA(cudaEventElapsedTime(ms3 + 0, m->ev[0], m->ev[3]));
  BNV_CUDA(cudaEventElapsedTime(ms3 + 1, m->ev[3], m->ev[1]));
  BNV_CUDA(cudaEventElapsedTime(ms3 + 2, m->ev[1], m->ev[2]));
  return BNV_OK;
}

int bnv_map_query(bnv_map_t* m, const int64_t* coords, int64_t n, float* feats, float* weights,
                  float* hits, uint8_t* found, void* stream) {
  if (!m || n < 0 || (n > 0 && (!coords || !feats || !weights || !hits))) { set_error("bnv_map_query: bad argument"); return BNV_E_ARG; }
  if (n == 0) return BNV_OK;
  const int64_t threads = n * 8;
  map_query_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m->d, coords, n, feats, weights, hits, found);
  BNV_LAUNCH_CHECK("map_query_kernel");
  return BNV_OK;
}

int bnv_map_insert(bnv_map_t* m, const int64_t* coords, const float* feats, const float* weights,
                   const float* hits, int64_t n, void* stream) {
  if (!m || n < 0 || (n > 0 && (!coords || !feats || !weights || !hits))) { set_error("bnv_map_insert: bad argument"); return BNV_E_ARG; }
  if (n == 0) return BNV_OK;  // sparse_volume.py:570-571
  map_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m->d, coords, feats, weights, hits, n);
  BNV_LAUNCH_CHECK("map_insert_kernel");
  return BNV_OK;
}

int bnv_map_export(bnv_map_t* m, int64_t n, int64_t* coords, float* feats, float* weights,
                   float* hits, void* stream) {
  if (!m || n < 0 || n > m->d.cap || (n > 0 && (!coords || !feats || !weights || !hits))) { set_error("bnv_map_export: bad argument"); return BNV_E_ARG; }
  if (n == 0) return BNV_OK;
  map_export_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m->d, n, coords, feats, weights, hits);
  BNV_LAUNCH_CHECK("map_export_kernel");
  return BNV_OK;
}

int bnv_map_count_optim_queries(bnv_map_t* m, const float* coords, int64_t n, int is_coords, float* weights_rows,
                                int64_t n_rows, void* stream) {
  if (!m || n < 0 || n_rows < 0 || n_rows > m->d.cap || (n > 0 && (!coords || !weights_rows))) { set_error("bnv_map_count_optim_queries: bad argument"); return BNV_E_ARG; }
  if (n == 0 || n_rows == 0) return BNV_OK;
  count_optim_mark_queries_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m->d, coords, n, is_coords, m->flags, n_rows);
  BNV_LAUNCH_CHECK("count_optim_mark_queries_kernel");
  count_optim_apply_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m->flags, weights_rows, n_rows);
  BNV_LAUNCH_CHECK("count_optim_apply_kernel");
  return BNV_OK;
}

int bnv_map_count_optim(bnv_map_t* m, const float* nbr, int64_t n, float* weights_rows, int64_t n_rows,
                        void* stream) {
  if (!m || n < 0 || n_rows < 0 || n_rows > m->d.cap || (n > 0 && (!nbr || !weights_rows))) { set_error("bnv_map_count_optim: bad argument"); return BNV_E_ARG; }
  if (n == 0 || n_rows == 0) return BNV_OK;
  count_optim_mark_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m->d, nbr, n, m->flags, n_rows);
  BNV_LAUNCH_CHECK("count_optim_mark_kernel");
  count_optim_apply_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m->flags, weights_rows, n_rows);
  BNV_LAUNCH_CHECK("count_optim_apply_kernel");
  return BNV_OK;
}

}  // extern "C"
