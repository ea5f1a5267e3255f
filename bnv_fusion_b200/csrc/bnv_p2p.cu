// Peer-memory halo exchange of the tile shard (EXPERIMENTAL, opt-in: TileShardedFusion(exchange="p2p")).
//
// The default exchange is one NCCL all-gather of every rank's boundary records per frame followed by an
// upsert kernel that discards what the rank does not need (bnv_map_insert_halo).  Here the sender does the
// routing: after the frame's finalize, `halo_push_kernel` stores each boundary record straight into the
// inbox of exactly the ranks that own a brick touching the voxel (st.global over NVLink into cudaIpc-mapped
// peer memory), then publishes per-peer counts and a frame sequence number with system-scope release
// stores.  The receiver's side stream waits for the sequence numbers of all peers (one spinning warp with a
// timeout, it holds no SM that the fusion kernels need) and upserts its inbox.  No collective kernel, no
// all-to-all traffic, and the whole sharded step is reachable from one C call.
//
// Everything but one event record / wait runs on the exchange's side stream, overlapped with the next frame.  The
// local record buffers and the peers' inboxes are double-buffered by frame parity; a sender reuses parity b for
// frame s only after every peer acknowledged frame s - 2 (ack sequence numbers written back the same way).
//
// The exchange is per EPOCH, not per frame: finalize remembers every shell voxel it integrates once (dirty list,
// bnv_map.cu), and an epoch (every K frames, or when the map is read) packs their current values and routes them.
#include <new>

#include "bnv_common.cuh"

namespace bnv {

constexpr int kMaxPeers = 16;
constexpr int kRecWords = 10;                 // int32 flat_id, float weight, float feat[8]
constexpr int kHdrWords = 64;                 // ready[16] | ack[16] | count[2][16]

struct PeerPtrs {
  int32_t* base[kMaxPeers];                   // base of every rank's exchange block (own rank: local pointer)
};

// layout of one rank's exchange block (int32 words)
__host__ __device__ inline int64_t ex_block_words(int world, int64_t cap) { return kHdrWords + 2 * (int64_t)world * cap * kRecWords; }
__device__ __forceinline__ uint32_t* ex_ready(int32_t* base) { return reinterpret_cast<uint32_t*>(base); }
__device__ __forceinline__ uint32_t* ex_ack(int32_t* base) { return reinterpret_cast<uint32_t*>(base) + kMaxPeers; }
__device__ __forceinline__ int32_t* ex_count(int32_t* base, int buf) { return base + 2 * kMaxPeers + buf * kMaxPeers; }
__device__ __forceinline__ int32_t* ex_inbox(int32_t* base, int buf, int world, int64_t cap, int from) {
  return base + kHdrWords + ((int64_t)buf * world + from) * cap * kRecWords;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// one warp: spin until flags[r] >= want for every r != rank (bounded; a timeout latches kErrExchange in the map status)
__global__ void wait_flags_kernel(const uint32_t* __restrict__ flags, uint32_t want, int world, int rank,
                                  long long timeout_cycles, int32_t* __restrict__ status) {
  const int r = threadIdx.x;
  if (r >= world || r == rank) return;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flags + r) - want) < 0) {
    if (clock64() - t0 > timeout_cycles) {
      atomicOr(status, kErrExchange);
      return;
    }
    __nanosleep(200);
  }
}

// sender: route this frame's boundary records to the ranks that need them
__global__ void __launch_bounds__(256) halo_push_kernel(MapDev m, const int32_t* __restrict__ packed /* bnv_map_halo_pack */,
                                                        PeerPtrs peers, int buf, uint32_t seq, int64_t cap,
                                                        int32_t* __restrict__ sent /*[world]*/, int32_t* __restrict__ done) {
  const GeomDev& g = m.g;
  const int n = packed[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t* rec = packed + 10 + i * kRecWords;
    const int32_t flat = rec[0];
    const int x = flat / g.nyz, rr = flat - x * g.nyz, y = rr / g.n[2], z = rr - y * g.n[2];
    uint32_t need = 0;                        // ranks owning a brick in the voxel's 26-neighbourhood
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz) {
          const int u = x + dx, v = y + dy, w = z + dz;
          if (u < 0 || v < 0 || w < 0 || u >= g.n[0] || v >= g.n[1] || w >= g.n[2]) continue;
          need |= 1u << owner_of(g, u, v, w);
        }
    need &= ~(1u << g.rank);
    const uint2* src = reinterpret_cast<const uint2*>(rec);          // 40-byte records, 8-byte aligned
    while (need) {
      const int p = __ffs(need) - 1;
      need &= need - 1;
      const int pos = atomicAdd(&sent[p], 1);
      if (pos < cap) {
        uint2* dst = reinterpret_cast<uint2*>(ex_inbox(peers.base[p], buf, g.world, cap, g.rank) + (int64_t)pos * kRecWords);
#pragma unroll
        for (int j = 0; j < 5; ++j) dst[j] = src[j];
      } else {
        atomicOr(&m.ctr[2], kErrCapacity);
      }
    }
  }
  // the last block publishes counts and the frame sequence number to every peer
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(done, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < g.world && threadIdx.x != g.rank) {
      const int p = threadIdx.x;
      const int c = atomicAdd(&sent[p], 0);
      ex_count(peers.base[p], buf)[g.rank] = c < cap ? c : (int)cap;
      __threadfence_system();
      st_release_sys(ex_ready(peers.base[p]) + g.rank, seq);
    }
    __syncthreads();
    if (threadIdx.x < g.world) sent[threadIdx.x] = 0;     // ready for the next frame (stream-ordered)
    if (threadIdx.x == 0) *done = 0;
  }
}

// receiver: upsert the inbox of frame `seq` (8 lanes per record), then acknowledge it to every sender
__global__ void __launch_bounds__(256) insert_inbox_kernel(MapDev m, PeerPtrs peers, int buf, uint32_t seq, int64_t cap,
                                                           int32_t* __restrict__ done) {
  const GeomDev& g = m.g;
  if (m.ctr[2] & kErrExchange) return;     // a peer's records never arrived (latched by wait_flags_kernel): no upsert, no ack
  int32_t* base = peers.base[g.rank];
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
  const int lane8 = threadIdx.x & 7;
  for (int r = 0; r < g.world; ++r) {
    if (r == g.rank) continue;
    const int n = min(ex_count(base, buf)[r], (int)cap);
    const int32_t* inbox = ex_inbox(base, buf, g.world, cap, r);
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < n;
         i += ((int64_t)gridDim.x * blockDim.x) >> 3) {
      const int32_t* rec = inbox + i * kRecWords;
      const int32_t flat = rec[0];
      int32_t slot = -1;
      if (lane8 == 0 && flat >= 0 && (int64_t)flat < g.n_vox) {
        slot = m.table[flat];
        if (slot < 0) {                      // a voxel reaches a rank from exactly one owner: keys are unique per frame
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
  __threadfence();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(done, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < g.world && threadIdx.x != g.rank) st_release_sys(ex_ack(peers.base[threadIdx.x]) + g.rank, seq);
    if (threadIdx.x == 0) *done = 0;
  }
}

}  // namespace bnv

using namespace bnv;

struct bnv_exchange {
  bnv_map_t* map;
  int rank, world;
  int64_t cap;
  int32_t* block;              // this rank's exchange block (cudaMalloc, IPC-exported)
  int32_t* scratch;            // [kMaxPeers] sent counters | push done | insert done
  PeerPtrs peers;
  bool connected;
  uint32_t seq;                // frames begun so far
  int32_t* halo[2];            // this rank's boundary records, double-buffered by frame parity (attached to the map)
  int64_t halo_cap;
  cudaStream_t side;
  cudaEvent_t fused, pushed[2], upserted;
  bool pushed_valid[2];
  bool any_upsert;
};

extern "C" {

int bnv_exchange_create(bnv_exchange_t** out, bnv_map_t* map, int64_t capacity_records) {
  if (!out || !map || capacity_records <= 0 || capacity_records > 0x7fffffff) { set_error("bnv_exchange_create: bad argument"); return BNV_E_ARG; }
  const int world = map->d.g.world, rank = map->d.g.rank;
  if (world < 2 || world > kMaxPeers) { set_error("bnv_exchange_create: world must be 2..%d (call bnv_map_set_shard first)", kMaxPeers); return BNV_E_ARG; }
  bnv_exchange* ex = new (std::nothrow) bnv_exchange();
  if (!ex) return BNV_E_ALLOC;
  memset(ex, 0, sizeof(*ex));
  ex->map = map; ex->rank = rank; ex->world = world; ex->cap = capacity_records;
  BNV_CUDA(cudaSetDevice(map->device));
  const size_t bytes = (size_t)ex_block_words(world, capacity_records) * 4;
  cudaError_t e = cudaMalloc((void**)&ex->block, bytes);
  if (e == cudaSuccess) e = cudaMemset(ex->block, 0, bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ex->scratch, (kMaxPeers + 2) * 4);
  if (e == cudaSuccess) e = cudaMemset(ex->scratch, 0, (kMaxPeers + 2) * 4);
  ex->halo_cap = capacity_records;
  for (int i = 0; i < 2; ++i) {
    if (e == cudaSuccess) e = cudaMalloc((void**)&ex->halo[i], (10 + (size_t)capacity_records * kRecWords) * 4);
    if (e == cudaSuccess) e = cudaMemset(ex->halo[i], 0, 40);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ex->pushed[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ex->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ex->fused, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ex->upserted, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    set_error("bnv_exchange_create: %s", cudaGetErrorString(e));
    ex->world = 0;                          // nothing was opened on peers
    bnv_exchange_destroy(ex);
    return BNV_E_ALLOC;
  }
  BNV_CUDA(cudaDeviceSynchronize());
  ex->peers.base[rank] = ex->block;
  *out = ex;
  return BNV_OK;
}

/* 64-byte cudaIpcMemHandle_t of this rank's exchange block: all-gather these between the ranks (any transport) */
int bnv_exchange_handle(bnv_exchange_t* ex, void* handle64_out) {
  if (!ex || !handle64_out) { set_error("bnv_exchange_handle: bad argument"); return BNV_E_ARG; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  BNV_CUDA(cudaIpcGetMemHandle(&h, ex->block));
  memcpy(handle64_out, &h, 64);
  return BNV_OK;
}

int bnv_exchange_connect(bnv_exchange_t* ex, const void* handles /* [world][64], rank order */) {
  if (!ex || !handles) { set_error("bnv_exchange_connect: bad argument"); return BNV_E_ARG; }
  BNV_CUDA(cudaSetDevice(ex->map->device));
  for (int r = 0; r < ex->world; ++r) {
    if (r == ex->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * r, 64);
    void* p = nullptr;
    BNV_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ex->peers.base[r] = (int32_t*)p;
  }
  ex->connected = true;
  return BNV_OK;
}

/* One exchange epoch.  On `stream`: pack the records of the shell voxels integrated since the last exchange
 * (bnv_map_halo_pack) into this epoch's record buffer.  Everything else happens on the exchange's side stream,
 * overlapped with the frames that follow -- route the records to the peers that need them, wait for every peer's
 * records of the same epoch, upsert them, acknowledge.  Every rank must call it at the same points of the stream. */
int bnv_map_halo_pack(bnv_map_t* m, void* buf, int64_t capacity_records, void* stream);
int bnv_exchange_push(bnv_exchange_t* ex, void* stream) {
  if (!ex || !ex->connected) { set_error("bnv_exchange_push: exchange is not connected"); return BNV_E_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  const uint32_t seq = ++ex->seq;
  const int buf = (int)(seq & 1u);
  const long long timeout = 4000000000ll;                       // ~2 s of SM clock
  uint32_t* flags_ready = reinterpret_cast<uint32_t*>(ex->block);
  uint32_t* flags_ack = flags_ready + kMaxPeers;
  if (ex->pushed_valid[buf]) BNV_CUDA(cudaStreamWaitEvent(s, ex->pushed[buf], 0));     // epoch seq - 2 has been routed
  int rc = bnv_map_halo_pack(ex->map, ex->halo[buf], ex->halo_cap, stream);
  if (rc) return rc;
  BNV_CUDA(cudaEventRecord(ex->fused, s));
  BNV_CUDA(cudaStreamWaitEvent(ex->side, ex->fused, 0));
  if (seq > 2) {                                                // the peers' inbox parity `buf` was last used by epoch seq - 2
    wait_flags_kernel<<<1, 32, 0, ex->side>>>(flags_ack, seq - 2, ex->world, ex->rank, timeout, &ex->map->d.ctr[2]);
    BNV_LAUNCH_CHECK("wait_flags_kernel");
  }
  MapDev d = ex->map->d;
  halo_push_kernel<<<64, 256, 0, ex->side>>>(d, ex->halo[buf], ex->peers, buf, seq, ex->cap, ex->scratch, ex->scratch + kMaxPeers);
  BNV_LAUNCH_CHECK("halo_push_kernel");
  BNV_CUDA(cudaEventRecord(ex->pushed[buf], ex->side));
  ex->pushed_valid[buf] = true;
  wait_flags_kernel<<<1, 32, 0, ex->side>>>(flags_ready, seq, ex->world, ex->rank, timeout, &ex->map->d.ctr[2]);
  BNV_LAUNCH_CHECK("wait_flags_kernel");
  insert_inbox_kernel<<<64, 256, 0, ex->side>>>(d, ex->peers, buf, seq, ex->cap, ex->scratch + kMaxPeers + 1);
  BNV_LAUNCH_CHECK("insert_inbox_kernel");
  BNV_CUDA(cudaEventRecord(ex->upserted, ex->side));
  ex->any_upsert = true;
  return BNV_OK;
}

/* `stream` waits for every upsert issued so far (call before reading the map: export, decode, size) */
int bnv_exchange_join(bnv_exchange_t* ex, void* stream) {
  if (!ex) { set_error("bnv_exchange_join: null"); return BNV_E_ARG; }
  if (ex->any_upsert) BNV_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ex->upserted, 0));
  return BNV_OK;
}

int bnv_exchange_destroy(bnv_exchange_t* ex) {
  if (!ex) return BNV_OK;
  cudaSetDevice(ex->map->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < ex->world; ++r)
    if (r != ex->rank && ex->peers.base[r]) cudaIpcCloseMemHandle(ex->peers.base[r]);
  if (ex->block) cudaFree(ex->block);
  if (ex->scratch) cudaFree(ex->scratch);
  for (int i = 0; i < 2; ++i) {
    if (ex->halo[i]) cudaFree(ex->halo[i]);
    if (ex->pushed[i]) cudaEventDestroy(ex->pushed[i]);
  }
  if (ex->side) cudaStreamDestroy(ex->side);
  if (ex->fused) cudaEventDestroy(ex->fused);
  if (ex->upserted) cudaEventDestroy(ex->upserted);
  delete ex;
  return BNV_OK;
}

}  // extern "C"
