// Finalize step of the fused per-frame path (finalize_fused_kernel, bnv_encode.cu): per touched voxel -> mean, count filter, running average into the
// persistent map (local_point_fusion.py:125,143-151,647-673); then the frame statistics.
#pragma once
#include "bnv_common.cuh"
#include "bnv_frame.cuh"

namespace bnv {

// mean of a scratch row (scatter_mean, local_point_fusion.py:125)
__device__ __forceinline__ float scratch_mean(const MapDev& m, int32_t row, int j, int32_t cnt, bool f32acc) {
  if (f32acc)   // tensor-core mode: fp32 partial sums (add_row_f32)
    return (float)((double)reinterpret_cast<const float*>(m.fsum)[(size_t)row * kFeat + j] / (double)cnt);
  const long long s = m.fsum[(size_t)row * kFeat + j];
  return (float)(((double)s / kFixScale) / (double)cnt);
}
__device__ __forceinline__ void scratch_clear(const MapDev& m, int32_t row, int j, bool f32acc) {
  if (f32acc) reinterpret_cast<float*>(m.fsum)[(size_t)row * kFeat + j] = 0.f;
  else m.fsum[(size_t)row * kFeat + j] = 0;
}

// _update (local_point_fusion.py:647-651), separately rounded like the reference's torch kernels
__device__ __forceinline__ float fuse_feat(float f_old, float w_old, float f_new, float w_new, float w) {
  return __fdiv_rn(__fadd_rn(__fmul_rn(f_old, w_old), __fmul_rn(f_new, w_new)), w);
}

// One thread per touched voxel (dense scratch rows first, first + stride, ...): the work is a chain of dependent
// scattered reads (key -> table entries -> map slot -> old features), i.e. latency x concurrency bound, so every
// thread keeps its own voxel's chain in flight; the 32-byte scratch and feature rows move as two 16-byte accesses.
// Returns the number of voxels this thread integrated.
__device__ __forceinline__ int finalize_rows(const MapDev& m, int min_pts, bool f32acc, int n_touched, int64_t first,
                                             int64_t stride) {
  int integrated = 0;
  for (int64_t t = first; t < n_touched; t += stride) {
    const int32_t key = m.fkeys[t];
    unsigned long long* ent = ft_entry(m, key, 0);
    const int32_t cnt = ft_count(*ent);
    int32_t slot = m.table[key];                                  // independent of the count: both reads in flight
    *ent = 0ull;
    float mean[kFeat];                                            // scatter_mean, local_point_fusion.py:125
    if (f32acc) {                                                 // tensor-core mode: fp32 partial sums (add_row_f32)
      float4* s4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(m.fsum) + (size_t)t * kFeat);
      const float4 a = s4[0], b = s4[1];
      s4[0] = s4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float s[kFeat] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      // float32 sums / exact integer count: one correctly rounded float32 division (the float64 form of the exact-
      // parity mode costs ~45 instructions per feature and was 90 % of this kernel's instruction count)
      const float fc = (float)cnt;
#pragma unroll
      for (int j = 0; j < kFeat; ++j) mean[j] = __fdiv_rn(s[j], fc);
    } else {
      longlong2* s2 = reinterpret_cast<longlong2*>(m.fsum + (size_t)t * kFeat);
#pragma unroll
      for (int j = 0; j < kFeat / 2; ++j) {
        const longlong2 v = s2[j];
        s2[j] = make_longlong2(0, 0);
        mean[2 * j] = (float)(((double)v.x / kFixScale) / (double)cnt);
        mean[2 * j + 1] = (float)(((double)v.y / kFixScale) / (double)cnt);
      }
    }
    if (cnt < min_pts) continue;                                  // local_point_fusion.py:143-147
    bool fresh = false;
    if (slot < 0) {
      slot = atomicAdd(&m.ctr[0], 1);
      if (slot >= m.cap) {
        atomicOr(&m.ctr[2], kErrCapacity);
        continue;
      }
      fresh = true;
      m.table[key] = slot;
      m.keys[slot] = key;
      m.hits[slot] = 0.f;
    }
    float4* f4 = reinterpret_cast<float4*>(m.feats + (size_t)slot * kFeat);
    float w_old = 0.f;
    float4 oa = make_float4(0.f, 0.f, 0.f, 0.f), ob = oa;
    if (!fresh) {
      w_old = m.weights[slot];
      oa = f4[0];
      ob = f4[1];
    }
    const float w_new = fminf(__fmul_rn((float)cnt, 0.03125f), 1.0f);   // clip(count/32, max=1)
    const float w = __fadd_rn(w_old, w_new);
    const float f_old[kFeat] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
    float f_new[kFeat];
#pragma unroll
    for (int j = 0; j < kFeat; ++j) f_new[j] = fuse_feat(f_old[j], w_old, mean[j], w_new, w);
    f4[0] = make_float4(f_new[0], f_new[1], f_new[2], f_new[3]);
    f4[1] = make_float4(f_new[4], f_new[5], f_new[6], f_new[7]);
    m.weights[slot] = w;
    ++integrated;
    if (m.dirty_list) {                                           // tile shard: another rank may need it as a corner
      const int kx = key / m.g.nyz, kr = key - kx * m.g.nyz, ky = kr / m.g.n[2], kz = kr - ky * m.g.n[2];
      if (on_brick_shell(m.g, kx, ky, kz) && atomicExch(&m.dirty_flag[slot], 1) == 0) {
        const int pos = atomicAdd(&m.ctr[5], 1);                  // once per voxel and exchange epoch
        if (pos < m.dirty_cap) m.dirty_list[pos] = slot;
        else atomicOr(&m.ctr[2], kErrCapacity);
      }
    }
  }
  return integrated;
}

// Frame batch (bnv_fuse_frames): scratch rows are (frame, voxel) pairs; ONE group of 8 lanes per voxel (a lane per
// feature) applies the voxel's frames in frame order, which is exactly the sequence of _update calls the per-frame path
// makes, with the voxel's weight and features held in registers in between (one read-modify-write of the map per voxel
// and batch instead of one per frame).  The voxel's rows are all requested before the first one is used -- the frame
// loop is then pure arithmetic -- and the stored voxel is read alongside them: three dependent memory round trips per
// voxel whatever the batch size.  (One thread per voxel walking its rows one after the other ran 213 us on 7 frames,
// 99 % of it waiting for the row of the next frame: profiles/r2e.)
// The group is chosen by the first of the voxel's rows to swap the batch sequence number into the cell's lock word
// (the last of its S table words); the lock is never released -- the next batch brings a new number -- so a row that
// arrives late can never see a half-cleared cell.  S = table words per cell (frames <= S - 1).
//
// finalize_batch_voxel: all frames of one voxel; `sub` = feature of this lane, `gmask` = the group's lanes (converged).
// Returns the number of (frame, voxel) pairs integrated (on the group's first lane, 0 on the others).
template <int S, bool F32>
__device__ __forceinline__ int finalize_batch_voxel(const MapDev& m, int min_pts, int32_t key, int sub, unsigned gmask) {
  unsigned long long* ent = ft_entry(m, key, 0);
  int32_t slot = m.table[key];
  unsigned long long e[S];
#pragma unroll
  for (int j = 0; j < S / 2; ++j) {                               // the cell's words: S / 2 16-byte loads, L1 bypassed
    const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(ent) + j);
    e[2 * j] = v.x;
    e[2 * j + 1] = v.y;
  }
  // this lane's feature of every frame's scratch row: independent loads, all in flight
  float rawf[S - 1];
  long long rawi[F32 ? 1 : S - 1];
#pragma unroll
  for (int fr = 0; fr < S - 1; ++fr) {
    rawf[fr] = 0.f;
    if (!F32) rawi[F32 ? 0 : fr] = 0;
    if (ft_count(e[fr]) != 0) {
      const size_t at = (size_t)ft_row(e[fr]) * kFeat + sub;
      if (F32) rawf[fr] = __ldcg(reinterpret_cast<const float*>(m.fsum) + at);
      else rawi[F32 ? 0 : fr] = __ldcg(m.fsum + at);
    }
  }
  // the stored voxel (if any): independent of the rows
  float w = 0.f, f = 0.f;
  if (slot >= 0) {
    w = m.weights[slot];
    f = m.feats[(size_t)slot * kFeat + sub];
  }
  // re-arm the scratch rows and the cell's frame words
#pragma unroll
  for (int fr = 0; fr < S - 1; ++fr)
    if (ft_count(e[fr]) != 0) {
      const size_t at = (size_t)ft_row(e[fr]) * kFeat + sub;
      if (F32) reinterpret_cast<float*>(m.fsum)[at] = 0.f;
      else m.fsum[at] = 0;
      if (sub == 0) ent[fr] = 0ull;
    }
  int integrated = 0;
  bool dead = false;
#pragma unroll
  for (int fr = 0; fr < S - 1; ++fr) {
    const int32_t cnt = ft_count(e[fr]);
    if (cnt < min_pts || cnt == 0 || dead) continue;              // local_point_fusion.py:143-147
    // scatter_mean, local_point_fusion.py:125 (same roundings as finalize_rows)
    const float mean = F32 ? __fdiv_rn(rawf[fr], (float)cnt) : (float)(((double)rawi[F32 ? 0 : fr] / kFixScale) / (double)cnt);
    if (slot < 0) {                                               // first frame that integrates a new voxel
      int32_t s2 = -1;
      if (sub == 0) {
        s2 = atomicAdd(&m.ctr[0], 1);
        if (s2 >= m.cap) {
          atomicOr(&m.ctr[2], kErrCapacity);
          s2 = -1;
        } else {
          m.table[key] = s2;
          m.keys[s2] = key;
          m.hits[s2] = 0.f;
        }
      }
      s2 = __shfl_sync(gmask, s2, __ffs(gmask) - 1);
      if (s2 < 0) {
        dead = true;
        continue;
      }
      slot = s2;
    }
    const float w_new = fminf(__fmul_rn((float)cnt, 0.03125f), 1.0f);   // clip(count/32, max=1)
    const float w_sum = __fadd_rn(w, w_new);
    f = fuse_feat(f, w, mean, w_new, w_sum);
    w = w_sum;
    ++integrated;
  }
  if (integrated == 0) return 0;
  m.feats[(size_t)slot * kFeat + sub] = f;
  if (sub != 0) return 0;
  m.weights[slot] = w;
  if (m.dirty_list) {                                             // tile shard: another rank may need it as a corner
    const int kx = key / m.g.nyz, kr = key - kx * m.g.nyz, ky = kr / m.g.n[2], kz = kr - ky * m.g.n[2];
    if (on_brick_shell(m.g, kx, ky, kz) && atomicExch(&m.dirty_flag[slot], 1) == 0) {
      const int pos = atomicAdd(&m.ctr[5], 1);
      if (pos < m.dirty_cap) m.dirty_list[pos] = slot;
      else atomicOr(&m.ctr[2], kErrCapacity);
    }
  }
  return integrated;
}

// A block takes chunks of kBatchRows scratch rows (handed out by an atomic counter, ctr[6]: the chunks' costs differ
// with their share of winners): (1) every thread swaps the sequence number into the lock words of kBatchRows / blockDim
// rows (independent atomics, all in flight) and the winners' voxels go to a shared-memory list; (2) the list is spread
// densely over the block's 8-lane groups.  Most rows lose (a voxel is touched by most frames of the batch), so the
// per-voxel work runs in full warps.
constexpr int kBatchRows = 512;
template <int S, bool F32>
__device__ __forceinline__ int finalize_batch_rows(const MapDev& m, int min_pts, unsigned int seq, int n_touched) {
  __shared__ int32_t s_lead[kBatchRows];
  __shared__ int s_n, s_chunk;
  constexpr int R = kBatchRows / 256;
  const int sub = threadIdx.x & 7;
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);
  const int n_chunks = (n_touched + kBatchRows - 1) / kBatchRows;
  int integrated = 0;
  for (;;) {
    if (threadIdx.x == 0) {
      s_n = 0;
      s_chunk = atomicAdd(&m.ctr[6], 1);
    }
    __syncthreads();
    const int chunk = s_chunk;
    if (chunk >= n_chunks) break;                                 // block-uniform
    const int64_t base = (int64_t)chunk * kBatchRows;
    int32_t key[R];
    unsigned int old[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t t = base + r * 256 + threadIdx.x;
      key[r] = t < n_touched ? m.fkeys[t] : -1;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      old[r] = seq;
      if (key[r] >= 0) old[r] = atomicExch(reinterpret_cast<unsigned int*>(ft_entry(m, key[r], S - 1)), seq);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (old[r] != seq) {
        s_lead[atomicAdd(&s_n, 1)] = key[r];
        asm volatile("prefetch.global.L2 [%0];" ::"l"(m.table + key[r]));      // the winner's group reads it next
      }
    __syncthreads();
    const int n_lead = s_n;
    for (int i = threadIdx.x >> 3; i < n_lead; i += 32) integrated += finalize_batch_voxel<S, F32>(m, min_pts, s_lead[i], sub, gmask);
    __syncthreads();
  }
  return integrated;
}

// Block-level end of finalize: one statistics atomic per block, the last block publishes the frame statistics and
// re-arms the per-frame counters.  (Running this step in the tail of the persistent encoder kernel, behind a grid-wide
// barrier, was measured in round 2: 8.8k instead of 9.9k frames/s -- the barrier waits for the slowest chain and the
// tail then runs with one CTA per SM.)
__device__ __forceinline__ void finalize_publish(const MapDev& m, int integrated, int n_touched, long long* __restrict__ stats,
                                                 long long* __restrict__ user_stats, float* __restrict__ user_navg) {
  // one statistics atomic per block (same-address atomics serialise in L2)
  __shared__ int s_integrated;
  if (threadIdx.x == 0) s_integrated = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) integrated += __shfl_xor_sync(0xffffffffu, integrated, o);
  if ((threadIdx.x & 31) == 0 && integrated) atomicAdd(&s_integrated, integrated);
  __syncthreads();
  if (threadIdx.x == 0 && s_integrated) atomicAdd(reinterpret_cast<unsigned long long*>(stats + 3), (unsigned long long)s_integrated);
  // last block publishes the frame statistics and re-arms the counters.  Only thread 0 needs the fence: the one thing
  // the last block reads from the others is the statistics atomic that thread 0 itself issued above (a fence executed
  // by all 256 threads of every block was 30 % of this kernel's stall samples, profiles/r2b)
  // (only thread 0 takes part: the other warps retire right away instead of waiting at a block barrier for thread 0's
  // fence + atomic round trip -- 25 % of the kernel's stall samples in profiles/r2d)
  if (threadIdx.x != 0) return;
  __threadfence();
  const bool last = atomicAdd(&m.ctr[3], 1) == (int)gridDim.x - 1;
  if (last) {
    __threadfence();
    const long long rows = stats[1];
    if (user_stats) {
      user_stats[0] = stats[0];
      user_stats[1] = rows;
      user_stats[2] = n_touched;
      user_stats[3] = *reinterpret_cast<volatile long long*>(stats + 3);
    }
    if (user_navg) *user_navg = n_touched > 0 ? (float)((double)rows / (double)n_touched) : 0.f;
    stats[0] = stats[1] = stats[3] = stats[4] = 0;
    m.ctr[1] = 0;
    m.ctr[3] = 0;
    m.ctr[4] = 0;
    m.ctr[6] = 0;                                                 // chunk counter of finalize_batch_rows
  }
}

}  // namespace bnv
