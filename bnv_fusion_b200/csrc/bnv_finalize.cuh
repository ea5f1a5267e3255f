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

// Tensor-core mode: the features already carry the fp16 operand rounding of the MLP (1e-3), so the two divisions per
// feature and frame -- mean = sum / count and the weighted average / (w_old + w_new) -- are done as one correctly
// rounded reciprocal per voxel and frame and a multiplication per feature (<= 1 ulp of fp32 from the true quotient);
// IEEE divisions were three quarters of the batch finalize's instructions.  The exact-parity mode keeps the reference's
// true divisions.  Both finalize flavours use the same arithmetic, so a batch equals its frames one by one.
__device__ __forceinline__ float fuse_feat_rcp(float f_old, float w_old, float f_new, float w_new, float rcp_w) {
  return __fmul_rn(__fadd_rn(__fmul_rn(f_old, w_old), __fmul_rn(f_new, w_new)), rcp_w);
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
      // float32 sums x 1 / count (the float64 form of the exact-parity mode costs ~45 instructions per feature and
      // was 90 % of this kernel's instruction count)
      const float rc = __frcp_rn((float)cnt);
#pragma unroll
      for (int j = 0; j < kFeat; ++j) mean[j] = __fmul_rn(s[j], rc);
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
    if (f32acc) {
      const float rw = __frcp_rn(w);
#pragma unroll
      for (int j = 0; j < kFeat; ++j) f_new[j] = fuse_feat_rcp(f_old[j], w_old, mean[j], w_new, rw);
    } else {
#pragma unroll
      for (int j = 0; j < kFeat; ++j) f_new[j] = fuse_feat(f_old[j], w_old, mean[j], w_new, w);
    }
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

// Frame batch (bnv_fuse_frames): scratch rows are (frame, voxel) pairs; ONE thread per voxel applies the voxel's frames
// in frame order, which is exactly the sequence of _update calls the per-frame path makes, with the voxel's weight and
// features held in registers in between (one read-modify-write of the map per voxel and batch instead of one per
// frame).  The voxel's rows are ALL requested before the first one is used (unconditional 16-byte loads, an untouched
// frame reads row 0 and ignores it) and the stored voxel is read alongside them: three dependent memory round trips
// per voxel whatever the batch size; the frame loop is then pure arithmetic.  (Walking the rows one after the other
// ran 213 us on 7 frames, 99 % of it waiting for the next frame's row; 8 lanes per voxel kept an eighth as many voxels
// in flight and ran 86 us -- profiles/r2e.)
// The thread is chosen by the first of the voxel's rows to swap the batch sequence number into the cell's lock word
// (the last of its kCellWords table words); the lock is never released -- the next batch brings a new number -- so a
// row that arrives late can never see a half-cleared cell.
constexpr int kCellWords = 8;                                     // 2^fshift: kMaxBatch frame words + the lock word
static_assert(kMaxBatch == kCellWords - 1, "one table word per frame of a batch plus the lock");

// all frames of one voxel; returns the number of (frame, voxel) pairs integrated
template <bool F32>
__device__ __forceinline__ int finalize_batch_voxel(const MapDev& m, int min_pts, int32_t key, int32_t slot) {
  constexpr int S = kCellWords;
  unsigned long long* ent = ft_entry(m, key, 0);
  unsigned long long e[S];
#pragma unroll
  for (int j = 0; j < S / 2; ++j) {                               // the cell's words: four 16-byte loads, L1 bypassed
    const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(ent) + j);
    e[2 * j] = v.x;
    e[2 * j + 1] = v.y;
  }
  float4 ra[F32 ? S - 1 : 1], rb[F32 ? S - 1 : 1];                 // tensor-core mode: the fp32 sums of every frame
  if (F32) {
#pragma unroll
    for (int fr = 0; fr < S - 1; ++fr) {
      const int32_t row = ft_count(e[fr]) != 0 ? ft_row(e[fr]) : 0;
      const float4* s4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(m.fsum) + (size_t)row * kFeat);
      ra[F32 ? fr : 0] = __ldcg(s4);
      rb[F32 ? fr : 0] = __ldcg(s4 + 1);
    }
  }
  // the stored voxel (if any): independent of the rows
  float w = 0.f;
  float f[kFeat];
#pragma unroll
  for (int j = 0; j < kFeat; ++j) f[j] = 0.f;
  if (slot >= 0) {
    const float4* f4 = reinterpret_cast<const float4*>(m.feats + (size_t)slot * kFeat);
    const float4 oa = f4[0], ob = f4[1];
    w = m.weights[slot];
    f[0] = oa.x; f[1] = oa.y; f[2] = oa.z; f[3] = oa.w;
    f[4] = ob.x; f[5] = ob.y; f[6] = ob.z; f[7] = ob.w;
  }
  int integrated = 0;
  bool dead = false;
#pragma unroll
  for (int fr = 0; fr < S - 1; ++fr) {
    const int32_t cnt = ft_count(e[fr]);
    if (cnt == 0) continue;                                       // frame fr did not touch the voxel
    const int32_t row = ft_row(e[fr]);
    ent[fr] = 0ull;                                               // re-arm the cell's frame word and the scratch row
    float mean[kFeat];                                            // scatter_mean, local_point_fusion.py:125
    if (F32) {
      float4* s4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(m.fsum) + (size_t)row * kFeat);
      s4[0] = s4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 a = ra[F32 ? fr : 0], b = rb[F32 ? fr : 0];
      const float sv[kFeat] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      const float rc = __frcp_rn((float)cnt);
#pragma unroll
      for (int j = 0; j < kFeat; ++j) mean[j] = __fmul_rn(sv[j], rc);
    } else {
      longlong2* s2 = reinterpret_cast<longlong2*>(m.fsum + (size_t)row * kFeat);
#pragma unroll
      for (int j = 0; j < kFeat / 2; ++j) {
        const longlong2 v = s2[j];
        s2[j] = make_longlong2(0, 0);
        mean[2 * j] = (float)(((double)v.x / kFixScale) / (double)cnt);
        mean[2 * j + 1] = (float)(((double)v.y / kFixScale) / (double)cnt);
      }
    }
    if (cnt < min_pts || dead) continue;                          // local_point_fusion.py:143-147
    if (slot < 0) {                                               // first frame that integrates a new voxel
      slot = atomicAdd(&m.ctr[0], 1);
      if (slot >= m.cap) {
        atomicOr(&m.ctr[2], kErrCapacity);
        dead = true;
        continue;
      }
      m.table[key] = slot;
      m.keys[slot] = key;
      m.hits[slot] = 0.f;
    }
    const float w_new = fminf(__fmul_rn((float)cnt, 0.03125f), 1.0f);   // clip(count/32, max=1)
    const float w_sum = __fadd_rn(w, w_new);
    if (F32) {
      const float rw = __frcp_rn(w_sum);
#pragma unroll
      for (int j = 0; j < kFeat; ++j) f[j] = fuse_feat_rcp(f[j], w, mean[j], w_new, rw);
    } else {
#pragma unroll
      for (int j = 0; j < kFeat; ++j) f[j] = fuse_feat(f[j], w, mean[j], w_new, w_sum);
    }
    w = w_sum;
    ++integrated;
  }
  if (integrated == 0) return 0;
  float4* f4 = reinterpret_cast<float4*>(m.feats + (size_t)slot * kFeat);
  f4[0] = make_float4(f[0], f[1], f[2], f[3]);
  f4[1] = make_float4(f[4], f[5], f[6], f[7]);
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

// The scratch rows are split evenly over the blocks (up to kBatchRowsMax per block and sweep).  (1) Every thread swaps
// the sequence number into the lock words of its rows -- independent atomics, four in flight per thread -- and the
// winners look their voxel up in the persistent table right away (the stored voxel's lines are on their way into L2
// while the other rows are still being resolved); the winners' (voxel, slot) pairs go to a shared-memory list.
// (2) The list is spread densely over the block's threads.  Most rows lose (a voxel is touched by most frames of the
// batch), so the heavy per-voxel work runs in full warps, and its memory round trips start from L2.
constexpr int kBatchRowsMax = 4096;
template <bool F32>
__device__ __forceinline__ int finalize_batch_rows(const MapDev& m, int min_pts, unsigned int seq, int n_touched) {
  __shared__ int32_t s_lead[kBatchRowsMax], s_slot[kBatchRowsMax];
  __shared__ int s_n;
  int per_block = (int)(((int64_t)n_touched + gridDim.x - 1) / gridDim.x);
  per_block = min(kBatchRowsMax, (per_block + 255) & ~255);
  int integrated = 0;
  for (int64_t base = (int64_t)blockIdx.x * per_block; base < n_touched; base += (int64_t)gridDim.x * per_block) {
    __syncthreads();                                              // the previous sweep's list has been consumed
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int r0 = 0; r0 < per_block; r0 += 4 * 256) {
      int32_t key[4], slot[4];
      unsigned int old[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int64_t t = base + r0 + r * 256 + threadIdx.x;
        key[r] = (r0 + r * 256 < per_block && t < n_touched) ? m.fkeys[t] : -1;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        old[r] = seq;
        if (key[r] >= 0) old[r] = atomicExch(reinterpret_cast<unsigned int*>(ft_entry(m, key[r], kCellWords - 1)), seq);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) slot[r] = old[r] != seq ? m.table[key[r]] : -1;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (old[r] != seq) {
          const int at = atomicAdd(&s_n, 1);
          s_lead[at] = key[r];
          s_slot[at] = slot[r];
          if (slot[r] >= 0) {                                     // the stored voxel: pull its lines into L2
            asm volatile("prefetch.global.L2 [%0];" ::"l"(m.feats + (size_t)slot[r] * kFeat));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(m.weights + slot[r]));
          }
        }
    }
    __syncthreads();
    const int n_lead = s_n;
    for (int i = threadIdx.x; i < n_lead; i += 256) integrated += finalize_batch_voxel<F32>(m, min_pts, s_lead[i], s_slot[i]);
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
  }
}

}  // namespace bnv
