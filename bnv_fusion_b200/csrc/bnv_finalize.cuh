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
    const int32_t cnt = ft_count(m.ftable[key]);
    int32_t slot = m.table[key];                                  // independent of the count: both reads in flight
    m.ftable[key] = 0ull;
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
