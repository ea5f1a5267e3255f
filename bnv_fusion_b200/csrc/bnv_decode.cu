// SDF decode at query points: 8-corner gather from the voxel map + decoder MLP per corner +
// trilinear blend + validity mask + TSDF-prior lookup, one kernel.
//
// Reference: SparseVolume.decode_pts (src/models/sparse_volume.py:768-833), fusion/utils.get_neighbors
// (src/models/fusion/utils.py:98-167), _query_tensor (sparse_volume.py:625-659), positional_encoding
// (src/models/fusion/modules.py:81-123), tcnnNeRFModel.geo_forward (modules.py:249-253),
// F.grid_sample(nearest) of the prior (sparse_volume.py:819-832), meshlize sampling (:717-731).
// This file holds the fp32 CUDA-core variant (BNV_MLP_FP32); the tcgen05 variant is in bnv_tc.cu.
#include <stdlib.h>

#include "bnv_common.cuh"
#include "bnv_decode_common.cuh"
#include "bnv_mlp_simt.cuh"

namespace bnv {

using DecMlp = SimtMlp<17, 1>;
constexpr int kDecThreads = 256;
constexpr size_t kDecSmem = (size_t)(DecMlp::kFloats + 64 * kDecThreads) * sizeof(float);

// One query: rules D1-D7 of SURVEY.md §8a.
__device__ __forceinline__ float decode_query_simt(const MapDev& m, const DecArgs& a, const float* sW, float* sH,
                                                   float cx, float cy, float cz, bool* mask_out) {
  const float c[3] = {cx, cy, cz};
  float fl[3], ce[3];
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    fl[ax] = floorf(c[ax]);
    ce[ax] = ceilf(c[ax]);
  }
  const float wsum = corner_weight_sum(c, fl, ce);
  float minw = 3.0e38f, sdf = 0.f, dsum = 0.f;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    float nb[3];
    corner_of(k, fl, ce, nb);
    float x[17];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      const float l = __fsub_rn(c[ax], nb[ax]);                  // D1
      x[ax] = l;
      x[3 + ax] = sinf(l);                                       // positional encoding, 1 frequency
      x[6 + ax] = cosf(l);
    }
    float wt;
    gather_corner(m, a, nb, x + 9, wt);                          // D3 (misses -> zeros)
    minw = fminf(minw, wt);
    float y[1];
    DecMlp::run(sW, sH, kDecThreads, x, y);                      // D7: evaluated for every corner
    const float wn = __fdiv_rn(corner_weight(c, nb), wsum);      // D2
    sdf = __fadd_rn(sdf, __fmul_rn(__fmul_rn(y[0], m.g.vs), wn));   // D4, D5
    if (a.tsdf) dsum = __fadd_rn(dsum, __fmul_rn(tsdf_nearest(a, m.g, nb), wn));
  }
  return finish_blend(sdf, dsum, minw, a, m.g.vs, mask_out);
}

__global__ void __launch_bounds__(kDecThreads) decode_simt_kernel(MapDev m, DecArgs a, const float* __restrict__ gW) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* sH = smem + DecMlp::kFloats + threadIdx.x;
  load_weights(sW, gW, DecMlp::kFloats);
  const int64_t q = (int64_t)blockIdx.x * kDecThreads + threadIdx.x;
  if (q >= a.n_queries) return;
  float c[3];
  query_coords(m, a, q, c);
  bool mask;
  const float sdf = decode_query_simt(m, a, sW, sH, c[0], c[1], c[2], &mask);
  a.out_sdf[q] = sdf;
  if (a.out_mask) a.out_mask[q] = mask ? 1 : 0;
}

// ---- backward: d(sum(grad_out * sdf)) / d(features)  (NeuralMap.optimize, src/run_e2e.py:111-156) -------------
// What torch autograd computes through SparseVolume.decode_pts when volume.features is the only leaf: per
// query and found corner, back-propagate dL/dy_k = g * w^_k * voxel_size (zero when the query's validity
// mask is false; the prior and the fallback value do not depend on the features) through the decoder MLP
// and accumulate the feature columns' gradient into grad_feats_rows[slot].  fp32 CUDA cores: the global
// optimisation decodes ~10^5 queries per step, tensor cores would not pay.
constexpr size_t kBwdSmem = (size_t)(DecMlp::kFloats + 2 * 64 * 64 + 64 * kDecThreads) * sizeof(float);

__device__ __forceinline__ unsigned long long relu_mask(const float (&acc)[64]) {
  unsigned long long mk = 0;
#pragma unroll
  for (int j = 0; j < 64; ++j) mk |= (unsigned long long)(acc[j] > 0.f) << j;
  return mk;
}

__global__ void __launch_bounds__(kDecThreads) decode_backward_simt_kernel(MapDev m, DecArgs a, const float* __restrict__ gW,
                                                                            const float* __restrict__ gRaw,
                                                                            const float* __restrict__ grad_out,
                                                                            float* __restrict__ grad_feats) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                                   // forward image (k-major)
  float* sR = smem + DecMlp::kFloats;                 // raw row-major W1 | W2
  float* sH = sR + 2 * 64 * 64 + threadIdx.x;
  {
    const float4* g4 = reinterpret_cast<const float4*>(gRaw + 64 * 32);      // skip W0 (64 x 32)
    float4* s4 = reinterpret_cast<float4*>(sR);
    for (int i = threadIdx.x; i < 2 * 64 * 64 / 4; i += blockDim.x) s4[i] = __ldg(g4 + i);
  }
  load_weights(sW, gW, DecMlp::kFloats);
  const int64_t q = (int64_t)blockIdx.x * kDecThreads + threadIdx.x;
  if (q >= a.n_queries) return;
  const float go = grad_out[q];
  if (go == 0.f) return;
  float c[3], fl[3], ce[3];
  query_coords(m, a, q, c);
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    fl[ax] = floorf(c[ax]);
    ce[ax] = ceilf(c[ax]);
  }
  const float wsum = corner_weight_sum(c, fl, ce);
  float minw = 3.0e38f;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    float nb[3], f8[8], wt;
    corner_of(k, fl, ce, nb);
    gather_corner(m, a, nb, f8, wt);
    minw = fminf(minw, wt);
  }
  if (!(minw >= a.min_pts)) return;                   // masked query: sdf = voxel_size, no gradient
  constexpr int S = kDecThreads;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    float nb[3];
    corner_of(k, fl, ce, nb);
    const int ix = (int)nb[0], iy = (int)nb[1], iz = (int)nb[2];
    int32_t slot = kEmpty;
    if (ix >= 0 && iy >= 0 && iz >= 0 && ix < m.g.n[0] && iy < m.g.n[1] && iz < m.g.n[2])
      slot = __ldg(m.table + ((int64_t)ix * m.g.nyz + iy * m.g.n[2] + iz));
    if (slot < 0 || slot >= a.n_rows) continue;       // a miss contributes no feature gradient
    float x[17], wt;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      const float l = __fsub_rn(c[ax], nb[ax]);
      x[ax] = l;
      x[3 + ax] = sinf(l);
      x[6 + ax] = cosf(l);
    }
    gather_corner(m, a, nb, x + 9, wt);
    // forward, recording the ReLU masks
    float acc[64];
    {
      const float4* b = reinterpret_cast<const float4*>(sW + DecMlp::kB0);
#pragma unroll
      for (int j4 = 0; j4 < 16; ++j4) {
        const float4 v = b[j4];
        acc[4 * j4] = v.x; acc[4 * j4 + 1] = v.y; acc[4 * j4 + 2] = v.z; acc[4 * j4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < 17; ++i) {
        const float4* w = reinterpret_cast<const float4*>(sW + DecMlp::kT0 + i * 64);
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const float4 v = w[j4];
          acc[4 * j4] = fmaf(v.x, x[i], acc[4 * j4]); acc[4 * j4 + 1] = fmaf(v.y, x[i], acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(v.z, x[i], acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(v.w, x[i], acc[4 * j4 + 3]);
        }
      }
    }
    const unsigned long long m1 = relu_mask(acc);
#pragma unroll
    for (int j = 0; j < 64; ++j) sH[j * S] = fmaxf(acc[j], 0.f);
    DecMlp::hidden_layer(sW + DecMlp::kT1, sH, S, acc);
    const unsigned long long m2 = relu_mask(acc);
#pragma unroll
    for (int j = 0; j < 64; ++j) sH[j * S] = fmaxf(acc[j], 0.f);
    DecMlp::hidden_layer(sW + DecMlp::kT2, sH, S, acc);
    const unsigned long long m3 = relu_mask(acc);
    // backward
    const float wn = __fdiv_rn(corner_weight(c, nb), wsum);
    const float dy = go * wn * m.g.vs;
#pragma unroll
    for (int j = 0; j < 64; ++j) sH[j * S] = ((m3 >> j) & 1ull) ? sW[DecMlp::kT3 + j] * dy : 0.f;   // d h3
    DecMlp::hidden_layer(sR + 64 * 64, sH, S, acc);                                                    // W2^T d h3
#pragma unroll
    for (int j = 0; j < 64; ++j) sH[j * S] = ((m2 >> j) & 1ull) ? acc[j] : 0.f;                       // d h2
    DecMlp::hidden_layer(sR, sH, S, acc);                                                              // W1^T d h2
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = ((m1 >> j) & 1ull) ? acc[j] : 0.f;                          // d h1
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      const float4* w = reinterpret_cast<const float4*>(sW + DecMlp::kT0 + (9 + f) * 64);             // W0[:, 9 + f]
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < 16; ++j4) {
        const float4 v = w[j4];
        s0 = fmaf(v.x, acc[4 * j4], s0); s1 = fmaf(v.y, acc[4 * j4 + 1], s1);
        s2 = fmaf(v.z, acc[4 * j4 + 2], s2); s3 = fmaf(v.w, acc[4 * j4 + 3], s3);
      }
      atomicAdd(grad_feats + (int64_t)slot * kFeat + f, (s0 + s1) + (s2 + s3));
    }
  }
}

}  // namespace bnv

using namespace bnv;

const float* bnv_internal_simt_weights(const bnv_mlp_t* mlp);
int bnv_internal_decode_tc(bnv_map_t* map, const bnv::DecArgs& a, const bnv_mlp_t* dec, cudaStream_t s);

static int decode_common(bnv_map_t* map, DecArgs& a, const bnv_mlp_t* dec, int mode, const float* tsdf,
                         const int32_t* tsdf_dims, cudaStream_t s) {
  if (!dec || dec->n_in != 17 || dec->n_out != 1) { set_error("decode: decoder MLP must be 17 -> 1"); return BNV_E_ARG; }
  if (a.n_rows < 0 || a.n_rows > map->d.cap || (a.n_rows > 0 && (!a.feats_rows || !a.weights_rows))) {
    set_error("decode: bad exported rows (n_rows=%lld)", (long long)a.n_rows);
    return BNV_E_ARG;
  }
  a.tsdf = tsdf;
  if (tsdf) {
    if (!tsdf_dims || tsdf_dims[0] <= 0 || tsdf_dims[1] <= 0 || tsdf_dims[2] <= 0) { set_error("decode: bad tsdf dims"); return BNV_E_ARG; }
    for (int i = 0; i < 3; ++i) {
      a.tsdf_dims[i] = tsdf_dims[i];
      a.nm1[i] = (float)(map->d.g.n[i] - 1);       // neighbor_coords / (self.n_xyz - 1)
      a.tm1[i] = (float)(tsdf_dims[i] - 1);
    }
  }
  if (a.n_queries == 0) return BNV_OK;
  if (mode == BNV_MLP_TC16) return bnv_internal_decode_tc(map, a, dec, s);
  if (mode != BNV_MLP_FP32) { set_error("decode: unknown MLP mode %d", mode); return BNV_E_ARG; }
  // per device and cheap: set on every call rather than cached per process
  BNV_CUDA(cudaFuncSetAttribute(decode_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDecSmem));
  decode_simt_kernel<<<(unsigned)((a.n_queries + kDecThreads - 1) / kDecThreads), kDecThreads, kDecSmem, s>>>(
      map->d, a, bnv_internal_simt_weights(dec));
  BNV_LAUNCH_CHECK("decode_simt_kernel");
  return BNV_OK;
}

extern "C" {

int bnv_decode_sdf_backward(bnv_map_t* map, const float* coords, int64_t n_queries, int is_coords,
                            const float* feats_rows, const float* weights_rows, int64_t n_rows, const bnv_mlp_t* dec,
                            int min_pts, const float* grad_out, float* grad_feats_rows, void* stream) {
  if (!map || n_queries < 0 || (n_queries > 0 && (!coords || !grad_out || !grad_feats_rows))) {
    set_error("bnv_decode_sdf_backward: bad argument");
    return BNV_E_ARG;
  }
  DecArgs a{};
  a.coords = coords;
  a.n_queries = n_queries;
  a.is_coords = is_coords;
  a.feats_rows = feats_rows;
  a.weights_rows = weights_rows;
  a.n_rows = n_rows;
  a.min_pts = (float)min_pts;
  if (!dec || dec->n_in != 17 || dec->n_out != 1) { set_error("decode backward: decoder MLP must be 17 -> 1"); return BNV_E_ARG; }
  if (n_rows <= 0 || n_rows > map->d.cap || !feats_rows || !weights_rows) { set_error("decode backward: bad exported rows"); return BNV_E_ARG; }
  if (n_queries == 0) return BNV_OK;
  // per device and cheap: set on every call rather than cached per process
  BNV_CUDA(cudaFuncSetAttribute(decode_backward_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
  decode_backward_simt_kernel<<<(unsigned)((n_queries + kDecThreads - 1) / kDecThreads), kDecThreads, kBwdSmem,
                                (cudaStream_t)stream>>>(map->d, a, dec->w32, dec->wraw, grad_out, grad_feats_rows);
  BNV_LAUNCH_CHECK("decode_backward_simt_kernel");
  return BNV_OK;
}

int bnv_decode_sdf(bnv_map_t* map, const float* coords, int64_t n_queries, int is_coords, const float* feats_rows,
                   const float* weights_rows, int64_t n_rows, const bnv_mlp_t* dec, int min_pts, int mode,
                   const float* tsdf, const int32_t* tsdf_dims, float* out_sdf, uint8_t* out_mask, void* stream) {
  if (!map || n_queries < 0 || (n_queries > 0 && (!coords || !out_sdf))) { set_error("bnv_decode_sdf: bad argument"); return BNV_E_ARG; }
  DecArgs a{};
  a.coords = coords;
  a.n_queries = n_queries;
  a.is_coords = is_coords;
  a.voxel_blocks = 0;
  a.feats_rows = feats_rows;
  a.weights_rows = weights_rows;
  a.n_rows = n_rows;
  a.min_pts = (float)min_pts;
  a.out_sdf = out_sdf;
  a.out_mask = out_mask;
  return decode_common(map, a, dec, mode, tsdf, tsdf_dims, (cudaStream_t)stream);
}

int bnv_decode_voxel_blocks(bnv_map_t* map, int64_t first, int64_t count, const float* feats_rows,
                            const float* weights_rows, int64_t n_rows, const bnv_mlp_t* dec, int min_pts, int mode,
                            const float* tsdf, const int32_t* tsdf_dims, float* out_sdf, void* stream) {
  if (!map || first < 0 || count < 0 || first + count > n_rows || (count > 0 && !out_sdf)) {
    set_error("bnv_decode_voxel_blocks: bad range [%lld, +%lld) of %lld rows", (long long)first, (long long)count, (long long)n_rows);
    return BNV_E_ARG;
  }
  DecArgs a{};
  a.coords = nullptr;
  a.n_queries = count * 27;
  a.is_coords = 1;
  a.voxel_blocks = 1;
  a.first_voxel = first;
  a.feats_rows = feats_rows;
  a.weights_rows = weights_rows;
  a.n_rows = n_rows;
  a.min_pts = (float)min_pts;
  a.out_sdf = out_sdf;
  a.out_mask = nullptr;
  return decode_common(map, a, dec, mode, tsdf, tsdf_dims, (cudaStream_t)stream);
}

}  // extern "C"
