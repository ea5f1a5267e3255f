// Pieces of SparseVolume.decode_pts shared by the CUDA-core and tcgen05 decode kernels.
#pragma once
#include "bnv_common.cuh"

namespace bnv {

struct DecArgs {
  const float* coords;        // [Q,3] (unused when voxel_blocks)
  int64_t n_queries;
  int is_coords;
  int voxel_blocks;           // 1: queries are the 27 meshlize samples of active voxels
  int64_t first_voxel;
  const float* feats_rows;    // exported [n_rows, 8]
  const float* weights_rows;  // exported [n_rows]
  int64_t n_rows;
  float min_pts;
  const float* tsdf;          // nullable [Tx,Ty,Tz]
  int32_t tsdf_dims[3];
  float nm1[3];               // (float)(n_xyz - 1)
  float tm1[3];               // (float)(T - 1)
  float* out_sdf;
  uint8_t* out_mask;
};

// corner order of fusion/utils.get_neighbors (src/models/fusion/utils.py:98-167):
// k: (f,f,f) (c,f,f) (f,c,f) (f,f,c) (c,c,f) (c,f,c) (f,c,c) (c,c,c)
__device__ __forceinline__ void corner_of(int k, const float (&fl)[3], const float (&ce)[3], float (&nb)[3]) {
  const bool cx = (k == 1) | (k == 4) | (k == 5) | (k == 7);
  const bool cy = (k == 2) | (k == 4) | (k == 6) | (k == 7);
  const bool cz = (k == 3) | (k == 5) | (k == 6) | (k == 7);
  nb[0] = cx ? ce[0] : fl[0];
  nb[1] = cy ? ce[1] : fl[1];
  nb[2] = cz ? ce[2] : fl[2];
}

// voxel-unit coordinates of query q (decode_pts :792-793, or meshlize's id + {-0.5,0,0.5}^3 :720-731)
__device__ __forceinline__ void query_coords(const MapDev& m, const DecArgs& a, int64_t q, float (&c)[3]) {
  if (a.voxel_blocks) {
    const int64_t v = a.first_voxel + q / 27;
    const int s = (int)(q % 27);
    const int32_t flat = m.keys[v];
    const int32_t x = flat / m.g.nyz;
    const int32_t r = flat - x * m.g.nyz;
    const int32_t y = r / m.g.n[2];
    const int32_t z = r - y * m.g.n[2];
    c[0] = (float)x + 0.5f * (float)(s / 9 - 1);
    c[1] = (float)y + 0.5f * (float)((s / 3) % 3 - 1);
    c[2] = (float)z + 0.5f * (float)(s % 3 - 1);
    return;
  }
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const float v = __ldg(a.coords + q * 3 + ax);
    c[ax] = a.is_coords ? v : __fmul_rn(__fsub_rn(v, m.g.bmin[ax]), m.g.inv_vs);
  }
}

// _query_tensor (sparse_volume.py:625-659): features + fusion weight of one corner, zeros on a miss
__device__ __forceinline__ void gather_corner(const MapDev& m, const DecArgs& a, const float (&nb)[3],
                                              float* feat8, float& wt) {
  const int ix = (int)nb[0], iy = (int)nb[1], iz = (int)nb[2];
  int32_t slot = kEmpty;
  if (ix >= 0 && iy >= 0 && iz >= 0 && ix < m.g.n[0] && iy < m.g.n[1] && iz < m.g.n[2])
    slot = __ldg(m.table + ((int64_t)ix * m.g.nyz + iy * m.g.n[2] + iz));
  if (slot >= 0 && slot < a.n_rows) {
    const float4* f = reinterpret_cast<const float4*>(a.feats_rows + (int64_t)slot * kFeat);
    const float4 f0 = f[0], f1 = f[1];
    feat8[0] = f0.x; feat8[1] = f0.y; feat8[2] = f0.z; feat8[3] = f0.w;
    feat8[4] = f1.x; feat8[5] = f1.y; feat8[6] = f1.z; feat8[7] = f1.w;
    wt = a.weights_rows[slot];
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) feat8[j] = 0.f;
    wt = 0.f;
  }
}

// F.grid_sample(mode="nearest", padding_mode="zeros", align_corners=True) at one corner
// (sparse_volume.py:819-828; ATen grid_sampler_unnormalize op order, nearbyint = half-to-even)
__device__ __forceinline__ float tsdf_nearest(const DecArgs& a, const GeomDev& g, const float (&nb)[3]) {
  int idx[3];
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    float t = __fdiv_rn(nb[ax], a.nm1[ax]);
    t = __fmul_rn(t, 2.f);
    t = __fsub_rn(t, 1.f);
    t = __fadd_rn(t, 1.f);
    t = __fmul_rn(t, 0.5f);
    t = __fmul_rn(t, a.tm1[ax]);
    const float r = nearbyintf(t);
    if (!(r >= 0.f && r < (float)a.tsdf_dims[ax])) return 0.f;
    idx[ax] = (int)r;
  }
  return __ldg(a.tsdf + ((int64_t)idx[0] * a.tsdf_dims[1] + idx[1]) * a.tsdf_dims[2] + idx[2]);
}

// trilinear weight of corner nb for voxel-unit point c: prod over xyz of (1 - |c - nb|)  (D1, D2)
__device__ __forceinline__ float corner_weight(const float (&c)[3], const float (&nb)[3]) {
  const float tx = __fsub_rn(1.f, fabsf(__fsub_rn(c[0], nb[0])));
  const float ty = __fsub_rn(1.f, fabsf(__fsub_rn(c[1], nb[1])));
  const float tz = __fsub_rn(1.f, fabsf(__fsub_rn(c[2], nb[2])));
  return __fmul_rn(__fmul_rn(tx, ty), tz);
}

// the normaliser sum_k w_k (sparse_volume.py:814-815), corners in reference order
__device__ __forceinline__ float corner_weight_sum(const float (&c)[3], const float (&fl)[3], const float (&ce)[3]) {
  float wsum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float nb[3];
    corner_of(k, fl, ce, nb);
    const float w = corner_weight(c, nb);
    wsum = k == 0 ? w : __fadd_rn(wsum, w);
  }
  return wsum;
}

// D5, D6: validity mask, fallback value, prior
__device__ __forceinline__ float finish_blend(float sdf, float dsum, float minw, const DecArgs& a, float vs,
                                              bool* mask_out) {
  const bool mask = minw >= a.min_pts;                            // sparse_volume.py:809
  if (mask_out) *mask_out = mask;
  sdf = mask ? sdf : vs;                                          // :818
  if (a.tsdf) sdf = __fadd_rn(sdf, dsum);                         // :831-832 (also where mask is false)
  return sdf;
}

}  // namespace bnv
