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
  static bool attr = false;
  if (!attr) {
    BNV_CUDA(cudaFuncSetAttribute(decode_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDecSmem));
    attr = true;
  }
  decode_simt_kernel<<<(unsigned)((a.n_queries + kDecThreads - 1) / kDecThreads), kDecThreads, kDecSmem, s>>>(
      map->d, a, bnv_internal_simt_weights(dec));
  BNV_LAUNCH_CHECK("decode_simt_kernel");
  return BNV_OK;
}

extern "C" {

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
