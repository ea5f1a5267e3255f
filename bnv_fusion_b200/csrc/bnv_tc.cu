// Host side of the tensor-core mode (BNV_MLP_TC16): weight packing into the UMMA shared-memory layout,
// dispatch to the chain kernels (bnv_tc_chain.cu), and the two non-MMA helpers of the decode path
// (fp16 repack of the exported rows, per-sample blend of the factored meshlize decode).
#include <cuda_fp16.h>
#include <limits.h>
#include <stdlib.h>

#include <vector>

#include "bnv_common.cuh"
#include "bnv_decode_common.cuh"
#include "bnv_frame.cuh"
#include "bnv_tc.cuh"

using namespace bnv;
using namespace bnv::tc;

namespace bnv {
namespace tc {

// exported rows -> packed fp16 features (16 B per row) for the gather
__global__ void pack_rows_kernel(const float* __restrict__ feats, int64_t n, uint4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = reinterpret_cast<const float4*>(feats)[2 * i], b = reinterpret_cast<const float4*>(feats)[2 * i + 1];
  out[i] = make_uint4(pack_f16x2(a.x, a.y), pack_f16x2(a.z, a.w), pack_f16x2(b.x, b.y), pack_f16x2(b.z, b.w));
}

// Per sample: 8 corner lookups into G + trilinear blend + mask + prior (rules D1-D6).  One warp per active
// voxel, lane s = sample s AND neighbour s of the voxel's 3x3x3 neighbourhood: the 27 table / weight lookups
// are done once per voxel and handed to the samples by shuffle (every corner of every sample is one of the
// 27 neighbours), instead of 8 x 27 lookups.  Arithmetic and corner order are those of the per-query kernel.
__global__ void __launch_bounds__(256) blend_blocks_kernel(MapDev m, DecArgs a, const float* __restrict__ G) {
  const int lane = threadIdx.x & 31;
  const int64_t vl = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // voxel of this warp
  const int64_t n_vox = a.n_queries / 27;
  if (vl >= n_vox) return;                                                         // warp-uniform
  const GeomDev& g = m.g;
  const int64_t v = a.first_voxel + vl;
  const int32_t flat0 = m.keys[v];
  const int id[3] = {flat0 / g.nyz, (flat0 % g.nyz) / g.n[2], flat0 % g.n[2]};
  const int s = lane < 27 ? lane : 0;
  const int o[3] = {s / 9 - 1, (s / 3) % 3 - 1, s % 3 - 1};     // sample offset / 0.5 == neighbour offset
  // ---- as neighbour `lane`: row in G (miss voxel = n_rows) and fusion weight -------------------------------
  int32_t my_gv = (int32_t)a.n_rows;
  float my_wt = 0.f;
  {
    const int ix = id[0] + o[0], iy = id[1] + o[1], iz = id[2] + o[2];
    if (lane < 27 && ix >= 0 && iy >= 0 && iz >= 0 && ix < g.n[0] && iy < g.n[1] && iz < g.n[2]) {
      const int32_t slot = __ldg(m.table + ((int64_t)ix * g.nyz + iy * g.n[2] + iz));
      if (slot >= 0 && slot < a.n_rows) {
        my_gv = slot;
        my_wt = __ldg(a.weights_rows + slot);
      }
    }
  }
  // ---- as sample `lane` -------------------------------------------------------------------------------------
  // per axis, floor (j = 0) and ceil (j = 1) corner: neighbour offset, l digit (0: -0.5, 1: 0, 2: +0.5), 1 - |l|
  int nbo[3][2], ld[3][2];
  float tt[3][2], nbf[3][2];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    nbo[d][0] = o[d] < 0 ? -1 : 0;
    nbo[d][1] = o[d] > 0 ? 1 : 0;
    ld[d][0] = o[d] == 0 ? 1 : 2;
    ld[d][1] = o[d] == 0 ? 1 : 0;
    tt[d][0] = tt[d][1] = o[d] == 0 ? 1.0f : 0.5f;
    nbf[d][0] = (float)(id[d] + nbo[d][0]);
    nbf[d][1] = (float)(id[d] + nbo[d][1]);
  }
  constexpr int SX[8] = {0, 1, 0, 0, 1, 1, 0, 1}, SY[8] = {0, 0, 1, 0, 1, 0, 1, 1}, SZ[8] = {0, 0, 0, 1, 0, 1, 1, 1};
  float wsum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float w = __fmul_rn(__fmul_rn(tt[0][SX[k]], tt[1][SY[k]]), tt[2][SZ[k]]);
    wsum = k == 0 ? w : __fadd_rn(wsum, w);
  }
  float minw = 3.0e38f, sdf = 0.f, dsum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int nb_lane = (nbo[0][SX[k]] + 1) * 9 + (nbo[1][SY[k]] + 1) * 3 + (nbo[2][SZ[k]] + 1);
    const int32_t gv = __shfl_sync(0xffffffffu, my_gv, nb_lane);
    const float wt = __shfl_sync(0xffffffffu, my_wt, nb_lane);
    minw = fminf(minw, wt);
    const float y = __ldg(G + (int64_t)gv * 27 + (ld[0][SX[k]] * 9 + ld[1][SY[k]] * 3 + ld[2][SZ[k]]));
    const float wn = __fdiv_rn(__fmul_rn(__fmul_rn(tt[0][SX[k]], tt[1][SY[k]]), tt[2][SZ[k]]), wsum);
    sdf = __fadd_rn(sdf, __fmul_rn(__fmul_rn(y, g.vs), wn));
    if (a.tsdf) {
      const float nb[3] = {nbf[0][SX[k]], nbf[1][SY[k]], nbf[2][SZ[k]]};
      dsum = __fadd_rn(dsum, __fmul_rn(tsdf_nearest(a, g, nb), wn));
    }
  }
  if (lane < 27) {
    bool mask;
    a.out_sdf[vl * 27 + lane] = finish_blend(sdf, dsum, minw, a, g.vs, &mask);
  }
}

}  // namespace tc
}  // namespace bnv

// ---- host side ------------------------------------------------------------------------------------------
int bnv_internal_pack_tc_weights(bnv_mlp_t* mlp, const float* params) {
  const WeightImage wi = weight_image(mlp->in_pad);
  std::vector<__half> img((size_t)wi.bytes / 2);
  const float* W = params;
  memcpy(reinterpret_cast<uint8_t*>(img.data()) + wi.off_w3f32,
         params + (size_t)64 * mlp->in_pad + 2 * 64 * 64, 64 * sizeof(float));   // W3 row 0
  for (int l = 0; l < 4; ++l) {
    const int K = wi.k[l], N = wi.n[l];
    const int lbo = (N / 8) * 128;
    uint8_t* base = reinterpret_cast<uint8_t*>(img.data()) + wi.off[l];
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        const size_t byte = (size_t)(k / 8) * lbo + (size_t)(n / 8) * 128 + (size_t)(n % 8) * 16 + (size_t)(k % 8) * 2;
        const int src = l > 0 ? k : (mlp->n_in == 17 ? dec_perm(k) : enc_perm(k));       // first layer: permuted inputs
        *reinterpret_cast<__half*>(base + byte) = __float2half_rn(W[(size_t)n * K + src]); // row-major [out, in]
      }
    W += (size_t)N * K;
  }
  cudaError_t e = cudaMalloc(&mlp->w16, wi.bytes);
  if (e == cudaSuccess) e = cudaMemcpy(mlp->w16, img.data(), wi.bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cuda_fail(e, "pack_tc_weights");
  mlp->w16_bytes = wi.bytes;
  return BNV_OK;
}

// warp-specialised chain kernels (bnv_tc_ws.cu)
int bnv_internal_mlp_forward_ws(const bnv_mlp_t* mlp, const float* x, int64_t n, float* y, cudaStream_t s);
int bnv_internal_encode_ws(bnv_map_t* map, int64_t max_records, const bnv_mlp_t* enc, cudaStream_t s);
int bnv_internal_decode_ws(bnv_map_t* map, const bnv::DecArgs& a, const bnv_mlp_t* dec, cudaStream_t s);
int bnv_internal_gtable_ws(bnv_map_t* map, int64_t n_rows, const bnv_mlp_t* dec, cudaStream_t s);

int bnv_internal_mlp_forward_tc(const bnv_mlp_t* mlp, const float* x, int64_t n, float* y, cudaStream_t s) {
  return bnv_internal_mlp_forward_ws(mlp, x, n, y, s);
}

int bnv_internal_encode_tc(bnv_map_t* map, int64_t max_records, const bnv_mlp_t* enc, cudaStream_t s) {
  return bnv_internal_encode_ws(map, max_records, enc, s);
}

int bnv_internal_decode_tc(bnv_map_t* map, const bnv::DecArgs& a, const bnv_mlp_t* dec, cudaStream_t s) {
  // gather-friendly copy of the exported rows: fp16x8 features, 16 B per row
  if (a.n_rows > 0) {
    pack_rows_kernel<<<(unsigned)((a.n_rows + 255) / 256), 256, 0, s>>>(a.feats_rows, a.n_rows, (uint4*)map->dec_pack);
    BNV_LAUNCH_CHECK("pack_rows_kernel");
  }
  if (a.voxel_blocks && !getenv("BNV_DECODE_DIRECT")) {
    // factored meshlize decode: G table on the tensor cores, then the per-sample blend
    if ((size_t)(a.n_rows + 1) * 27 * sizeof(float) > map->gtable_bytes) {
      set_error("decode: G table of %lld rows exceeds the map's capacity", (long long)a.n_rows);
      return BNV_E_CAPACITY;
    }
    int rc = bnv_internal_gtable_ws(map, a.n_rows, dec, s);
    if (rc) return rc;
    blend_blocks_kernel<<<(unsigned)((a.n_queries / 27 + 7) / 8), 256, 0, s>>>(map->d, a, (const float*)map->gtable);   // one warp per voxel
    BNV_LAUNCH_CHECK("blend_blocks_kernel");
    return BNV_OK;
  }
  return bnv_internal_decode_ws(map, a, dec, s);
}
