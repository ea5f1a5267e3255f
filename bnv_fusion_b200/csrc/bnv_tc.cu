// tcgen05 tensor-core variants of the hot kernels (BNV_MLP_TC16): plain MLP forward, fused encode
// (backproject -> 8 corner rows -> encoder MLP -> scatter) and fused decode (8-corner gather ->
// decoder MLP -> trilinear blend + prior).  fp16 operands, fp32 accumulation in tensor memory.
// Building blocks and the data flow are described in bnv_tc.cuh.
#include <cuda_fp16.h>
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

constexpr int kNWG = 4;                       // row warpgroups (= TMEM slots) per CTA
constexpr int kThreads = kNWG * 128;
constexpr uint32_t kOnes = 0x3C003C00u;       // fp16x2 {1.0, 1.0}: tcnn pads the input with ones

struct alignas(16) TcSmem {
  TcShared<kNWG> sh;
};

__device__ __forceinline__ uint8_t* weights_smem(uint8_t* smem) { return smem + ((sizeof(TcSmem) + 127) / 128) * 128; }

static size_t tc_smem_bytes(int in_pad) { return ((sizeof(TcSmem) + 127) / 128) * 128 + weight_image(in_pad).bytes; }

static int tc_grid(int64_t n_tiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t need = (n_tiles + kNWG - 1) / kNWG;
  return (int)(need < sms ? (need < 1 ? 1 : need) : sms);
}

// ---- plain forward ------------------------------------------------------------------------------
template <int NIN, int INW, int NOUT>
__global__ void __launch_bounds__(kThreads, 1) mlp_forward_tc_kernel(const uint8_t* __restrict__ gW, int w_bytes,
                                                                     const float* __restrict__ x, int64_t n,
                                                                     float* __restrict__ y) {
  extern __shared__ __align__(128) uint8_t smem[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem);
  RowChain c = tc_setup<kNWG>(S.sh, weights_smem(smem), gW, w_bytes);
  const int wg = threadIdx.x >> 7, r = threadIdx.x & 127;
  const int64_t n_tiles = (n + 127) / 128;
  for (int64_t tile = (int64_t)blockIdx.x * kNWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * kNWG) {
    const int64_t i = tile * 128 + r;
    float xi[2 * INW];
#pragma unroll
    for (int k = 0; k < 2 * INW; ++k) xi[k] = 1.0f;
    if (i < n) {
#pragma unroll
      for (int k = 0; k < NIN; ++k) xi[k] = __ldg(x + i * NIN + k);
    }
    uint32_t in[INW];
#pragma unroll
    for (int k = 0; k < INW; ++k) in[k] = pack_f16x2(xi[2 * k], xi[2 * k + 1]);
    float out[NOUT];
    chain_run<INW, NOUT>(c, in, out);
    if (i < n) {
#pragma unroll
      for (int o = 0; o < NOUT; ++o) y[i * NOUT + o] = out[o];
    }
  }
  tc_teardown<kNWG>(S.sh);
}

// ---- fused encode -----------------------------------------------------------------------------------
// thread r of a warpgroup owns point (tile * 128 + r) and walks its 8 corner rows; a corner whose
// voxel is owned by nobody in the warpgroup (tile shard) is skipped by a warpgroup-uniform vote.
template <bool FROM_DEPTH>
__global__ void __launch_bounds__(kThreads, 1) encode_tc_kernel(MapDev m, EncSrc src, const uint8_t* __restrict__ gW,
                                                                int w_bytes, int64_t n_threads,
                                                                long long* __restrict__ stats) {
  extern __shared__ __align__(128) uint8_t smem[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem);
  RowChain c = tc_setup<kNWG>(S.sh, weights_smem(smem), gW, w_bytes);
  const int wg = threadIdx.x >> 7, r = threadIdx.x & 127;
  const GeomDev& g = m.g;
  const int64_t n_tiles = (n_threads + 127) / 128;
  for (int64_t tile = (int64_t)blockIdx.x * kNWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * kNWG) {
    const int64_t idx = tile * 128 + r;
    float p[6];
    bool valid = false;
    if (FROM_DEPTH) {
      if (idx < n_threads)
        valid = backproject_pixel(src.depth, src.cam, (int)(idx % src.cam.W), (int)(idx / src.cam.W), p);
    } else if (idx < n_threads) {
      valid = true;
#pragma unroll
      for (int j = 0; j < 6; ++j) p[j] = __ldg(src.pts6 + idx * 6 + j);
    }
    bool inb = valid;
#pragma unroll
    for (int a = 0; a < 3; ++a) inb = inb && (p[a] < g.hi[a]) && (p[a] > g.lo[a]);     // rule A1
    float cc[3], fl[3], ce[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      cc[a] = inb ? __fmul_rn(__fsub_rn(p[a], g.bmin[a]), g.inv_vs) : 0.f;             // rule A2
      fl[a] = floorf(cc[a]);
      ce[a] = ceilf(cc[a]);
    }
    const uint32_t nrm12 = pack_f16x2(inb ? p[4] : 0.f, inb ? p[5] : 0.f);
    const float nrm0 = inb ? p[3] : 0.f;
    // phase A: claim the 8 corner voxels' scratch rows (8 independent CAS round trips in flight)
    int32_t slot[8];
    int n_rows = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float nb[3];
      corner_of(k, fl, ce, nb);
      const int ix = (int)nb[0], iy = (int)nb[1], iz = (int)nb[2];
      slot[k] = -1;
      if (inb && owns(g, ix)) {
        slot[k] = claim_row(m, ix * g.nyz + iy * g.n[2] + iz, (int32_t)(idx * 8 + k));      // rule A5
        ++n_rows;
      }
    }
    // phase B: encoder MLP per corner row on the tensor core, accumulate into the claimed rows
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float nb[3];
      corner_of(k, fl, ce, nb);
      if (g.world > 1) {                       // warpgroup-uniform skip of corners nobody here owns
        int any;
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbarrier.cta.red.or.pred q, %2, 128, p;\n\t"
                     "selp.u32 %0, 1, 0, q;\n\t}"
                     : "=r"(any)
                     : "r"((int)(slot[k] >= 0)), "r"(c.bar_id)
                     : "memory");
        if (!any) continue;
      }
      float xr[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float rel = __fmul_rn(__fsub_rn(cc[a], nb[a]), g.vs);                     // rule A4
        xr[a] = __fmul_rn(rel, g.inv_vs);
      }
      // row = [x0 x1 x2 n0 n1 n2 | 1 x 10]  (tcnn pads the 6 inputs to 16 with ones)
      const uint32_t in[8] = {pack_f16x2(xr[0], xr[1]), pack_f16x2(xr[2], nrm0), nrm12, kOnes, kOnes, kOnes, kOnes, kOnes};
      float y[8];
      chain_run<8, 8>(c, in, y);
      if (slot[k] >= 0) add_row(m, slot[k], y);
    }
    const unsigned mv = __ballot_sync(0xffffffffu, valid);
    const unsigned mi = __ballot_sync(0xffffffffu, inb);
    int rr = n_rows;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rr += __shfl_xor_sync(0xffffffffu, rr, o);
    if ((threadIdx.x & 31) == 0 && (mv | mi)) {
      atomicAdd(reinterpret_cast<unsigned long long*>(stats + 0), (unsigned long long)__popc(mv));
      atomicAdd(reinterpret_cast<unsigned long long*>(stats + 1), (unsigned long long)rr);
      atomicAdd(reinterpret_cast<unsigned long long*>(stats + 4), (unsigned long long)__popc(mi));
    }
  }
  tc_teardown<kNWG>(S.sh);
}

// ---- fused decode -----------------------------------------------------------------------------------
template <int NWG>
__global__ void __launch_bounds__(NWG * 128, 1) decode_tc_kernel(MapDev m, DecArgs a, const uint8_t* __restrict__ gW,
                                                                 int w_bytes) {
  constexpr int kNWG = NWG;
  extern __shared__ __align__(128) uint8_t smem[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem);
  RowChain c = tc_setup<4>(S.sh, weights_smem(smem), gW, w_bytes);
  const int wg = threadIdx.x >> 7, r = threadIdx.x & 127;
  const int64_t n_tiles = (a.n_queries + 127) / 128;
  for (int64_t tile = (int64_t)blockIdx.x * kNWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * kNWG) {
    const int64_t q = tile * 128 + r;
    const bool live = q < a.n_queries;
    float cq[3] = {0.f, 0.f, 0.f};
    if (live) query_coords(m, a, q, cq);
    float fl[3], ce[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      fl[ax] = floorf(cq[ax]);
      ce[ax] = ceilf(cq[ax]);
    }
    const float wsum = corner_weight_sum(cq, fl, ce);
    float minw = 3.0e38f, sdf = 0.f, dsum = 0.f;
    // software pipeline: the gather of corner k+1 is in flight while corner k runs on the tensor core
    float feat[8], wt;
    {
      float nb[3];
      corner_of(0, fl, ce, nb);
      if (live && a.debug != 2) gather_corner(m, a, nb, feat, wt);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) feat[j] = 0.f;
        wt = 0.f;
      }
    }
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      float nb[3];
      corner_of(k, fl, ce, nb);
      float l[3], sn[3], cs[3];
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        l[ax] = __fsub_rn(cq[ax], nb[ax]);                                               // D1
        __sincosf(l[ax], &sn[ax], &cs[ax]);                                              // |l| <= 1
      }
      uint32_t in[16];
      in[0] = pack_f16x2(l[0], l[1]);
      in[1] = pack_f16x2(l[2], sn[0]);
      in[2] = pack_f16x2(sn[1], sn[2]);
      in[3] = pack_f16x2(cs[0], cs[1]);
      in[4] = pack_f16x2(cs[2], feat[0]);
      in[5] = pack_f16x2(feat[1], feat[2]);
      in[6] = pack_f16x2(feat[3], feat[4]);
      in[7] = pack_f16x2(feat[5], feat[6]);
      in[8] = pack_f16x2(feat[7], 1.0f);
#pragma unroll
      for (int j = 9; j < 16; ++j) in[j] = kOnes;
      minw = fminf(minw, wt);                                                            // D3
      if (k < 7) {                                                                       // prefetch next corner
        float nb2[3];
        corner_of(k + 1, fl, ce, nb2);
        if (live && a.debug != 2) gather_corner(m, a, nb2, feat, wt);
      }
      float y[1];
      if (a.debug == 1) y[0] = __uint_as_float(in[0] ^ in[5] ^ in[8]); else
      chain_run<16, 1>(c, in, y);                                                        // D7
      const float wn = __fdiv_rn(corner_weight(cq, nb), wsum);                           // D2
      sdf = __fadd_rn(sdf, __fmul_rn(__fmul_rn(y[0], m.g.vs), wn));                      // D4, D5
      if (a.tsdf) dsum = __fadd_rn(dsum, __fmul_rn(tsdf_nearest(a, m.g, nb), wn));       // D6
    }
    if (live) {
      bool mask;
      a.out_sdf[q] = finish_blend(sdf, dsum, minw, a, m.g.vs, &mask);
      if (a.out_mask) a.out_mask[q] = mask ? 1 : 0;
    }
  }
  tc_teardown<4>(S.sh);
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
        *reinterpret_cast<__half*>(base + byte) = __float2half_rn(W[(size_t)n * K + k]);   // row-major [out, in]
      }
    W += (size_t)N * K;
  }
  cudaError_t e = cudaMalloc(&mlp->w16, wi.bytes);
  if (e == cudaSuccess) e = cudaMemcpy(mlp->w16, img.data(), wi.bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cuda_fail(e, "pack_tc_weights");
  mlp->w16_bytes = wi.bytes;
  return BNV_OK;
}

template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
  BNV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return BNV_OK;
}

int bnv_internal_mlp_forward_tc(const bnv_mlp_t* mlp, const float* x, int64_t n, float* y, cudaStream_t s) {
  const size_t smem = tc_smem_bytes(mlp->in_pad);
  const int grid = tc_grid((n + 127) / 128);
  if (mlp->n_in == 6) {
    int rc = set_smem(mlp_forward_tc_kernel<6, 8, 8>, smem);
    if (rc) return rc;
    mlp_forward_tc_kernel<6, 8, 8><<<grid, kThreads, smem, s>>>((const uint8_t*)mlp->w16, (int)mlp->w16_bytes, x, n, y);
  } else {
    int rc = set_smem(mlp_forward_tc_kernel<17, 16, 1>, smem);
    if (rc) return rc;
    mlp_forward_tc_kernel<17, 16, 1><<<grid, kThreads, smem, s>>>((const uint8_t*)mlp->w16, (int)mlp->w16_bytes, x, n, y);
  }
  BNV_LAUNCH_CHECK("mlp_forward_tc_kernel");
  return BNV_OK;
}

int bnv_internal_encode_tc(bnv_map_t* map, const void* srcp, int from_depth, int64_t n_threads, const bnv_mlp_t* enc,
                           cudaStream_t s) {
  const EncSrc& src = *reinterpret_cast<const EncSrc*>(srcp);
  const size_t smem = tc_smem_bytes(enc->in_pad);
  const int grid = tc_grid((n_threads + 127) / 128);
  if (from_depth) {
    int rc = set_smem(encode_tc_kernel<true>, smem);
    if (rc) return rc;
    encode_tc_kernel<true><<<grid, kThreads, smem, s>>>(map->d, src, (const uint8_t*)enc->w16, (int)enc->w16_bytes,
                                                        n_threads, (long long*)map->stats);
  } else {
    int rc = set_smem(encode_tc_kernel<false>, smem);
    if (rc) return rc;
    encode_tc_kernel<false><<<grid, kThreads, smem, s>>>(map->d, src, (const uint8_t*)enc->w16, (int)enc->w16_bytes,
                                                         n_threads, (long long*)map->stats);
  }
  BNV_LAUNCH_CHECK("encode_tc_kernel");
  return BNV_OK;
}

int bnv_internal_decode_tc(bnv_map_t* map, const bnv::DecArgs& a, const bnv_mlp_t* dec, cudaStream_t s) {
  const size_t smem = tc_smem_bytes(dec->in_pad);
  const char* e = getenv("BNV_TC_NWG");
  const int nwg = e ? atoi(e) : 4;
  const int grid = tc_grid((a.n_queries + 127) / 128);
  int rc;
  if (nwg == 2) {
    rc = set_smem(decode_tc_kernel<2>, smem); if (rc) return rc;
    decode_tc_kernel<2><<<grid, 256, smem, s>>>(map->d, a, (const uint8_t*)dec->w16, (int)dec->w16_bytes);
  } else if (nwg == 3) {
    rc = set_smem(decode_tc_kernel<3>, smem); if (rc) return rc;
    decode_tc_kernel<3><<<grid, 384, smem, s>>>(map->d, a, (const uint8_t*)dec->w16, (int)dec->w16_bytes);
  } else {
    rc = set_smem(decode_tc_kernel<4>, smem); if (rc) return rc;
    decode_tc_kernel<4><<<grid, 512, smem, s>>>(map->d, a, (const uint8_t*)dec->w16, (int)dec->w16_bytes);
  }
  BNV_LAUNCH_CHECK("decode_tc_kernel");
  return BNV_OK;
}
