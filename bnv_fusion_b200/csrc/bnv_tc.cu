// tcgen05 tensor-core variants of the hot kernels (BNV_MLP_TC16): plain MLP forward, fused encode
// (backproject -> 8 corner rows -> encoder MLP -> scatter) and fused decode (8-corner gather ->
// decoder MLP -> trilinear blend + prior).  fp16 operands, fp32 accumulation in tensor memory.
// Building blocks and the data flow are described in bnv_tc.cuh.
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
    for (int k = 0; k < 2 * INW; ++k) {
      const int src = NIN == 17 ? dec_perm(k) : enc_perm(k);     // column order of the packed W0
      xi[k] = (src < NIN && i < n) ? __ldg(x + i * NIN + src) : 1.0f;
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
                                                                long long* __restrict__ stats, int debug) {
  extern __shared__ __align__(128) uint8_t smem[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem);
  RowChain c = tc_setup<kNWG>(S.sh, weights_smem(smem), gW, w_bytes);
  const int wg = threadIdx.x >> 7, r = threadIdx.x & 127;
  const GeomDev& g = m.g;
  const int64_t n_tiles = (n_threads + 127) / 128;
  int st_valid = 0, st_inb = 0, st_rows = 0;           // frame statistics, flushed once per warp at the end
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
    const uint32_t nrm01 = pack_f16x2(inb ? p[3] : 0.f, inb ? p[4] : 0.f);
    const uint32_t nrm2o = pack_f16x2(inb ? p[5] : 0.f, 1.f);
    // phase A: claim the 8 corner voxels' scratch rows (8 independent CAS round trips in flight)
    int32_t slot[8];
    int n_rows = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float nb[3];
      corner_of(k, fl, ce, nb);
      const int ix = (int)nb[0], iy = (int)nb[1], iz = (int)nb[2];
      slot[k] = -1;
      if (inb && owns(g, ix, iy, iz)) {
        slot[k] = debug == 3 ? 0 : claim_row(m, ix * g.nyz + iy * g.n[2] + iz, (int32_t)(idx * 8 + k));      // rule A5
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
      // row = [x 1 | y 1 | z 1 | n0 n1 | n2 1 | 1 x 6]  (enc_perm: tcnn pads the 6 inputs to 16 with ones)
      const uint32_t in[8] = {pack_f16x2(xr[0], 1.f), pack_f16x2(xr[1], 1.f), pack_f16x2(xr[2], 1.f), nrm01, nrm2o, kOnes, kOnes, kOnes};
      float y[8];
      if (debug == 1) { for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(in[j & 3]); } else
      chain_run<8, 8>(c, in, y);
      if (debug == 5) add_row_f32_runs(m, slot[k], y); else
      if (slot[k] >= 0 && debug < 2) add_row_f32(m, slot[k], y);
    }
    st_valid += valid ? 1 : 0;
    st_inb += inb ? 1 : 0;
    st_rows += n_rows;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    st_valid += __shfl_xor_sync(0xffffffffu, st_valid, o);
    st_inb += __shfl_xor_sync(0xffffffffu, st_inb, o);
    st_rows += __shfl_xor_sync(0xffffffffu, st_rows, o);
  }
  if ((threadIdx.x & 31) == 0 && (st_valid | st_inb)) {
    atomicAdd(reinterpret_cast<unsigned long long*>(stats + 0), (unsigned long long)st_valid);
    atomicAdd(reinterpret_cast<unsigned long long*>(stats + 1), (unsigned long long)st_rows);
    atomicAdd(reinterpret_cast<unsigned long long*>(stats + 4), (unsigned long long)st_inb);
  }
  tc_teardown<kNWG>(S.sh);
}

// ---- fused decode -----------------------------------------------------------------------------------
// exported rows -> packed fp16 features (16 B per row) for the gather
__global__ void pack_rows_kernel(const float* __restrict__ feats, int64_t n, uint4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = reinterpret_cast<const float4*>(feats)[2 * i], b = reinterpret_cast<const float4*>(feats)[2 * i + 1];
  out[i] = make_uint4(pack_f16x2(a.x, a.y), pack_f16x2(a.z, a.w), pack_f16x2(b.x, b.y), pack_f16x2(b.z, b.w));
}

struct AxisPre {       // one axis of a query, floor (s = 0) and ceil (s = 1) flavours
  uint32_t w_ls[2];    // fp16x2 {l, sin l}
  uint32_t w_c1[2];    // fp16x2 {cos l, 1}
  float t[2];          // 1 - |l|
  int32_t tab[2];      // voxel index * table stride of this axis, or INT_MIN when outside the grid
  int32_t ts[2];       // TSDF-prior index * its stride, or INT_MIN when outside (nearest lookup)
};

template <int NWG>
__global__ void __launch_bounds__(NWG * 128, 1) decode_tc_kernel(MapDev m, DecArgs a, const uint4* __restrict__ packed,
                                                                 const uint8_t* __restrict__ gW, int w_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem);
  RowChain c = tc_setup<4>(S.sh, weights_smem(smem), gW, w_bytes);
  const int wg = threadIdx.x >> 7, r = threadIdx.x & 127;
  const int64_t n_tiles = (a.n_queries + 127) / 128;
  const GeomDev& g = m.g;
  constexpr int32_t kOut = INT_MIN;
  for (int64_t tile = (int64_t)blockIdx.x * NWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * NWG) {
    const int64_t q = tile * 128 + r;
    const bool live = q < a.n_queries;
    float cq[3] = {0.f, 0.f, 0.f};
    if (live) query_coords(m, a, q, cq);
    // ---- once per query: everything that depends on one axis only ---------------------------------
    AxisPre ax[3];
    const int32_t tstride[3] = {g.nyz, g.n[2], 1};
    const int32_t pstride[3] = {a.tsdf_dims[1] * a.tsdf_dims[2], a.tsdf_dims[2], 1};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float nbv[2] = {floorf(cq[d]), ceilf(cq[d])};
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float l = __fsub_rn(cq[d], nbv[s]);                                        // D1
        float sn, cs;
        __sincosf(l, &sn, &cs);                                                          // |l| <= 1
        ax[d].w_ls[s] = pack_f16x2(l, sn);
        ax[d].w_c1[s] = pack_f16x2(cs, 1.0f);
        ax[d].t[s] = __fsub_rn(1.f, fabsf(l));
        const int iv = (int)nbv[s];
        ax[d].tab[s] = (live && iv >= 0 && iv < g.n[d]) ? iv * tstride[d] : kOut;
        ax[d].ts[s] = kOut;
        if (a.tsdf) {                                                                    // grid_sample(nearest), D6
          float t = __fdiv_rn(nbv[s], a.nm1[d]);
          t = __fmul_rn(t, 2.f);
          t = __fsub_rn(t, 1.f);
          t = __fadd_rn(t, 1.f);
          t = __fmul_rn(t, 0.5f);
          t = __fmul_rn(t, a.tm1[d]);
          const float rr = nearbyintf(t);
          if (rr >= 0.f && rr < (float)a.tsdf_dims[d]) ax[d].ts[s] = (int)rr * pstride[d];
        }
      }
    }
    // corner k uses flavour (sx, sy, sz) = get_neighbors' order (src/models/fusion/utils.py:98-167)
    constexpr int SX[8] = {0, 1, 0, 0, 1, 1, 0, 1}, SY[8] = {0, 0, 1, 0, 1, 0, 1, 1}, SZ[8] = {0, 0, 0, 1, 0, 1, 1, 1};
    float wsum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float w = __fmul_rn(__fmul_rn(ax[0].t[SX[k]], ax[1].t[SY[k]]), ax[2].t[SZ[k]]);
      wsum = k == 0 ? w : __fadd_rn(wsum, w);                                            // D2 normaliser
    }
    // ---- 8 independent table lookups in flight (_query_tensor, D3) ---------------------------------
    int32_t slot[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int32_t tx = ax[0].tab[SX[k]], ty = ax[1].tab[SY[k]], tz = ax[2].tab[SZ[k]];
      slot[k] = kEmpty;
      if (tx != kOut && ty != kOut && tz != kOut) slot[k] = __ldg(m.table + ((int64_t)tx + ty + tz));
    }
    float minw = 3.0e38f, sdf = 0.f, dsum = 0.f;
    uint4 f_cur = make_uint4(0, 0, 0, 0), f_nxt = make_uint4(0, 0, 0, 0);
    float w_cur = 0.f, w_nxt = 0.f;
    if (slot[0] >= 0 && slot[0] < a.n_rows) {
      f_cur = __ldg(packed + slot[0]);
      w_cur = __ldg(a.weights_rows + slot[0]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t in[16] = {f_cur.x, f_cur.y, f_cur.z, f_cur.w,
                               ax[0].w_ls[SX[k]], ax[0].w_c1[SX[k]], ax[1].w_ls[SY[k]], ax[1].w_c1[SY[k]],
                               ax[2].w_ls[SZ[k]], ax[2].w_c1[SZ[k]], kOnes, kOnes, kOnes, kOnes, kOnes, kOnes};
      minw = fminf(minw, w_cur);                                                         // D3
      float y[1];
      chain_run<16, 1>(c, in, y, [&]() {
        // in the shadow of the first MMA round trip: fetch the next corner's features
        if (k < 7) {
          f_nxt = make_uint4(0, 0, 0, 0);
          w_nxt = 0.f;
          const int32_t s = slot[k < 7 ? k + 1 : 7];
          if (s >= 0 && s < a.n_rows) {
            f_nxt = __ldg(packed + s);
            w_nxt = __ldg(a.weights_rows + s);
          }
        }
      });                                                                                // D7
      const float wk = __fmul_rn(__fmul_rn(ax[0].t[SX[k]], ax[1].t[SY[k]]), ax[2].t[SZ[k]]);
      const float wn = __fdiv_rn(wk, wsum);                                              // D2
      sdf = __fadd_rn(sdf, __fmul_rn(__fmul_rn(y[0], g.vs), wn));                        // D4, D5
      if (a.tsdf) {
        const int32_t px = ax[0].ts[SX[k]], py = ax[1].ts[SY[k]], pz = ax[2].ts[SZ[k]];
        const float dl = (px != kOut && py != kOut && pz != kOut) ? __ldg(a.tsdf + ((int64_t)px + py + pz)) : 0.f;
        dsum = __fadd_rn(dsum, __fmul_rn(dl, wn));                                       // D6
      }
      f_cur = f_nxt;
      w_cur = w_nxt;
    }
    if (live) {
      bool mask;
      a.out_sdf[q] = finish_blend(sdf, dsum, minw, a, g.vs, &mask);
      if (a.out_mask) a.out_mask[q] = mask ? 1 : 0;
    }
  }
  tc_teardown<4>(S.sh);
}

// ---- factored decode of the meshlize sample blocks ------------------------------------------------------
// SparseVolume.meshlize samples id + {-0.5, 0, 0.5}^3 around every active voxel (sparse_volume.py:717-731),
// so every (query, corner) row of decode_pts is MLP(l, feat_V) with V a voxel and l in {-0.5, 0, 0.5}^3:
// only 27 distinct rows per voxel exist, each shared by up to 8 queries of neighbouring voxels.  The
// factored path evaluates G[V][l] once on the tensor cores (27 rows per exported voxel + one "miss" voxel
// with zero features, rule D7) and then blends per sample with the reference's op order (D2-D6).  Same MLP
// rows, same blend => bit-identical to decode_tc_kernel on the same coordinates, with 8x fewer MLP rows.
__global__ void __launch_bounds__(kThreads, 1) gtable_tc_kernel(const uint4* __restrict__ packed, int64_t n_rows,
                                                                const uint8_t* __restrict__ gW, int w_bytes,
                                                                float* __restrict__ G) {
  extern __shared__ __align__(128) uint8_t smem[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem);
  RowChain c = tc_setup<kNWG>(S.sh, weights_smem(smem), gW, w_bytes);
  const int wg = threadIdx.x >> 7, r = threadIdx.x & 127;
  const int64_t total = (n_rows + 1) * 27;                       // voxel n_rows = the miss voxel
  const int64_t n_tiles = (total + 127) / 128;
  uint32_t w_ls[3], w_c1[3];                                     // l = -0.5, 0, +0.5
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float l = 0.5f * (float)(d - 1);
    float sn, cs;
    __sincosf(l, &sn, &cs);
    w_ls[d] = pack_f16x2(l, sn);
    w_c1[d] = pack_f16x2(cs, 1.0f);
  }
  for (int64_t tile = (int64_t)blockIdx.x * kNWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * kNWG) {
    const int64_t row = tile * 128 + r;
    const int64_t v = row / 27;
    const int li = (int)(row - v * 27);
    uint4 f = make_uint4(0, 0, 0, 0);
    if (row < total && v < n_rows) f = __ldg(packed + v);
    const int dx = li / 9, dy = (li / 3) % 3, dz = li % 3;
    const uint32_t in[16] = {f.x, f.y, f.z, f.w, w_ls[dx], w_c1[dx], w_ls[dy], w_c1[dy], w_ls[dz], w_c1[dz],
                             kOnes, kOnes, kOnes, kOnes, kOnes, kOnes};
    float y[1];
    chain_run<16, 1>(c, in, y);
    if (row < total) G[row] = y[0];
  }
  tc_teardown<kNWG>(S.sh);
}

// per sample: 8 corner lookups into G + trilinear blend + mask + prior (rules D1-D6)
__global__ void __launch_bounds__(256) blend_blocks_kernel(MapDev m, DecArgs a, const float* __restrict__ G) {
  const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (q >= a.n_queries) return;
  const GeomDev& g = m.g;
  const int64_t v = a.first_voxel + q / 27;
  const int s = (int)(q % 27);
  const int32_t flat0 = m.keys[v];
  const int id[3] = {flat0 / g.nyz, (flat0 % g.nyz) / g.n[2], flat0 % g.n[2]};
  const int o[3] = {s / 9 - 1, (s / 3) % 3 - 1, s % 3 - 1};     // sample offset / 0.5
  // per axis, floor (j = 0) and ceil (j = 1) corner: voxel index, l digit (0: -0.5, 1: 0, 2: +0.5), 1 - |l|
  int nbv[3][2], ld[3][2];
  float tt[3][2], nbf[3][2];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    nbv[d][0] = o[d] < 0 ? id[d] - 1 : id[d];
    nbv[d][1] = o[d] > 0 ? id[d] + 1 : id[d];
    ld[d][0] = o[d] == 0 ? 1 : 2;
    ld[d][1] = o[d] == 0 ? 1 : 0;
    tt[d][0] = tt[d][1] = o[d] == 0 ? 1.0f : 0.5f;
    nbf[d][0] = (float)nbv[d][0];
    nbf[d][1] = (float)nbv[d][1];
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
    const int ix = nbv[0][SX[k]], iy = nbv[1][SY[k]], iz = nbv[2][SZ[k]];
    int64_t gv = a.n_rows;                                        // miss voxel
    float wt = 0.f;
    if (ix >= 0 && iy >= 0 && iz >= 0 && ix < g.n[0] && iy < g.n[1] && iz < g.n[2]) {
      const int32_t slot = __ldg(m.table + ((int64_t)ix * g.nyz + iy * g.n[2] + iz));
      if (slot >= 0 && slot < a.n_rows) {
        gv = slot;
        wt = __ldg(a.weights_rows + slot);
      }
    }
    minw = fminf(minw, wt);
    const float y = __ldg(G + gv * 27 + (ld[0][SX[k]] * 9 + ld[1][SY[k]] * 3 + ld[2][SZ[k]]));
    const float wn = __fdiv_rn(__fmul_rn(__fmul_rn(tt[0][SX[k]], tt[1][SY[k]]), tt[2][SZ[k]]), wsum);
    sdf = __fadd_rn(sdf, __fmul_rn(__fmul_rn(y, g.vs), wn));
    if (a.tsdf) {
      const float nb[3] = {nbf[0][SX[k]], nbf[1][SY[k]], nbf[2][SZ[k]]};
      dsum = __fadd_rn(dsum, __fmul_rn(tsdf_nearest(a, g, nb), wn));
    }
  }
  bool mask;
  a.out_sdf[q] = finish_blend(sdf, dsum, minw, a, g.vs, &mask);
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
  const char* e = getenv("BNV_DEBUG_ENCODE");     // profiling experiments only
  const int dbg = e ? atoi(e) : 0;
  const size_t smem = tc_smem_bytes(enc->in_pad);
  const int grid = tc_grid((n_threads + 127) / 128);
  if (from_depth) {
    int rc = set_smem(encode_tc_kernel<true>, smem);
    if (rc) return rc;
    encode_tc_kernel<true><<<grid, kThreads, smem, s>>>(map->d, src, (const uint8_t*)enc->w16, (int)enc->w16_bytes,
                                                        n_threads, (long long*)map->stats, dbg);
  } else {
    int rc = set_smem(encode_tc_kernel<false>, smem);
    if (rc) return rc;
    encode_tc_kernel<false><<<grid, kThreads, smem, s>>>(map->d, src, (const uint8_t*)enc->w16, (int)enc->w16_bytes,
                                                         n_threads, (long long*)map->stats, dbg);
  }
  BNV_LAUNCH_CHECK("encode_tc_kernel");
  return BNV_OK;
}

int bnv_internal_decode_tc(bnv_map_t* map, const bnv::DecArgs& a, const bnv_mlp_t* dec, cudaStream_t s) {
  const size_t smem = tc_smem_bytes(dec->in_pad);
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
    int rc = set_smem(gtable_tc_kernel, smem);
    if (rc) return rc;
    const int64_t tiles = ((a.n_rows + 1) * 27 + 127) / 128;
    gtable_tc_kernel<<<tc_grid(tiles), kThreads, smem, s>>>((const uint4*)map->dec_pack, a.n_rows, (const uint8_t*)dec->w16,
                                                            (int)dec->w16_bytes, (float*)map->gtable);
    BNV_LAUNCH_CHECK("gtable_tc_kernel");
    blend_blocks_kernel<<<(unsigned)((a.n_queries + 255) / 256), 256, 0, s>>>(map->d, a, (const float*)map->gtable);
    BNV_LAUNCH_CHECK("blend_blocks_kernel");
    return BNV_OK;
  }
  const char* e = getenv("BNV_TC_NWG");      // profiling experiments only
  const int nwg = e ? atoi(e) : 4;
  const int grid = tc_grid((a.n_queries + 127) / 128);
  int rc;
  if (nwg == 2) {
    rc = set_smem(decode_tc_kernel<2>, smem); if (rc) return rc;
    decode_tc_kernel<2><<<grid, 256, smem, s>>>(map->d, a, (const uint4*)map->dec_pack, (const uint8_t*)dec->w16, (int)dec->w16_bytes);
  } else if (nwg == 3) {
    rc = set_smem(decode_tc_kernel<3>, smem); if (rc) return rc;
    decode_tc_kernel<3><<<grid, 384, smem, s>>>(map->d, a, (const uint4*)map->dec_pack, (const uint8_t*)dec->w16, (int)dec->w16_bytes);
  } else {
    rc = set_smem(decode_tc_kernel<4>, smem); if (rc) return rc;
    decode_tc_kernel<4><<<grid, 512, smem, s>>>(map->d, a, (const uint4*)map->dec_pack, (const uint8_t*)dec->w16, (int)dec->w16_bytes);
  }
  BNV_LAUNCH_CHECK("decode_tc_kernel");
  return BNV_OK;
}
