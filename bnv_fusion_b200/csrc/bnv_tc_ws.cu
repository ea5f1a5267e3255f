// Warp-specialised tcgen05 kernels (bnv_tc_ws.cuh): the encoder MLP over the frame's point records and the fused SDF
// decode.  Epilogue warpgroups run nothing but the MLP chain; helper warpgroups feed it (record / table / feature
// gathers, input rows into TMEM) and drain it (outputs out of TMEM, scatter-add or trilinear blend).
//
// Reference semantics are those of the single-role kernels they replace (bnv_tc_chain.cu):
//   encode: local_point_fusion.py:81-165 rules A2-A6 per (point, corner) row, scatter_mean sums
//   decode: sparse_volume.py:768-833 rules D1-D7
#include <cuda_fp16.h>
#include <limits.h>

#include "bnv_common.cuh"
#include "bnv_decode_common.cuh"
#include "bnv_frame.cuh"
#include "bnv_tc_ws.cuh"

using namespace bnv;
using namespace bnv::ws;

namespace bnv {
namespace wsk {

constexpr int kRows = kNC * 128;              // helper threads (= rows in flight) per CTA
constexpr uint32_t kOnes = 0x3C003C00u;       // fp16x2 {1.0, 1.0}: tcnn pads the input with ones

struct alignas(16) SmemBase {
  WsShared sh;
};
__device__ __forceinline__ uint8_t* weights_smem(uint8_t* smem) { return smem + ((sizeof(SmemBase) + 127) / 128) * 128; }
static size_t weights_off() { return ((sizeof(SmemBase) + 127) / 128) * 128; }

static int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
static int grid_for(int64_t n_items) {
  const int sms = sm_count();
  const int64_t need = (n_items + kNC - 1) / kNC;
  return (int)(need < sms ? (need < 1 ? 1 : need) : sms);
}

// ---- encode ---------------------------------------------------------------------------------------------------------
// Work unit = (128-record tile, corner); chain j of J takes the contiguous unit range [U j / J, U (j + 1) / J).
__device__ __forceinline__ void enc_input(int k, const float (&cc)[3], const float (&fl)[3], const float (&ce)[3], float vs,
                                          float inv_vs, uint32_t nrm01, uint32_t nrm2o, uint32_t (&in)[8]) {
  float nb[3];
  corner_of(k, fl, ce, nb);
  float xr[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float rel = __fmul_rn(__fsub_rn(cc[a], nb[a]), vs);                         // rule A4
    xr[a] = __fmul_rn(rel, inv_vs);
  }
  // row = [x 1 | y 1 | z 1 | n0 n1 | n2 1 | 1 x 6]  (enc_perm: tcnn pads the 6 inputs to 16 with ones)
  in[0] = pack_f16x2(xr[0], 1.f);
  in[1] = pack_f16x2(xr[1], 1.f);
  in[2] = pack_f16x2(xr[2], 1.f);
  in[3] = nrm01;
  in[4] = nrm2o;
  in[5] = in[6] = in[7] = kOnes;
}

__global__ void __launch_bounds__(kThreads, 1) encode_ws_kernel(MapDev m, const uint8_t* __restrict__ gW, int w_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  SmemBase& S = *reinterpret_cast<SmemBase*>(smem);
  Role c = ws_setup(S.sh, weights_smem(smem), gW, w_bytes);
  grid_dependency_wait();                               // the prepass (setup above overlapped its tail)
  if (!c.helper) {
    e_run<8>(c);
  } else {
    const GeomDev& g = m.g;
    const int r = c.row, warp_in_wg = r >> 5;
    const int64_t n_rec = m.ctr[4];                     // point records written by frame_prepass_kernel
    const int64_t n_units = ((n_rec + 127) / 128) * 8;
    const int64_t chain = (int64_t)blockIdx.x * kNC + c.chain, n_chains = (int64_t)gridDim.x * kNC;
    const int64_t u_end = n_units * (chain + 1) / n_chains;
    HelperSeq seq;
    auto consume = [&](int row) {
      float y[8];
      h_read_out<8>(c, y);
      if (row >= 0) add_row_f32(m, row, y);
    };
    int flip = 0;
    for (int64_t u = n_units * chain / n_chains; u < u_end;) {
      const int64_t tile = u >> 3;
      const int k0 = (int)(u & 7);
      const int k1 = (int)(u_end - u < (int64_t)(8 - k0) ? k0 + (u_end - u) : 8);
      u += k1 - k0;
      const int64_t idx = tile * 128 + r;
      float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
      if (idx < n_rec) {
        const float4* r4 = reinterpret_cast<const float4*>(m.prec + (size_t)idx * 8);
        ra = __ldg(r4);
        rb = __ldg(r4 + 1);
      }
      const float cc[3] = {ra.x, ra.y, ra.z};
      float fl[3], ce[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        fl[a] = floorf(cc[a]);
        ce[a] = ceilf(cc[a]);
      }
      const uint32_t nrm01 = pack_f16x2(ra.w, rb.x);
      const uint32_t nrm2o = pack_f16x2(rb.y, 1.f);
      // corners of this chain's unit range that this rank owns (the prepass stored the ownership mask)
      const uint32_t range = ((1u << k1) - 1u) & ~((1u << k0) - 1u);
      const uint32_t own = (idx < n_rec ? (uint32_t)__float_as_int(rb.z) : 0u) & range;
      const int frame = record_frame(rb.w);               // position of the record's frame in its batch
      int32_t rows[8];                                    // dense scratch rows of the owned corners (8 reads in flight)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        rows[k] = -1;
        if ((own >> k) & 1u) {
          float nb[3];
          corner_of(k, fl, ce, nb);
          rows[k] = scratch_row_of(m, (int)nb[0] * g.nyz + (int)nb[1] * g.n[2] + (int)nb[2], frame);   // rule A5
        }
      }
      // corners to run: all of [k0, k1) on one GPU; in the tile shard only those somebody in the warpgroup owns
      uint32_t live = range;
      if (g.world > 1) {
        const uint32_t wown = __reduce_or_sync(0xffffffffu, own);
        if ((r & 31) == 0) S.sh.live[c.chain][flip][warp_in_wg] = wown;
        wg_sync(c.bar_hw);
        live = S.sh.live[c.chain][flip][0] | S.sh.live[c.chain][flip][1] | S.sh.live[c.chain][flip][2] | S.sh.live[c.chain][flip][3];
        flip ^= 1;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!((live >> k) & 1u)) continue;                // warpgroup-uniform
        seq.emit(
            c, rows[k],
            [&]() {
              uint32_t in[8];
              enc_input(k, cc, fl, ce, g.vs, g.inv_vs, nrm01, nrm2o, in);
              h_stage<8>(c, in);
            },
            consume);
      }
    }
    seq.finish(c, consume);
  }
  ws_teardown(S.sh);
}

// ---- decode ---------------------------------------------------------------------------------------------------------
// Per-query state of the helper threads in shared memory ([word][helper thread], conflict-free, thread-private):
// everything that depends on one axis only exists in a floor (s = 0) and a ceil (s = 1) flavour computed once per
// query; a corner row is then 4 gathered words + 6 selected words + 6 constants.
struct DecState {
  uint32_t w_ls[3][2][kRows];      // fp16x2 {l, sin l}
  uint32_t w_c1[3][2][kRows];      // fp16x2 {cos l, 1}
  float t[3][2][kRows];            // 1 - |l|
  int32_t ts[3][2][kRows];         // TSDF-prior index * its stride, or INT_MIN when outside (nearest lookup)
  int32_t slot[8][kRows];          // table lookup of corner k (_query_tensor)
};

__global__ void __launch_bounds__(kThreads, 1) decode_ws_kernel(MapDev m, DecArgs a, const uint4* __restrict__ packed,
                                                                 const uint8_t* __restrict__ gW, int w_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  SmemBase& S = *reinterpret_cast<SmemBase*>(smem);
  Role c = ws_setup(S.sh, weights_smem(smem), gW, w_bytes);
  if (!c.helper) {
    e_run<16>(c);
  } else {
    DecState& Q = *reinterpret_cast<DecState*>(weights_smem(smem) + ((w_bytes + 127) / 128) * 128);
    const int hid = c.chain * 128 + c.row;              // helper thread index in the state arrays
    const int64_t n_tiles = (a.n_queries + 127) / 128;
    const GeomDev& g = m.g;
    constexpr int32_t kOut = INT_MIN;
    const bool has_prior = a.tsdf != nullptr;
    auto weight_of = [&](int k) {                                                          // D2, corner k
      return __fmul_rn(__fmul_rn(Q.t[0][corner_sx(k)][hid], Q.t[1][corner_sy(k)][hid]), Q.t[2][corner_sz(k)][hid]);
    };
    auto prior_of = [&](int k) {                                                           // D6, corner k
      const int32_t px = Q.ts[0][corner_sx(k)][hid], py = Q.ts[1][corner_sy(k)][hid], pz = Q.ts[2][corner_sz(k)][hid];
      return (px != kOut && py != kOut && pz != kOut) ? __ldg(a.tsdf + ((int64_t)px + py + pz)) : 0.f;
    };
    auto gather = [&](int k, uint4& f, float& w) {                                         // D3
      f = make_uint4(0, 0, 0, 0);
      w = 0.f;
      const int32_t s = Q.slot[k][hid];
      if (s >= 0 && s < a.n_rows) {
        f = __ldg(packed + s);
        w = __ldg(a.weights_rows + s);
      }
    };
    // blend state: outputs come back two items after their rows were staged, so the weights of the two corners in
    // flight wait in a two-entry FIFO and the query meta data of two tiles (by tile parity) is kept
    float fw0 = 0.f, fd0 = 0.f, fw1 = 0.f, fd1 = 0.f;
    int fcount = 0;
    float sdf = 0.f, dsum = 0.f;
    int64_t mq[2] = {0, 0};
    float mminw[2] = {0.f, 0.f};
    bool mlive[2] = {false, false};
    auto consume = [&](int tag) {
      const int k = tag & 7, par = (tag >> 3) & 1;
      float y[1];
      h_read_out<1>(c, y);
      const float wn = fw0, dl = fd0;
      fw0 = fw1;
      fd0 = fd1;
      --fcount;
      if (k == 0) sdf = dsum = 0.f;
      sdf = __fadd_rn(sdf, __fmul_rn(__fmul_rn(y[0], g.vs), wn));                         // D4, D5
      if (has_prior) dsum = __fadd_rn(dsum, __fmul_rn(dl, wn));                           // D6
      if (k == 7 && mlive[par]) {
        bool mask;
        a.out_sdf[mq[par]] = finish_blend(sdf, dsum, mminw[par], a, g.vs, &mask);
        if (a.out_mask) a.out_mask[mq[par]] = mask ? 1 : 0;
      }
    };
    HelperSeq seq;
    int par = 0;
    for (int64_t tile = (int64_t)blockIdx.x * kNC + c.chain; tile < n_tiles; tile += (int64_t)gridDim.x * kNC, par ^= 1) {
      const int64_t q = tile * 128 + c.row;
      const bool live = q < a.n_queries;
      float cq[3] = {0.f, 0.f, 0.f};
      if (live) query_coords(m, a, q, cq);
      // ---- once per query: everything that depends on one axis only ---------------------------------
      int32_t tab[3][2];                // voxel index * table stride of this axis, or INT_MIN when outside the grid
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
          Q.w_ls[d][s][hid] = pack_f16x2(l, sn);
          Q.w_c1[d][s][hid] = pack_f16x2(cs, 1.0f);
          Q.t[d][s][hid] = __fsub_rn(1.f, fabsf(l));
          const int iv = (int)nbv[s];
          tab[d][s] = (live && iv >= 0 && iv < g.n[d]) ? iv * tstride[d] : kOut;
          int32_t ts = kOut;
          if (has_prior) {                                                                 // grid_sample(nearest), D6
            float t = __fdiv_rn(nbv[s], a.nm1[d]);
            t = __fmul_rn(t, 2.f);
            t = __fsub_rn(t, 1.f);
            t = __fadd_rn(t, 1.f);
            t = __fmul_rn(t, 0.5f);
            t = __fmul_rn(t, a.tm1[d]);
            const float rr = nearbyintf(t);
            if (rr >= 0.f && rr < (float)a.tsdf_dims[d]) ts = (int)rr * pstride[d];
          }
          Q.ts[d][s][hid] = ts;
        }
      }
      float wsum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float w = weight_of(k);
        wsum = k == 0 ? w : __fadd_rn(wsum, w);                                            // D2 normaliser
      }
      // ---- 8 independent table lookups in flight (_query_tensor, D3) ---------------------------------
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int32_t tx = tab[0][corner_sx(k)], ty = tab[1][corner_sy(k)], tz = tab[2][corner_sz(k)];
        int32_t sl = kEmpty;
        if (tx != kOut && ty != kOut && tz != kOut) sl = __ldg(m.table + ((int64_t)tx + ty + tz));
        Q.slot[k][hid] = sl;
      }
      uint4 f_cur, f_nxt = make_uint4(0, 0, 0, 0);
      float w_cur, w_nxt = 0.f;
      gather(0, f_cur, w_cur);
      float minw = 3.0e38f;
#pragma unroll 1
      for (int k = 0; k < 8; ++k) {
        if (k < 7) gather(k + 1, f_nxt, w_nxt);           // lands while this corner's row waits for its turn
        minw = fminf(minw, w_cur);                                                         // D3
        const float wn = __fdiv_rn(weight_of(k), wsum);                                    // D2
        const float dl = has_prior ? prior_of(k) : 0.f;
        seq.emit(
            c, k | (par << 3),
            [&]() {
              const int sx = corner_sx(k), sy = corner_sy(k), sz = corner_sz(k);
              const uint32_t in[16] = {f_cur.x, f_cur.y, f_cur.z, f_cur.w,
                                       Q.w_ls[0][sx][hid], Q.w_c1[0][sx][hid], Q.w_ls[1][sy][hid], Q.w_c1[1][sy][hid],
                                       Q.w_ls[2][sz][hid], Q.w_c1[2][sz][hid], kOnes, kOnes, kOnes, kOnes, kOnes, kOnes};
              h_stage<16>(c, in);
            },
            consume);
        // this corner's blend weight joins the FIFO behind the (at most one) corner still in flight
        if (fcount == 0) { fw0 = wn; fd0 = dl; } else { fw1 = wn; fd1 = dl; }
        ++fcount;
        f_cur = f_nxt;
        w_cur = w_nxt;
      }
      mq[par] = q;
      mminw[par] = minw;
      mlive[par] = live;
    }
    seq.finish(c, consume);
  }
  ws_teardown(S.sh);
}

// ---- factored decode of the meshlize sample blocks: G[V][l] table (see bnv_tc.cu) ----------------------
// rows = (exported voxel V, offset l of the 27): items = 128-row tiles; the helper builds each row from the voxel's
// packed features and the three per-axis word pairs of l = -0.5, 0, +0.5, and stores the MLP outputs into G
__global__ void __launch_bounds__(kThreads, 1) gtable_ws_kernel(const uint4* __restrict__ packed, int64_t n_rows,
                                                                 const uint8_t* __restrict__ gW, int w_bytes,
                                                                 float* __restrict__ G) {
  extern __shared__ __align__(128) uint8_t smem[];
  SmemBase& S = *reinterpret_cast<SmemBase*>(smem);
  Role c = ws_setup(S.sh, weights_smem(smem), gW, w_bytes);
  if (!c.helper) {
    e_run<16>(c);
  } else {
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
    auto consume = [&](int tile) {
      float y[1];
      h_read_out<1>(c, y);
      const int64_t row = (int64_t)tile * 128 + c.row;
      if (row < total) G[row] = y[0];
    };
    HelperSeq seq;
    for (int64_t tile = (int64_t)blockIdx.x * kNC + c.chain; tile < n_tiles; tile += (int64_t)gridDim.x * kNC) {
      const int64_t row = tile * 128 + c.row;
      const int64_t v = row / 27;
      const int li = (int)(row - v * 27);
      uint4 f = make_uint4(0, 0, 0, 0);
      if (row < total && v < n_rows) f = __ldg(packed + v);
      const int dx = li / 9, dy = (li / 3) % 3, dz = li % 3;
      seq.emit(
          c, (int)tile,
          [&]() {
            // dynamic index into 3-element register arrays -> selects
            const uint32_t lx = dx == 0 ? w_ls[0] : dx == 1 ? w_ls[1] : w_ls[2], cx = dx == 0 ? w_c1[0] : dx == 1 ? w_c1[1] : w_c1[2];
            const uint32_t ly = dy == 0 ? w_ls[0] : dy == 1 ? w_ls[1] : w_ls[2], cy = dy == 0 ? w_c1[0] : dy == 1 ? w_c1[1] : w_c1[2];
            const uint32_t lz = dz == 0 ? w_ls[0] : dz == 1 ? w_ls[1] : w_ls[2], cz = dz == 0 ? w_c1[0] : dz == 1 ? w_c1[1] : w_c1[2];
            const uint32_t in[16] = {f.x, f.y, f.z, f.w, lx, cx, ly, cy, lz, cz, kOnes, kOnes, kOnes, kOnes, kOnes, kOnes};
            h_stage<16>(c, in);
          },
          consume);
    }
    seq.finish(c, consume);
  }
  ws_teardown(S.sh);
}

// ---- plain forward (tcnnPointNetEncoder.forward / tcnnNeRFModel.geo_forward): items = 128-row tiles --------------
template <int NIN, int INW>
__device__ __forceinline__ void load_row(const float* __restrict__ x, int64_t i, int64_t n, uint32_t (&in)[INW]) {
  float xi[2 * INW];
#pragma unroll
  for (int k = 0; k < 2 * INW; ++k) {
    const int src = NIN == 17 ? dec_perm(k) : enc_perm(k);       // column order of the packed W0
    xi[k] = (src < NIN && i < n) ? __ldg(x + i * NIN + src) : 1.0f;
  }
#pragma unroll
  for (int k = 0; k < INW; ++k) in[k] = pack_f16x2(xi[2 * k], xi[2 * k + 1]);
}

template <int NIN, int INW, int NOUT>
__global__ void __launch_bounds__(kThreads, 1) mlp_forward_ws_kernel(const uint8_t* __restrict__ gW, int w_bytes,
                                                                      const float* __restrict__ x, int64_t n,
                                                                      float* __restrict__ y) {
  extern __shared__ __align__(128) uint8_t smem[];
  SmemBase& S = *reinterpret_cast<SmemBase*>(smem);
  Role c = ws_setup(S.sh, weights_smem(smem), gW, w_bytes);
  if (!c.helper) {
    e_run<INW>(c);
  } else {
    const int64_t n_tiles = (n + 127) / 128;
    auto consume = [&](int tile) {
      float out[NOUT];
      h_read_out<NOUT>(c, out);
      const int64_t row = (int64_t)tile * 128 + c.row;
      if (row < n) {
#pragma unroll
        for (int o = 0; o < NOUT; ++o) y[row * NOUT + o] = out[o];
      }
    };
    HelperSeq seq;
    for (int64_t tile = (int64_t)blockIdx.x * kNC + c.chain; tile < n_tiles; tile += (int64_t)gridDim.x * kNC) {
      uint32_t in[INW];
      load_row<NIN, INW>(x, tile * 128 + c.row, n, in);
      seq.emit(c, (int)tile, [&]() { h_stage<INW>(c, in); }, consume);
    }
    seq.finish(c, consume);
  }
  ws_teardown(S.sh);
}

}  // namespace wsk
}  // namespace bnv

// ---- host side ------------------------------------------------------------------------------------------
using namespace bnv::wsk;

int bnv_internal_encode_ws(bnv_map_t* map, int64_t max_records, const bnv_mlp_t* enc, cudaStream_t s) {
  const size_t smem = weights_off() + weight_image(enc->in_pad).bytes;
  BNV_CUDA(cudaFuncSetAttribute(encode_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = grid_for(((max_records + 127) / 128) * 8);        // units = (tile, corner)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // see grid_dependency_wait (bnv_frame.cuh)
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BNV_CUDA(cudaLaunchKernelEx(&cfg, encode_ws_kernel, map->d, (const uint8_t*)enc->w16, (int)enc->w16_bytes));
  BNV_LAUNCH_CHECK("encode_ws_kernel");
  return BNV_OK;
}

int bnv_internal_decode_ws(bnv_map_t* map, const bnv::DecArgs& a, const bnv_mlp_t* dec, cudaStream_t s) {
  const size_t smem = ((weights_off() + weight_image(dec->in_pad).bytes + 127) / 128) * 128 + sizeof(DecState);
  BNV_CUDA(cudaFuncSetAttribute(decode_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  decode_ws_kernel<<<grid_for((a.n_queries + 127) / 128), kThreads, smem, s>>>(map->d, a, (const uint4*)map->dec_pack,
                                                                              (const uint8_t*)dec->w16, (int)dec->w16_bytes);
  BNV_LAUNCH_CHECK("decode_ws_kernel");
  return BNV_OK;
}

int bnv_internal_gtable_ws(bnv_map_t* map, int64_t n_rows, const bnv_mlp_t* dec, cudaStream_t s) {
  const size_t smem = weights_off() + weight_image(dec->in_pad).bytes;
  BNV_CUDA(cudaFuncSetAttribute(gtable_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = ((n_rows + 1) * 27 + 127) / 128;
  gtable_ws_kernel<<<grid_for(tiles), kThreads, smem, s>>>((const uint4*)map->dec_pack, n_rows, (const uint8_t*)dec->w16,
                                                          (int)dec->w16_bytes, (float*)map->gtable);
  BNV_LAUNCH_CHECK("gtable_ws_kernel");
  return BNV_OK;
}


int bnv_internal_mlp_forward_ws(const bnv_mlp_t* mlp, const float* x, int64_t n, float* y, cudaStream_t s) {
  const size_t smem = weights_off() + weight_image(mlp->in_pad).bytes;
  const int grid = grid_for((n + 127) / 128);
  if (mlp->n_in == 6) {
    BNV_CUDA(cudaFuncSetAttribute(mlp_forward_ws_kernel<6, 8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_forward_ws_kernel<6, 8, 8><<<grid, kThreads, smem, s>>>((const uint8_t*)mlp->w16, (int)mlp->w16_bytes, x, n, y);
  } else {
    BNV_CUDA(cudaFuncSetAttribute(mlp_forward_ws_kernel<17, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_forward_ws_kernel<17, 16, 1><<<grid, kThreads, smem, s>>>((const uint8_t*)mlp->w16, (int)mlp->w16_bytes, x, n, y);
  }
  BNV_LAUNCH_CHECK("mlp_forward_ws_kernel");
  return BNV_OK;
}
