// tcgen05 kernels on the software-pipelined chain (bnv_tc.cuh): plain MLP forward, encode
// (point records -> 8 corner rows -> encoder MLP -> scatter), fused decode (8-corner gather -> decoder MLP ->
// trilinear blend + prior) and the G table of the factored meshlize decode.
//
//  * every item (one 128-row tile through the MLP) costs three exposed MMA round trips: the output layer runs on
//    the tensor core and is issued together with the next item's first layer, its result is consumed in the
//    shadow of the next item;
//  * the encode kernel runs over the point records the frame prepass compacted (bnv_encode.cu) and splits them at
//    (tile, corner) granularity: every chain gets the same number of corner rows (+-1);
//  * in the tile shard, the corners nobody in the warpgroup owns are dropped from the item list up front.
#include <cuda_fp16.h>
#include <limits.h>
#include <stdlib.h>

#include "bnv_common.cuh"
#include "bnv_decode_common.cuh"
#include "bnv_frame.cuh"
#include "bnv_tc.cuh"

using namespace bnv;
using namespace bnv::tc;

namespace bnv {
namespace tcc {

constexpr int kNWG = 4;                       // row warpgroups (= chains = TMEM slots) per CTA
constexpr int kThreads = kNWG * 128;
constexpr uint32_t kOnes = 0x3C003C00u;       // fp16x2 {1.0, 1.0}: tcnn pads the input with ones

struct alignas(16) Smem {         // plain forward, decode, G table: the chain's barriers only
  TcShared<kNWG> sh;
};
struct alignas(16) SmemEnc {      // encode
  TcShared<kNWG> sh;
  int32_t slot[8][kThreads];      // dense scratch row of (this thread's point, corner k); thread-private, [k][tid]
  float4 rec[2][2][kThreads];     // point records of this / the next tile (cp.async double buffer), [buf][half][tid]
};

template <class S>
__device__ __forceinline__ uint8_t* weights_smem(uint8_t* smem) { return smem + ((sizeof(S) + 127) / 128) * 128; }
template <class S>
static size_t smem_bytes(int in_pad) { return ((sizeof(S) + 127) / 128) * 128 + weight_image(in_pad).bytes; }

static int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
static int grid_for(int64_t n_items) {
  const int sms = sm_count();
  const int64_t need = (n_items + kNWG - 1) / kNWG;
  return (int)(need < sms ? (need < 1 ? 1 : need) : sms);
}

// ---- plain forward: items = 128-row tiles --------------------------------------------------------
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
__global__ void __launch_bounds__(kThreads, 1) mlp_forward_tc_kernel(const uint8_t* __restrict__ gW, int w_bytes,
                                                                      const float* __restrict__ x, int64_t n,
                                                                      float* __restrict__ y) {
  extern __shared__ __align__(128) uint8_t smem[];
  Smem& S = *reinterpret_cast<Smem*>(smem);
  RowChain c = tc_setup<kNWG>(S.sh, weights_smem<Smem>(smem), gW, w_bytes);
  const int wg = threadIdx.x >> 7, r = threadIdx.x & 127;
  const int64_t n_tiles = (n + 127) / 128;
  const int64_t stride = (int64_t)gridDim.x * kNWG;
  int64_t tile = (int64_t)blockIdx.x * kNWG + wg;
  int64_t pend = -1;                                  // row whose output is still in D_out
  auto drain = [&]() {
    if (pend >= 0) {
      float out[NOUT];
      chain_output<NOUT>(c, out);
      if (pend < n) {
#pragma unroll
        for (int o = 0; o < NOUT; ++o) y[pend * NOUT + o] = out[o];
      }
    }
    pend = -1;
  };
  if (tile < n_tiles) {
    uint32_t in[INW];
    load_row<NIN, INW>(x, tile * 128 + r, n, in);
    chain_stage<INW>(c, in);
    chain_begin<INW>(c);
    for (; tile < n_tiles; tile += stride) {
      const bool has_next = tile + stride < n_tiles;
      chain_hidden<INW>(
          c,
          [&]() {
            drain();
            if (has_next) load_row<NIN, INW>(x, (tile + stride) * 128 + r, n, in);
          },
          [&]() {
            if (has_next) chain_stage<INW>(c, in);
          });
      chain_finish<INW>(c, has_next);
      pend = tile * 128 + r;
    }
    drain();
  }
  tc_teardown<kNWG>(S.sh);
}

// ---- cp.async (LDGSTS): global -> shared memory without staging registers -------------------------------
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// 16 / 4 bytes, zero-filled when !pred (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool pred) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(pred ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(void* smem_dst, const void* gsrc, bool pred) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(pred ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }


// ---- fused encode -----------------------------------------------------------------------------------
// Work unit = (128-point tile, corner).  Chain j of J processes the contiguous unit range
// [U j / J, U (j + 1) / J): whole tiles in the middle, partial tiles (a corner sub-range) at both ends, so
// all chains carry the same number of MLP rounds.  Thread r of the warpgroup owns point tile * 128 + r.
__device__ __forceinline__ void enc_input(int k, const float (&cc)[3], const float (&fl)[3], const float (&ce)[3],
                                          float vs, float inv_vs, uint32_t nrm01, uint32_t nrm2o, uint32_t (&in)[8]) {
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

__global__ void __launch_bounds__(kThreads, 1) encode_chain_kernel(MapDev m, const uint8_t* __restrict__ gW, int w_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  SmemEnc& S = *reinterpret_cast<SmemEnc*>(smem);
  RowChain c = tc_setup<kNWG>(S.sh, weights_smem<SmemEnc>(smem), gW, w_bytes);
  const int wg = threadIdx.x >> 7, r = threadIdx.x & 127, warp_in_wg = r >> 5, tid = threadIdx.x;
  const GeomDev& g = m.g;
  grid_dependency_wait();                               // the prepass (setup above overlapped its tail)
  const int64_t n_rec = m.ctr[4];                       // point records written by frame_prepass_kernel
  const int64_t n_units = ((n_rec + 127) / 128) * 8;
  const int64_t chain = (int64_t)blockIdx.x * kNWG + wg, n_chains = (int64_t)gridDim.x * kNWG;
  const int64_t u_end = n_units * (chain + 1) / n_chains;
  int32_t pend_row = -1;
  bool pending = false;                                // an output layer is in flight / unread in D_out
  int flip = 0, buf = 0;
  auto drain = [&]() {
    if (pending) {
      float y[8];
      chain_output<8>(c, y);
      if (pend_row >= 0) add_row_f32(m, pend_row, y);
      pending = false;
    }
  };
  // point record of tile `tile` -> shared memory (32 bytes per thread, no registers held across the chain)
  auto fetch_record = [&](int64_t tile, int b) {
    const int64_t idx = tile * 128 + r;
    const bool have = idx < n_rec;
    const float4* r4 = reinterpret_cast<const float4*>(m.prec + (size_t)(have ? idx : 0) * 8);
    cp_async16_zfill(&S.rec[b][0][tid], r4, have);
    cp_async16_zfill(&S.rec[b][1][tid], r4 + 1, have);
    cp_async_commit();
  };
  BNV_PROF_MARK(p_life);
  int64_t u = n_units * chain / n_chains;
  if (u < u_end) fetch_record(u >> 3, buf);
  while (u < u_end) {
    BNV_PROF_MARK(p_pre);
    const int64_t tile = u >> 3;
    const int k0 = (int)(u & 7);
    const int k1 = (int)(u_end - u < (int64_t)(8 - k0) ? k0 + (u_end - u) : 8);
    u += k1 - k0;
    const int64_t idx = tile * 128 + r;
    cp_async_wait_all();
    const float4 ra = S.rec[buf][0][tid], rb = S.rec[buf][1][tid];     // zeros past the last record
    if (u < u_end) fetch_record(tile + 1, buf ^ 1);                     // lands while this tile's corners run
    buf ^= 1;
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
    // dense scratch rows of the owned corners: 8 independent table reads (high word of ftable[flat]) whose results
    // are parked in shared memory in the shadow of the tile's first MLP round, where they are first needed
    int32_t rows[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      rows[k] = -1;
      if ((own >> k) & 1u) {
        float nb[3];
        corner_of(k, fl, ce, nb);
        rows[k] = scratch_row_of(m, (int)nb[0] * g.nyz + (int)nb[1] * g.n[2] + (int)nb[2]);   // rule A5
      }
    }
    // corners to run: all of [k0, k1) on one GPU; in the tile shard only those somebody here owns
    uint32_t live = range;
    if (g.world > 1) {
      const uint32_t wown = __reduce_or_sync(0xffffffffu, own);
      if ((r & 31) == 0) S.sh.live[wg][flip][warp_in_wg] = wown;
      wg_sync(c.bar_id);
      live = S.sh.live[wg][flip][0] | S.sh.live[wg][flip][1] | S.sh.live[wg][flip][2] | S.sh.live[wg][flip][3];
      flip ^= 1;
    }
    if (live == 0) continue;              // nobody here owns anything of this tile: `own` is 0 for every thread
    // the chain is idle here (the previous tile's last corner was finished without a next item; its
    // output, if still unread, is drained in the shadow of this tile's first corner)
    {
      uint32_t in[8];
      enc_input(__ffs(live) - 1, cc, fl, ce, g.vs, g.inv_vs, nrm01, nrm2o, in);
      chain_stage<8>(c, in);
      chain_begin<8>(c);
    }
    BNV_PROF_ADD(9, p_pre);
    bool first = true;
#pragma unroll 1
    for (uint32_t rem = live; rem;) {
      const int k = __ffs(rem) - 1;
      rem &= rem - 1;
      const bool has_next = rem != 0;
      chain_hidden<8>(
          c,
          [&]() {
            drain();                      // reads the PREVIOUS tile's row from pend_row (a register): safe to overwrite
            if (first) {
#pragma unroll
              for (int j = 0; j < 8; ++j) S.slot[j][tid] = rows[j];
            }
          },
          [&]() {
            if (has_next) {
              uint32_t in[8];
              enc_input(__ffs(rem) - 1, cc, fl, ce, g.vs, g.inv_vs, nrm01, nrm2o, in);
              chain_stage<8>(c, in);
            }
          });
      first = false;
      chain_finish<8>(c, has_next);
      pending = true;
      pend_row = S.slot[k][tid];
    }
  }
  drain();
  cp_async_wait_all();
  BNV_PROF_ADD(10, p_life);
  tc_teardown<kNWG>(S.sh);
}

// ---- fused decode -----------------------------------------------------------------------------------
// (A variant that prefetched the table entries and feature rows ONE TILE ahead with cp.async into shared memory was
// measured in round 2: 5.6 instead of 6.1 Gqueries/s -- the chain is bound by the instructions each warp issues per
// corner, not by the gather latency, and the extra shared-memory traffic costs more than the exposed latency saved.)
// Per-query state lives in shared memory ([word][thread], conflict-free, thread-private): everything that
// depends on one axis only exists in a floor (s = 0) and a ceil (s = 1) flavour computed once per query,
// a corner row is then 4 gathered words + 6 selected words + 6 constants.  Keeping it out of the register
// file leaves room for the 96 transient registers of the hidden-layer epilogue (no spills) and lets the
// 8-corner loop stay rolled (8x less code, no instruction-cache misses).
struct DecState {
  uint32_t w_ls[3][2][kThreads];   // fp16x2 {l, sin l}
  uint32_t w_c1[3][2][kThreads];   // fp16x2 {cos l, 1}
  float t[3][2][kThreads];         // 1 - |l|
  int32_t ts[3][2][kThreads];      // TSDF-prior index * its stride, or INT_MIN when outside (nearest lookup)
  int32_t slot[8][kThreads];       // table lookup of corner k (_query_tensor)
};

__global__ void __launch_bounds__(kThreads, 1) decode_tc_kernel(MapDev m, DecArgs a, const uint4* __restrict__ packed,
                                                                 const uint8_t* __restrict__ gW, int w_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  Smem& S = *reinterpret_cast<Smem*>(smem);
  RowChain c = tc_setup<kNWG>(S.sh, weights_smem<Smem>(smem), gW, w_bytes);
  DecState& Q = *reinterpret_cast<DecState*>(weights_smem<Smem>(smem) + ((w_bytes + 127) / 128) * 128);
  const int tid = threadIdx.x, wg = tid >> 7, r = tid & 127;
  const int64_t n_tiles = (a.n_queries + 127) / 128;
  const GeomDev& g = m.g;
  constexpr int32_t kOut = INT_MIN;
  const bool has_prior = a.tsdf != nullptr;
  auto weight_of = [&](int k) {                                                          // D2, corner k
    return __fmul_rn(__fmul_rn(Q.t[0][corner_sx(k)][tid], Q.t[1][corner_sy(k)][tid]), Q.t[2][corner_sz(k)][tid]);
  };
  auto prior_of = [&](int k) {                                                           // D6, corner k
    const int32_t px = Q.ts[0][corner_sx(k)][tid], py = Q.ts[1][corner_sy(k)][tid], pz = Q.ts[2][corner_sz(k)][tid];
    return (px != kOut && py != kOut && pz != kOut) ? __ldg(a.tsdf + ((int64_t)px + py + pz)) : 0.f;
  };
  auto gather = [&](int k, uint4& f, float& w) {                                         // D3
    f = make_uint4(0, 0, 0, 0);
    w = 0.f;
    const int32_t s = Q.slot[k][tid];
    if (s >= 0 && s < a.n_rows) {
      f = __ldg(packed + s);
      w = __ldg(a.weights_rows + s);
    }
  };
  auto stage_corner = [&](int k, const uint4& f) {
    const int sx = corner_sx(k), sy = corner_sy(k), sz = corner_sz(k);
    const uint32_t in[16] = {f.x, f.y, f.z, f.w,
                             Q.w_ls[0][sx][tid], Q.w_c1[0][sx][tid], Q.w_ls[1][sy][tid], Q.w_c1[1][sy][tid],
                             Q.w_ls[2][sz][tid], Q.w_c1[2][sz][tid], kOnes, kOnes, kOnes, kOnes, kOnes, kOnes};
    chain_stage<16>(c, in);
  };
  // the query whose last corner is still in D_out
  bool pending = false, p_live = false;
  int64_t p_q = 0;
  float p_sdf = 0.f, p_dsum = 0.f, p_minw = 0.f, p_wn = 0.f, p_dl = 0.f;
  auto drain = [&]() {
    if (pending) {
      float y[1];
      chain_output<1>(c, y);
      p_sdf = __fadd_rn(p_sdf, __fmul_rn(__fmul_rn(y[0], g.vs), p_wn));                   // D4, D5 (corner 7)
      if (has_prior) p_dsum = __fadd_rn(p_dsum, __fmul_rn(p_dl, p_wn));                  // D6
      if (p_live) {
        bool mask;
        a.out_sdf[p_q] = finish_blend(p_sdf, p_dsum, p_minw, a, g.vs, &mask);
        if (a.out_mask) a.out_mask[p_q] = mask ? 1 : 0;
      }
      pending = false;
    }
  };
  BNV_PROF_MARK(p_life);
  for (int64_t tile = (int64_t)blockIdx.x * kNWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * kNWG) {
    BNV_PROF_MARK(p_pre);
    const int64_t q = tile * 128 + r;
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
        Q.w_ls[d][s][tid] = pack_f16x2(l, sn);
        Q.w_c1[d][s][tid] = pack_f16x2(cs, 1.0f);
        Q.t[d][s][tid] = __fsub_rn(1.f, fabsf(l));
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
        Q.ts[d][s][tid] = ts;
      }
    }
    float wsum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float w = weight_of(k);
      wsum = k == 0 ? w : __fadd_rn(wsum, w);                                            // D2 normaliser
    }
    // ---- 8 independent table lookups in flight (_query_tensor, D3) ---------------------------------
    int32_t slot0 = kEmpty;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int32_t tx = tab[0][corner_sx(k)], ty = tab[1][corner_sy(k)], tz = tab[2][corner_sz(k)];
      int32_t sl = kEmpty;
      if (tx != kOut && ty != kOut && tz != kOut) sl = __ldg(m.table + ((int64_t)tx + ty + tz));
      Q.slot[k][tid] = sl;
      if (k == 0) slot0 = sl;
    }
    uint4 f_nxt = make_uint4(0, 0, 0, 0);
    float w_nxt = 0.f;
    if (slot0 >= 0 && slot0 < a.n_rows) {
      f_nxt = __ldg(packed + slot0);
      w_nxt = __ldg(a.weights_rows + slot0);
    }
    // the previous query's last corner has been in flight during all of the above
    drain();
    float minw = 3.0e38f, sdf = 0.f, dsum = 0.f;
    stage_corner(0, f_nxt);
    chain_begin<16>(c);
    BNV_PROF_ADD(9, p_pre);
    float w_cur = w_nxt;
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      minw = fminf(minw, w_cur);                                                         // D3
      chain_hidden<16>(
          c,
          [&]() {
            // shadow of the second layer: issue the next corner's gather, blend the previous corner
            if (k < 7) gather(k + 1, f_nxt, w_nxt);
            if (k > 0) {
              float y[1];
              chain_output<1>(c, y);                                                    // D7: all 8 rows are evaluated
              const float wn = __fdiv_rn(weight_of(k - 1), wsum);                        // D2
              sdf = __fadd_rn(sdf, __fmul_rn(__fmul_rn(y[0], g.vs), wn));                // D4, D5
              if (has_prior) dsum = __fadd_rn(dsum, __fmul_rn(prior_of(k - 1), wn));     // D6
            }
          },
          [&]() {
            // shadow of the third layer: the gather has landed -> stage the next corner's row
            if (k < 7) stage_corner(k + 1, f_nxt);
          });
      chain_finish<16>(c, k < 7);
      w_cur = w_nxt;
    }
    // corner 7 is in flight: finish this query at the top of the next tile (or after the loop)
    pending = true;
    p_live = live;
    p_q = q;
    p_sdf = sdf;
    p_dsum = dsum;
    p_minw = minw;
    p_wn = __fdiv_rn(weight_of(7), wsum);
    p_dl = has_prior ? prior_of(7) : 0.f;
  }
  drain();
  BNV_PROF_ADD(10, p_life);
  tc_teardown<kNWG>(S.sh);
}

// ---- factored decode of the meshlize sample blocks: G[V][l] table (see bnv_tc.cu) ----------------------
__global__ void __launch_bounds__(kThreads, 1) gtable_tc_kernel(const uint4* __restrict__ packed, int64_t n_rows,
                                                                 const uint8_t* __restrict__ gW, int w_bytes,
                                                                 float* __restrict__ G) {
  extern __shared__ __align__(128) uint8_t smem[];
  Smem& S = *reinterpret_cast<Smem*>(smem);
  RowChain c = tc_setup<kNWG>(S.sh, weights_smem<Smem>(smem), gW, w_bytes);
  const int wg = threadIdx.x >> 7, r = threadIdx.x & 127;
  const int64_t total = (n_rows + 1) * 27;                       // voxel n_rows = the miss voxel
  const int64_t n_tiles = (total + 127) / 128;
  const int64_t stride = (int64_t)gridDim.x * kNWG;
  uint32_t w_ls[3], w_c1[3];                                     // l = -0.5, 0, +0.5
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float l = 0.5f * (float)(d - 1);
    float sn, cs;
    __sincosf(l, &sn, &cs);
    w_ls[d] = pack_f16x2(l, sn);
    w_c1[d] = pack_f16x2(cs, 1.0f);
  }
  auto make_row = [&](int64_t row, uint32_t (&in)[16]) {
    const int64_t v = row / 27;
    const int li = (int)(row - v * 27);
    uint4 f = make_uint4(0, 0, 0, 0);
    if (row < total && v < n_rows) f = __ldg(packed + v);
    const int dx = li / 9, dy = (li / 3) % 3, dz = li % 3;
    // dynamic index into 3-element register arrays -> selects
    const uint32_t lx = dx == 0 ? w_ls[0] : dx == 1 ? w_ls[1] : w_ls[2], cx = dx == 0 ? w_c1[0] : dx == 1 ? w_c1[1] : w_c1[2];
    const uint32_t ly = dy == 0 ? w_ls[0] : dy == 1 ? w_ls[1] : w_ls[2], cy = dy == 0 ? w_c1[0] : dy == 1 ? w_c1[1] : w_c1[2];
    const uint32_t lz = dz == 0 ? w_ls[0] : dz == 1 ? w_ls[1] : w_ls[2], cz = dz == 0 ? w_c1[0] : dz == 1 ? w_c1[1] : w_c1[2];
    in[0] = f.x; in[1] = f.y; in[2] = f.z; in[3] = f.w;
    in[4] = lx; in[5] = cx; in[6] = ly; in[7] = cy; in[8] = lz; in[9] = cz;
#pragma unroll
    for (int j = 10; j < 16; ++j) in[j] = kOnes;
  };
  int64_t tile = (int64_t)blockIdx.x * kNWG + wg;
  int64_t pend = -1;
  auto drain = [&]() {
    if (pend >= 0) {
      float y[1];
      chain_output<1>(c, y);
      if (pend < total) G[pend] = y[0];
    }
    pend = -1;
  };
  if (tile < n_tiles) {
    uint32_t in[16];
    make_row(tile * 128 + r, in);
    chain_stage<16>(c, in);
    chain_begin<16>(c);
    for (; tile < n_tiles; tile += stride) {
      const bool has_next = tile + stride < n_tiles;
      chain_hidden<16>(
          c,
          [&]() {
            drain();
            if (has_next) make_row((tile + stride) * 128 + r, in);
          },
          [&]() {
            if (has_next) chain_stage<16>(c, in);
          });
      chain_finish<16>(c, has_next);
      pend = tile * 128 + r;
    }
    drain();
  }
  tc_teardown<kNWG>(S.sh);
}

}  // namespace tcc
}  // namespace bnv

// ---- host side ------------------------------------------------------------------------------------------
using namespace bnv::tcc;

template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
  BNV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return BNV_OK;
}

int bnv_internal_mlp_forward_chain(const bnv_mlp_t* mlp, const float* x, int64_t n, float* y, cudaStream_t s) {
  const size_t smem = smem_bytes<Smem>(mlp->in_pad);
  const int grid = grid_for((n + 127) / 128);
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

int bnv_internal_encode_chain(bnv_map_t* map, int64_t max_records, const bnv_mlp_t* enc, cudaStream_t s) {
  // the record count lives on the device (ctr[4]); the grid is sized for the most the prepass can have written
  const size_t smem = smem_bytes<SmemEnc>(enc->in_pad);
  int rc = set_smem(encode_chain_kernel, smem);
  if (rc) return rc;
  const int grid = grid_for(((max_records + 127) / 128) * 8);        // units = (tile, corner)
  {
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
    BNV_CUDA(cudaLaunchKernelEx(&cfg, encode_chain_kernel, map->d, (const uint8_t*)enc->w16, (int)enc->w16_bytes));
  }
  BNV_LAUNCH_CHECK("encode_chain_kernel");
  return BNV_OK;
}

int bnv_internal_decode_chain(bnv_map_t* map, const bnv::DecArgs& a, const bnv_mlp_t* dec, cudaStream_t s) {
  const size_t smem = ((smem_bytes<Smem>(dec->in_pad) + 127) / 128) * 128 + sizeof(DecState);
  int rc = set_smem(decode_tc_kernel, smem);
  if (rc) return rc;
  decode_tc_kernel<<<grid_for((a.n_queries + 127) / 128), kThreads, smem, s>>>(
      map->d, a, (const uint4*)map->dec_pack, (const uint8_t*)dec->w16, (int)dec->w16_bytes);
  BNV_LAUNCH_CHECK("decode_tc_kernel");
  return BNV_OK;
}

int bnv_internal_gtable_chain(bnv_map_t* map, int64_t n_rows, const bnv_mlp_t* dec, cudaStream_t s) {
  const size_t smem = smem_bytes<Smem>(dec->in_pad);
  int rc = set_smem(gtable_tc_kernel, smem);
  if (rc) return rc;
  const int64_t tiles = ((n_rows + 1) * 27 + 127) / 128;
  gtable_tc_kernel<<<grid_for(tiles), kThreads, smem, s>>>((const uint4*)map->dec_pack, n_rows, (const uint8_t*)dec->w16,
                                                            (int)dec->w16_bytes, (float*)map->gtable);
  BNV_LAUNCH_CHECK("gtable_tc_kernel");
  return BNV_OK;
}

#if BNV_CHAIN_PROFILE
// profiling build only: read (and reset) the phase timers of this translation unit's chain kernels
extern "C" int bnv_debug_chain_profile(unsigned long long* out16, int reset) {
  BNV_CUDA(cudaDeviceSynchronize());
  BNV_CUDA(cudaMemcpyFromSymbol(out16, bnv::tc::g_chain_prof, 16 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[16] = {0};
    BNV_CUDA(cudaMemcpyToSymbol(bnv::tc::g_chain_prof, z, sizeof(z)));
  }
  return BNV_OK;
}
#endif
