// Warp-specialised form of the tensor-core MLP chain (bnv_tc.cuh): every chain (one 128-column TMEM slot, 128 rows in
// flight) is served by TWO warpgroups.
//
//   epilogue warpgroup (4 warps, thread r <-> TMEM lane r): nothing but the chain itself --
//       wait D -> tcgen05.ld -> ReLU + cvt -> tcgen05.st A -> issue the next layer's UMMAs, three times per item,
//       then the output layer of this item together with the first layer of the next one;
//   helper warpgroup (4 warps, same lane quarters): everything around the chain -- gathers and table lookups, the
//       input row of the NEXT item into A_in, the output of the PREVIOUS item out of D_out and its blend / scatter.
//
// In the single-role kernels of round 1 that surrounding work ran in the "shadows" of the chain's own MMA round
// trips; it is ~2/3 of the instructions a row thread issues per item and stretched a layer round from the 577 cycles
// of the bare chain (89 % tensor-pipe utilisation, profiles/r1c_umma_microbench2.txt) to 1 000-1 150 cycles
// (47-51 %).  Here it runs on its own warps, concurrently, and the epilogue warpgroup's round is the bare one.
//
// Hand-offs per chain (mbarriers in shared memory):
//   bar_d  (tcgen05.commit)     hidden-layer accumulator complete                     -> epilogue warps
//   bar_o  (tcgen05.commit)     phase p: L0 of item p complete (A_in is free again) and, for p > 0, the output layer of
//                               item p - 1 complete (D_out holds its result)          -> helper warps
//   bar_h  (4 warp arrivals)    phase p: output of item p - 2 has been read out of D_out, the input row of item p is in
//                               A_in; the control word says whether item p exists      -> epilogue warps
// TMEM slot of a chain (128 columns) as before: D [0,64) | A_h [64,96) | A_in [96,112) | D_out [112,128).
#pragma once
#include "bnv_tc.cuh"

namespace bnv {
namespace ws {

using namespace bnv::tc;

constexpr int kNC = 4;                    // chains per CTA
constexpr int kThreads = kNC * 256;       // per chain: 4 epilogue warps + 4 helper warps

struct WsShared {
  uint64_t bar_d[kNC];
  uint64_t bar_o[kNC];
  uint64_t bar_h[kNC];
  uint64_t bar_w;                         // weight image arrived (cp.async.bulk complete_tx)
  uint32_t tmem_base;
  volatile int32_t next[kNC];             // control word of bar_h's current phase: 1 = the next item is staged, 0 = no more
  uint32_t live[kNC][2][4];               // tile shard: per-helper-warp corner masks (double-buffered)
};

struct Role {
  bool helper;                            // warp-uniform
  int chain, row;                         // chain of this warp, row of this thread in the 128-row tile
  uint32_t t_d, t_a, t_in, t_o, d_slot;
  uint64_t *bar_d, *bar_o, *bar_h;
  uint32_t par_d, par_o, par_h;
  uint32_t w_saddr;
  int bar_e, bar_hw;                      // named barriers of the epilogue / helper warpgroup (128 threads each)
  bool issuer_warp;
  volatile int32_t* next;
};

__device__ __forceinline__ Role ws_setup(WsShared& sh, uint8_t* s_weights, const uint8_t* __restrict__ g_weights, int w_bytes) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int g = 0; g < kNC; ++g) {
      mbar_init(&sh.bar_d[g], 1);
      mbar_init(&sh.bar_o[g], 1);
      mbar_init(&sh.bar_h[g], 4);
      sh.next[g] = 0;
    }
    mbar_init(&sh.bar_w, 1);
    mbar_fence_init();
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sh.bar_w)), "r"((uint32_t)w_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(s_weights)), "l"(g_weights), "r"((uint32_t)w_bytes), "r"(smem_u32(&sh.bar_w))
                 : "memory");
  }
  if (warp == 0) tmem_alloc(&sh.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  mbar_wait(&sh.bar_w, 0);
  Role r;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  r.helper = warp_u >= 4 * kNC;
  r.chain = (warp_u & (4 * kNC - 1)) >> 2;
  r.row = (warp_u & 3) * 32 + (tid & 31);
  r.d_slot = __shfl_sync(0xffffffffu, sh.tmem_base, 0) + r.chain * kSlotCols;
  r.t_d = r.d_slot + ((uint32_t)((warp_u & 3) * 32) << 16);     // a warp reaches TMEM lanes 32 (warp % 4) .. + 31
  r.t_a = r.t_d + kACol;
  r.t_in = r.t_d + kInCol;
  r.t_o = r.t_d + kOutCol;
  r.bar_d = &sh.bar_d[r.chain];
  r.bar_o = &sh.bar_o[r.chain];
  r.bar_h = &sh.bar_h[r.chain];
  r.par_d = r.par_o = r.par_h = 0;
  r.w_saddr = smem_u32(s_weights);
  r.bar_e = 1 + r.chain;
  r.bar_hw = 1 + kNC + r.chain;
  r.issuer_warp = !r.helper && (warp_u & 3) == (r.chain & 3);   // issuers of the four chains sit in different SM sub-partitions
  r.next = &sh.next[r.chain];
  return r;
}

__device__ __forceinline__ void ws_teardown(WsShared& sh) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc(sh.tmem_base, 512);
}

template <int N>
__device__ __forceinline__ void ws_umma(const Role& c, uint32_t b0, int d_col, int a_col, int kk) {
  constexpr uint32_t lbo = (uint32_t)(N / 8) * 128u;
  umma_ts_f16(c.d_slot + d_col, c.d_slot + a_col + kk * 8, smem_desc_kmajor(b0 + kk * 2 * lbo, lbo, 128), idesc_f16_m128(N),
              kk > 0 ? 1u : 0u);
}

// ---- epilogue warpgroup ------------------------------------------------------------------------------------------
__device__ __forceinline__ void e_wait_d(Role& c) {
  mbar_wait(c.bar_d, c.par_d);
  c.par_d ^= 1;
  tc_fence_after();
}
// helper's hand-off of the current phase: returns whether another item has been staged
__device__ __forceinline__ bool e_wait_h(Role& c) {
  mbar_wait(c.bar_h, c.par_h);
  c.par_h ^= 1;
  return *c.next != 0;
}
// 64 fp32 accumulator columns -> ReLU -> 32 packed fp16x2 columns of A_h, in four 16-column chunks so that at most
// 48 registers are live (1 024 threads per CTA leave 64 per thread); chunk i + 1 is in flight while chunk i converts
__device__ __forceinline__ void e_epilogue(Role& c) {
  uint32_t x[16], y[16], a[16];
  tmem_ld16(c.t_d, x);
  tmem_wait_ld();
  tmem_ld16(c.t_d + 16, y);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = pack_relu_f16x2(x[2 * i], x[2 * i + 1]);
  tmem_wait_ld();
  tmem_ld16(c.t_d + 32, x);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[8 + i] = pack_relu_f16x2(y[2 * i], y[2 * i + 1]);
  tmem_st16(c.t_a, a);
  tmem_wait_ld();
  tmem_ld16(c.t_d + 48, y);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = pack_relu_f16x2(x[2 * i], x[2 * i + 1]);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 8; ++i) a[8 + i] = pack_relu_f16x2(y[2 * i], y[2 * i + 1]);
  tmem_st16(c.t_a + 16, a);
}
template <class F>
__device__ __forceinline__ void e_sync_issue(Role& c, F&& f) {
  tmem_wait_st();
  tc_fence_before();
  wg_sync(c.bar_e);
  if (c.issuer_warp) {
    tc_fence_after();
    if (elect_one()) f();
  }
}

// The whole life of an epilogue warpgroup: run the chain for as many items as the helper stages.  INW = packed input
// words (8: in_pad 16, 16: in_pad 32).
template <int INW>
__device__ __forceinline__ void e_run(Role& c) {
  constexpr int off1 = 2 * INW * 64 * 2, off2 = off1 + 64 * 64 * 2, off3 = off2 + 64 * 64 * 2;
  if (!e_wait_h(c)) return;                              // phase 0: is there a first item at all ?
  e_sync_issue(c, [&]() {
#pragma unroll
    for (int kk = 0; kk < INW / 8; ++kk) ws_umma<64>(c, c.w_saddr, 0, kInCol, kk);
    umma_commit(c.bar_d);
    umma_commit(c.bar_o);                                // phase 0 of bar_o: A_in is free once L0 of item 0 is done
  });
  while (true) {
    e_wait_d(c);
    e_epilogue(c);
    e_sync_issue(c, [&]() {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) ws_umma<64>(c, c.w_saddr + off1, 0, kACol, kk);
      umma_commit(c.bar_d);
    });
    e_wait_d(c);
    e_epilogue(c);
    e_sync_issue(c, [&]() {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) ws_umma<64>(c, c.w_saddr + off2, 0, kACol, kk);
      umma_commit(c.bar_d);
    });
    e_wait_d(c);
    e_epilogue(c);
    // the helper has read the previous item's output out of D_out and staged the next item's input (or said "none")
    const bool has_next = e_wait_h(c);
    e_sync_issue(c, [&]() {
      if (has_next) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          ws_umma<16>(c, c.w_saddr + off3, kOutCol, kACol, kk);
          if (kk < INW / 8) ws_umma<64>(c, c.w_saddr, 0, kInCol, kk);
        }
        umma_commit(c.bar_o);
        umma_commit(c.bar_d);
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ws_umma<16>(c, c.w_saddr + off3, kOutCol, kACol, kk);
        umma_commit(c.bar_o);
      }
    });
    if (!has_next) break;
  }
}

// ---- helper warpgroup ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void h_wait_o(Role& c) {
  mbar_wait(c.bar_o, c.par_o);
  c.par_o ^= 1;
  tc_fence_after();
}
template <int INW>
__device__ __forceinline__ void h_stage(Role& c, const uint32_t (&in)[INW]) {
  static_assert(INW == 8 || INW == 16, "in_pad must be 16 or 32");
  if constexpr (INW == 8) tmem_st8(c.t_in, in); else tmem_st16(c.t_in, in);
}
template <int NOUT>
__device__ __forceinline__ void h_read_out(Role& c, float (&out)[NOUT]) {
  static_assert(NOUT == 8 || NOUT == 1, "n_out must be 8 or 1");
  if constexpr (NOUT == 8) {
    uint32_t r[8];
    tmem_ld8(c.t_o, r);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = __uint_as_float(r[j]);
  } else {
    uint32_t r;
    tmem_ld1(c.t_o, r);
    tmem_wait_ld();
    out[0] = __uint_as_float(r);
  }
}
// end of a helper phase: this warp's TMEM loads / stores are done; one arrival per warp
__device__ __forceinline__ void h_publish(Role& c, bool has_next) {
  tmem_wait_st();
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    *c.next = has_next ? 1 : 0;           // all four warps write the same value
    mbar_arrive(c.bar_h);
  }
}

// Sequencer of a helper warpgroup.  Call `emit(stage, consume)` once per item in order -- `stage()` writes the item's
// input row (h_stage), `consume(tag)` handles an output with the tag given when its item was emitted -- then `finish`.
// Item j's input is staged once L0 of item j - 1 has completed; its output is consumed two emits later.
struct HelperSeq {
  int n = 0;                // items emitted
  int tag_prev = 0, tag_cur = 0;
  template <class Stage, class Consume>
  __device__ __forceinline__ void emit(Role& c, int tag, Stage&& stage, Consume&& consume) {
    if (n > 0) {
      h_wait_o(c);                         // phase n - 1: L0 of item n - 1 done, output layer of item n - 2 done
      if (n >= 2) consume(tag_prev);
    }
    stage();
    h_publish(c, true);                    // bar_h phase n
    tag_prev = tag_cur;
    tag_cur = tag;
    ++n;
  }
  template <class Consume>
  __device__ __forceinline__ void finish(Role& c, Consume&& consume) {
    if (n == 0) {
      h_publish(c, false);
      return;
    }
    h_wait_o(c);                           // phase n - 1
    if (n >= 2) consume(tag_prev);
    h_publish(c, false);                   // bar_h phase n: no further item
    h_wait_o(c);                           // phase n: output layer of the last item
    consume(tag_cur);
  }
};

}  // namespace ws
}  // namespace bnv
