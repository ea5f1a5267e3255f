// tcgen05 / TMEM / mbarrier building blocks for the tensor-core MLP chain (sm_100a only).
//
// The tiny tcnn-style MLPs (n_in -> 64 -> 64 -> 64 -> n_out, ReLU, no bias) are evaluated for
// 128-row tiles as four chained UMMAs whose A operand never leaves tensor memory:
//
//   registers --tcgen05.st--> TMEM A (fp16 [128 x K]) --tcgen05.mma (B = weights in smem)--> TMEM D
//   (fp32 [128 x 64]) --tcgen05.ld--> registers (ReLU + cvt.f16x2) --tcgen05.st--> TMEM A  ...
//
// Row r of a tile is TMEM lane r and is owned by thread r of a 128-thread "row warpgroup", so the
// activations need no shared-memory round trip and no swizzled layouts; only the weights (B, K-major,
// no swizzle) sit in shared memory, loaded once per persistent CTA.  One elected lane of one warp of
// each row warpgroup issues that warpgroup's UMMAs; four warpgroups (one 128-column TMEM slot each) keep
// the tensor pipe busy while the others run their epilogues.  See "The software-pipelined chain" below.
#pragma once
#include <cuda_fp16.h>

#include "bnv_common.cuh"

namespace bnv {
namespace tc {

constexpr int kRowsPerTile = 128;
constexpr int kSlotCols = 128;       // TMEM columns per chain (layout below)
constexpr int kACol = 64;

// ---- shared-memory weight image (built on the host, bnv_tc.cu) ---------------------------------
// Per layer: B [N x K] fp16, K-major, SWIZZLE_NONE canonical layout:
//   byte(n, k) = (k / 8) * LBO + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2,   LBO = (N / 8) * 128
// i.e. 8x8 core matrices of 128 contiguous bytes; SBO (between 8-row groups) = 128 B.
struct WeightImage {
  int k[4];          // K of each layer (in_pad, 64, 64, 64)
  int n[4];          // N of each layer (64, 64, 64, 16)
  int off[4];        // byte offset of each layer's block
  int off_w3f32;     // byte offset of the fp32 copy of W3 row 0 (64 floats): single-output layers run on
                     // the CUDA cores straight from the fp32 accumulators (no fp16 rounding of h3)
  int bytes;
};

__host__ __device__ inline WeightImage weight_image(int in_pad) {
  WeightImage w{};
  const int ks[4] = {in_pad, 64, 64, 64};
  const int ns[4] = {64, 64, 64, 16};
  int o = 0;
  for (int l = 0; l < 4; ++l) {
    w.k[l] = ks[l];
    w.n[l] = ns[l];
    w.off[l] = o;
    o += ks[l] * ns[l] * 2;
  }
  w.off_w3f32 = o;
  o += 64 * 4;
  w.bytes = o;
  return w;
}

// ---- first-layer input order -------------------------------------------------------------------
// Input-column order of the decoder's first layer as the tensor-core path feeds it.  The 17 inputs
// [l(3) sin(3) cos(3) feat(8)] + 15 ones are permuted (W0's columns are permuted identically when the
// weight image is packed) so that every fp16x2 word of a row depends on ONE axis or is a gathered
// feature pair: [f0..f7 | lx sx cx 1 | ly sy cy 1 | lz sz cz 1 | 1 x 12].  The per-axis words exist in a
// floor and a ceil flavour computed once per query; a corner row is then 4 gathered words + 6 selected
// words + 6 constants -- no per-corner sincos / conversions.
__host__ __device__ constexpr int dec_perm(int pos) {
  return pos < 8 ? 9 + pos
         : pos < 20 ? ((pos - 8) % 4 == 3 ? 17 + (pos - 8) / 4 : (pos - 8) / 4 + 3 * ((pos - 8) % 4))
                    : pos;
}
// encoder: [x 1 | y 1 | z 1 | n0 n1 | n2 1 | 1 x 6]
__host__ __device__ constexpr int enc_perm(int pos) {
  return pos == 0 ? 0 : pos == 1 ? 6 : pos == 2 ? 1 : pos == 3 ? 7 : pos == 4 ? 2 : pos == 5 ? 8 : pos == 6 ? 3
         : pos == 7 ? 4 : pos == 8 ? 5 : pos;
}

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// wait of a latency-tolerant warp: the hardware may keep the thread suspended for up to `hint_ns` per attempt, so a
// waiting warp does not compete for issue slots with the warps on the critical path
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
  }
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]; issued by ONE thread
__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all tcgen05 ops issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// shared-memory matrix descriptor: K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}
// instruction descriptor: f16 x f16 -> f32, both K-major, M = 128 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t idesc_f16_m128(int n, bool acc_f32 = true) {
  return (acc_f32 ? (1u << 4) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ---- TMEM <-> registers: 32 lanes x 32-bit, thread i of the warp <-> lane (32 * (warp % 4) + i) --
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 registers from the .pack::16b view (fp16 accumulators; layout probed by tools/umma_bench.cu)
__device__ __forceinline__ void tmem_ld32_pack16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// two fp32 -> packed fp16x2 (lo = a, hi = b), optionally with ReLU, one instruction
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_relu_f16x2(uint32_t lo_bits, uint32_t hi_bits) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
  return r;
}

__device__ __forceinline__ void wg_sync(int bar_id) { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); }

// TMEM slot of a chain (128 columns): D [0,64) fp32 | A_h [64,96) packed fp16 hidden activations |
// A_in [96,112) packed fp16 input row of the next item | D_out [112,128) fp32 output layer.
constexpr int kInCol = 96;
constexpr int kOutCol = 112;


// one lane of a converged warp (elect.sync): with a warp-uniform enclosing branch and warp-uniform
// operands ptxas emits straight-line UTCHMMA / UTCBAR; a thread-divergent `if (tid == x)` instead makes it
// wrap every tcgen05.mma / commit into an ELECT + R2UR.BROADCAST waterfall loop (~45 cycles per MMA and
// ~250 for the commit, tools/umma_bench2.cu)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

}  // namespace tc
}  // namespace bnv
