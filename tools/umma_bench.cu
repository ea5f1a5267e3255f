// Micro-benchmarks that shaped the tensor-core MLP chain (results in profiles/r1_umma_microbench.txt):
//   1. cycles per tcgen05.mma (M=128, K=16, A in TMEM) as a function of N, 1 and 4 issuing threads
//   2. single-MMA round trip: issue -> commit -> mbarrier wake
//   3. tcgen05.ld / tcgen05.st (32x32b.x32) cost per warp, 1 / 4 / 8 / 16 warps
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../bnv_fusion_b200/csrc umma_bench.cu -o umma_bench
#include <cstdio>
#include "bnv_tc.cuh"
using namespace bnv::tc;

struct Sh { uint64_t bar[16]; uint32_t tmem; };

template <int N>
__global__ void k_mma(int reps, int n_issuers, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Sh& sh = *reinterpret_cast<Sh*>(smem);
  uint8_t* w = smem + 1024;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(w)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&sh.bar[i], 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&sh.tmem, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int warp = threadIdx.x >> 5;
  long long t0 = 0, t1 = 0;
  if ((threadIdx.x & 31) == 0 && warp < n_issuers) {
    const uint32_t d = sh.tmem + warp * 64 * (N > 64 ? 0 : 1);   // distinct D regions when they fit
    const uint32_t a = sh.tmem + 448;
    const uint32_t lbo = (N / 8) * 128;
    const uint64_t desc = smem_desc_kmajor(smem_u32(w), lbo, 128);
    const uint32_t idesc = idesc_f16_m128(N);
    t0 = clock64();
    for (int r = 0; r < reps; ++r) umma_ts_f16(d, a, desc, idesc, 1u);
    umma_commit(&sh.bar[warp]);
    long long ti = clock64();
    mbar_wait(&sh.bar[warp], 0);
    t1 = clock64();
    out[warp * 2] = t1 - t0;
    out[warp * 2 + 1] = ti - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(sh.tmem, 512);
}

__global__ void k_roundtrip(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Sh& sh = *reinterpret_cast<Sh*>(smem);
  uint8_t* w = smem + 1024;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(w)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(&sh.bar[0], 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&sh.tmem, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    const uint64_t desc = smem_desc_kmajor(smem_u32(w), 1024, 128);
    uint32_t par = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int kk = 0; kk < 4; ++kk) umma_ts_f16(sh.tmem, sh.tmem + 448, desc, idesc_f16_m128(64), kk);
      umma_commit(&sh.bar[0]);
      mbar_wait(&sh.bar[0], par); par ^= 1;
      tc_fence_after();
    }
    out[0] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(sh.tmem, 512);
}

// mode 0: ld x32 only; 1: st x32 + wait; 2: ld x32 x2 + 32 cvt + st x32 (the v1 hidden epilogue)
__global__ void k_ldst(int reps, int mode, long long* out) {
  __shared__ Sh sh;
  if (threadIdx.x < 32) tmem_alloc(&sh.tmem, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int warp = threadIdx.x >> 5;
  const uint32_t t = sh.tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 96 % 416;
  uint32_t v[32], w[32];
  for (int i = 0; i < 32; ++i) { v[i] = i + threadIdx.x; w[i] = i; }
  tmem_st32(t, v); tmem_st32(t + 32, v); tmem_wait_st();
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) { tmem_ld32(t, v); tmem_wait_ld(); }
    else if (mode == 1) { tmem_st32(t, v); tmem_wait_st(); }
    else {
      tmem_ld32(t, v); tmem_ld32(t + 32, w); tmem_wait_ld();
      uint32_t a[32];
      for (int i = 0; i < 16; ++i) { a[i] = pack_relu_f16x2(v[2 * i], v[2 * i + 1]); a[16 + i] = pack_relu_f16x2(w[2 * i], w[2 * i + 1]); }
      tmem_st32(t + 64, a); tmem_wait_st();
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 32; ++i) s += v[i] + w[i];
  if ((threadIdx.x & 31) == 0) { out[warp] = t1 - t0; out[32 + warp] = s; }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(sh.tmem, 512);
}

// fp16 accumulators: where do the 16-bit D elements live in TMEM?  A = ones [128 x 16], B[n][k] = (k == 0) * (n + 1) / 64
// => D[m][n] = (n + 1) / 64.  Dump lane 0: raw 32-bit columns 0..63, and the .pack::16b view of the same columns.
__global__ void k_probe_f16acc(uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Sh& sh = *reinterpret_cast<Sh*>(smem);
  __half* w = reinterpret_cast<__half*>(smem + 1024);
  for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) w[i] = __float2half(0.f);
  __syncthreads();
  if (threadIdx.x < 64) {            // canonical K-major no-swizzle: byte(n,k) = (k/8)*LBO + (n/8)*128 + (n%8)*16 + (k%8)*2, LBO = 1024
    const int n = threadIdx.x;
    w[((0 / 8) * 1024 + (n / 8) * 128 + (n % 8) * 16) / 2] = __float2half((n + 1) / 64.f);
  }
  if (threadIdx.x == 0) { mbar_init(&sh.bar[0], 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&sh.tmem, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int warp = threadIdx.x >> 5;
  const uint32_t tl = sh.tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t ones[8];
  for (int i = 0; i < 8; ++i) ones[i] = 0x3C003C00u;
  uint32_t junk[32];
  for (int i = 0; i < 32; ++i) junk[i] = 0xDEAD0000u + i;
  tmem_st32(tl, junk); tmem_st32(tl + 32, junk);
  tmem_st8(tl + 448, ones);
  tmem_wait_st(); tc_fence_before(); __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    umma_ts_f16(sh.tmem, sh.tmem + 448, smem_desc_kmajor(smem_u32(w), 1024, 128), idesc_f16_m128(64, false), 0u);
    umma_commit(&sh.bar[0]);
  }
  mbar_wait(&sh.bar[0], 0);
  tc_fence_after();
  uint32_t a[32], b[32], p[32];
  tmem_ld32(tl, a); tmem_ld32(tl + 32, b); tmem_wait_ld();
  tmem_ld32_pack16(tl, p); tmem_wait_ld();
  if (threadIdx.x == 0) for (int i = 0; i < 32; ++i) { out[i] = a[i]; out[32 + i] = b[i]; out[64 + i] = p[i]; }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(sh.tmem, 512);
}

template <int N> void run_mma(long long* d, long long* h) {
  cudaFuncSetAttribute(k_mma<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int iss : {1, 4}) {
    if (N > 64 && iss > 1) continue;
    const int reps = 512;
    k_mma<N><<<1, 128, 40000>>>(reps, iss, d);
    cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
    printf("mma M=128 N=%3d K=16 TS  issuers=%d reps=%d : %.1f cyc/mma (issue loop %.1f cyc/mma)  [nominal %d]\n", N, iss,
           reps, (double)h[0] / reps / 1.0, (double)h[1] / reps, 128 * N / 256);
  }
}

int main() {
  long long *d, h[64];
  cudaMalloc(&d, 64 * 8);
  run_mma<16>(d, h); run_mma<32>(d, h); run_mma<64>(d, h); run_mma<128>(d, h); run_mma<256>(d, h);
  cudaFuncSetAttribute(k_roundtrip, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  k_roundtrip<<<1, 128, 40000>>>(200, d);
  cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("round trip 4x(M128 N64 K16) + commit + mbarrier wait: %.1f cyc (MMA nominal 128)\n", (double)h[0] / 200);
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      k_ldst<<<1, warps * 32>>>(200, mode, d);
      cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("%s warps=%2d : %.1f cyc/iter (slowest warp)\n", mode == 0 ? "LDTM.x32+wait      " : mode == 1 ? "STTM.x32+wait      " : "ld64+cvt32+st32    ", warps, (double)mx / 200);
    }
  {
    uint32_t *dp, hp[96];
    cudaMalloc(&dp, 96 * 4);
    cudaFuncSetAttribute(k_probe_f16acc, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    k_probe_f16acc<<<1, 128, 40000>>>(dp);
    cudaMemcpy(hp, dp, 96 * 4, cudaMemcpyDeviceToHost);
    printf("fp16-accumulator probe, lane 0 (expect D[n] = (n+1)/64 as fp16: 0x2400 0x2800 0x2A00 0x2C00 ...)\n raw cols 0..15 :");
    for (int i = 0; i < 16; ++i) printf(" %08x", hp[i]);
    printf("\n raw cols 32..39:");
    for (int i = 32; i < 40; ++i) printf(" %08x", hp[i]);
    printf("\n pack::16b 0..15:");
    for (int i = 64; i < 80; ++i) printf(" %08x", hp[i]);
    printf("\n");
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
