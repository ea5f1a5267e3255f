// Follow-up micro-benchmarks (round 1, second session): what bounds one MMA round trip of the chain?
//   A. one issuing thread, MMAs alternating over NACC independent accumulators (is 55 cyc/MMA a
//      dependent-accumulate latency or a per-thread issue rate?)
//   B. issue -> commit -> mbarrier wake latency for n MMAs (n = 1, 2, 4, 6), N = 64 and 16
//   C. same with the waiter being a different warp than the issuer
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../bnv_fusion_b200/csrc umma_bench2.cu -o umma_bench2
#include <cstdio>
#include "bnv_tc.cuh"
using namespace bnv::tc;

struct Sh { uint64_t bar[16]; uint32_t tmem; };

__device__ __forceinline__ void setup(Sh& sh, uint8_t* w) {
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(w)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&sh.bar[i], 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&sh.tmem, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
}
__device__ __forceinline__ void teardown(Sh& sh) {
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(sh.tmem, 512);
}

// A: NACC independent accumulators (N columns each), one issuer
template <int N, int NACC>
__global__ void k_indep(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Sh& sh = *reinterpret_cast<Sh*>(smem);
  uint8_t* w = smem + 1024;
  setup(sh, w);
  if (threadIdx.x == 0) {
    const uint32_t a = sh.tmem + 448;
    const uint32_t lbo = (N / 8) * 128;
    const uint64_t desc = smem_desc_kmajor(smem_u32(w), lbo, 128);
    const uint32_t idesc = idesc_f16_m128(N);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int j = 0; j < NACC; ++j) umma_ts_f16(sh.tmem + j * N, a, desc, idesc, 1u);
    }
    long long ti = clock64();
    umma_commit(&sh.bar[0]);
    mbar_wait(&sh.bar[0], 0);
    long long t1 = clock64();
    out[0] = t1 - t0; out[1] = ti - t0;
  }
  teardown(sh);
}

// B/C: latency of n dependent MMAs + commit + wake; waiter = issuer (other = 0) or thread 32 * other
template <int N>
__global__ void k_latency(int n_mma, int other, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Sh& sh = *reinterpret_cast<Sh*>(smem);
  uint8_t* w = smem + 1024;
  setup(sh, w);
  const uint32_t lbo = (N / 8) * 128;
  const uint64_t desc = smem_desc_kmajor(smem_u32(w), lbo, 128);
  const uint32_t idesc = idesc_f16_m128(N);
  uint32_t par = 0;
  long long acc = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x == 0) {
      for (int kk = 0; kk < n_mma; ++kk) umma_ts_f16(sh.tmem, sh.tmem + 448, desc, idesc, kk);
      umma_commit(&sh.bar[0]);
    }
    if (threadIdx.x == 32 * other) {
      mbar_wait(&sh.bar[0], par);
      tc_fence_after();
      acc += clock64() - t0;
    }
    par ^= 1;
    __syncthreads();
  }
  if (threadIdx.x == 32 * other) out[0] = acc;
  teardown(sh);
}

// D: issue cost only: how long is the issuing thread busy for n MMAs + commit (no wait)
template <int N>
__global__ void k_issue_cost(int n_mma, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Sh& sh = *reinterpret_cast<Sh*>(smem);
  uint8_t* w = smem + 1024;
  setup(sh, w);
  const uint32_t lbo = (N / 8) * 128;
  const uint64_t desc = smem_desc_kmajor(smem_u32(w), lbo, 128);
  const uint32_t idesc = idesc_f16_m128(N);
  uint32_t par = 0;
  long long acc = 0;
  if (threadIdx.x == 0) {
    for (int r = 0; r < reps; ++r) {
      long long t0 = clock64();
      for (int kk = 0; kk < n_mma; ++kk) umma_ts_f16(sh.tmem, sh.tmem + 448, desc, idesc, kk);
      umma_commit(&sh.bar[0]);
      acc += clock64() - t0;
      mbar_wait(&sh.bar[0], par); par ^= 1;
      tc_fence_after();
    }
    out[0] = acc;
  }
  teardown(sh);
}

// E: the same round trip with the issue in a warp-uniform branch + elect.sync (straight-line UTCHMMA)
template <int N>
__global__ void k_latency_elect(int n_mma, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Sh& sh = *reinterpret_cast<Sh*>(smem);
  uint8_t* w = smem + 1024;
  setup(sh, w);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t base = __shfl_sync(0xffffffffu, sh.tmem, 0);
  const uint32_t lbo = (N / 8) * 128;
  const uint64_t desc = smem_desc_kmajor(smem_u32(w), lbo, 128);
  const uint32_t idesc = idesc_f16_m128(N);
  uint32_t par = 0;
  long long acc = 0, busy = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    long long t0 = clock64();
    if (warp == 0) {
      if (elect_one()) {
        for (int kk = 0; kk < n_mma; ++kk) umma_ts_f16(base, base + 448, desc, idesc, kk);
        umma_commit(&sh.bar[0]);
      }
      busy += clock64() - t0;
    }
    mbar_wait(&sh.bar[0], par);
    tc_fence_after();
    if (threadIdx.x == 32) acc += clock64() - t0;
    par ^= 1;
    __syncthreads();
  }
  if (threadIdx.x == 32) out[0] = acc;
  if (threadIdx.x == 0) out[1] = busy;
  teardown(sh);
}

// F: where do the ~200 cycles of the issue path go?  timestamps around elect / MMAs / commit / wake
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
template <int N, int NMMA>
__global__ void k_dissect(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Sh& sh = *reinterpret_cast<Sh*>(smem);
  uint8_t* w = smem + 1024;
  setup(sh, w);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t base = __shfl_sync(0xffffffffu, sh.tmem, 0);
  const uint32_t lbo = (N / 8) * 128;
  const uint32_t wa = smem_u32(w);
  uint32_t par = 0;
  long long a1 = 0, a2 = 0, a3 = 0, a4 = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    const long long t0 = clk();
    if (warp == 0) {
      if (elect_one()) {
        const long long t1 = clk();
#pragma unroll
        for (int kk = 0; kk < NMMA; ++kk)
          umma_ts_f16(base, base + 448 + (kk & 3) * 8, smem_desc_kmajor(wa + (kk & 3) * 2 * lbo, lbo, 128), idesc_f16_m128(N), kk > 0);
        const long long t2 = clk();
        umma_commit(&sh.bar[0]);
        const long long t3 = clk();
        a1 += t1 - t0; a2 += t2 - t0; a3 += t3 - t0;
      }
    }
    mbar_wait(&sh.bar[0], par);
    tc_fence_after();
    if (threadIdx.x == 32) a4 += clk() - t0;
    par ^= 1;
    __syncthreads();
  }
  // the elected lane is lane 0 in practice; report from whoever holds non-zero sums
  if (a3 != 0) { out[0] = a1; out[1] = a2; out[2] = a3; }
  if (threadIdx.x == 32) out[3] = a4;
  teardown(sh);
}
template <int N, int NMMA> void run_dissect(long long* d, long long* h) {
  cudaFuncSetAttribute(k_dissect<N, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  k_dissect<N, NMMA><<<1, 128, 40000>>>(200, d);
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("F: N=%2d, %d MMAs (unrolled): elect done +%.0f, MMAs issued +%.0f, commit issued +%.0f, other warp awake +%.0f cyc\n", N, NMMA,
         h[0] / 200.0, h[1] / 200.0, h[2] / 200.0, h[3] / 200.0);
}

template <int N, int NACC> void run_indep(long long* d, long long* h) {
  cudaFuncSetAttribute(k_indep<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const int reps = 256;
  k_indep<N, NACC><<<1, 128, 40000>>>(reps, d);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("A: 1 issuer, N=%3d, %d independent accumulators: %.1f cyc/mma (issue loop %.1f)\n", N, NACC,
         (double)h[0] / reps / NACC, (double)h[1] / reps / NACC);
}
template <int N> void run_lat(long long* d, long long* h) {
  cudaFuncSetAttribute(k_latency<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  cudaFuncSetAttribute(k_issue_cost<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int n : {1, 2, 4, 6, 8})
    for (int other : {0, 1}) {
      k_latency<N><<<1, 128, 40000>>>(n, other, 200, d);
      cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
      printf("B: N=%2d, %d dependent MMAs + commit + wake (%s): %.1f cyc\n", N, n, other ? "waiter = other warp" : "waiter = issuer", (double)h[0] / 200);
    }
  for (int n : {1, 2, 4, 6, 8}) {
    k_issue_cost<N><<<1, 128, 40000>>>(n, 200, d);
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("D: N=%2d, issuing %d MMAs + commit keeps the thread busy %.1f cyc\n", N, n, (double)h[0] / 200);
  }
}

template <int N> void run_elect(long long* d, long long* h) {
  cudaFuncSetAttribute(k_latency_elect<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int n : {1, 2, 4, 6, 8, 16, 32}) {
    k_latency_elect<N><<<1, 128, 40000>>>(n, 200, d);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("E: N=%2d, elect.sync issue of %2d dependent MMAs + commit + wake (other warp): %.1f cyc, issuing warp busy %.1f cyc\n", N, n,
           (double)h[0] / 200, (double)h[1] / 200);
  }
}

int main() {
  long long *d, h[8];
  cudaMalloc(&d, 64);
  run_dissect<64, 1>(d, h); run_dissect<64, 2>(d, h); run_dissect<64, 4>(d, h); run_dissect<64, 6>(d, h); run_dissect<16, 4>(d, h);
  run_elect<64>(d, h); run_elect<16>(d, h);
  run_indep<64, 1>(d, h); run_indep<64, 2>(d, h); run_indep<64, 4>(d, h);
  run_indep<16, 1>(d, h); run_indep<16, 2>(d, h); run_indep<16, 4>(d, h);
  run_lat<64>(d, h); run_lat<16>(d, h);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
