// Where do the ~950 cycles of one MLP round of the chain go?  Four row warpgroups per CTA (like the real
// kernels) run hidden-layer rounds (K = 64, N = 64) back to back; lane 0 of warp 0 of every warpgroup
// timestamps the phases of its round.  Uses the production chain code (bnv_tc.cuh).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../bnv_fusion_b200/csrc chain_profile.cu -o chain_profile
#include <cstdio>
#include "bnv_tc.cuh"
using namespace bnv::tc;

__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }

struct alignas(16) Smem { TcShared<4> sh; };

template <int NWG>
__global__ void __launch_bounds__(NWG * 128, 1) k_chain(const uint8_t* gW, int w_bytes, int rounds, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  Smem& S = *reinterpret_cast<Smem*>(smem);
  uint8_t* s_w = smem + 256;
  RowChain c = tc_setup<4>(S.sh, s_w, gW, w_bytes);
  const int wg = threadIdx.x >> 7;
  const bool rec = (threadIdx.x & 127) == 0;
  uint32_t in[16];
  for (int i = 0; i < 16; ++i) in[i] = 0x3C003C00u;
  long long a_ld = 0, a_cvt = 0, a_st = 0, a_sync = 0, a_issue = 0, a_wait = 0;
  constexpr int off1 = 32 * 64 * 2;
  chain_stage<16>(c, in);
  chain_begin<16>(c);
  long long t_issue = clk();
  const long long t_start = t_issue;
  for (int r = 0; r < rounds; ++r) {
    chain_wait_d(c);
    const long long t0 = clk();
    uint32_t v[32], w[32];
    tmem_ld32(c.t_d, v);
    tmem_ld32(c.t_d + 32, w);
    tmem_wait_ld();
    const long long t1 = clk();
    uint32_t a[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      a[i] = pack_relu_f16x2(v[2 * i], v[2 * i + 1]);
      a[16 + i] = pack_relu_f16x2(w[2 * i], w[2 * i + 1]);
    }
    tmem_st32(c.t_a, a);
    const long long t2 = clk();
    tmem_wait_st();
    tc_fence_before();
    const long long t3 = clk();
    wg_sync(c.bar_id);
    const long long t4 = clk();
    if (c.issuer_warp) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_step<64>(c, c.w_saddr + off1, 0, kACol, kk);
        umma_commit(c.bar_d);
      }
    }
    const long long t5 = clk();
    a_wait += t0 - t_issue; a_ld += t1 - t0; a_cvt += t2 - t1; a_st += t3 - t2; a_sync += t4 - t3; a_issue += t5 - t4;
    t_issue = t5;
  }
  chain_wait_d(c);
  const long long t_end = clk();
  if (rec && wg < NWG) {
    long long* o = out + (blockIdx.x * 4 + wg) * 8;
    o[0] = a_wait; o[1] = a_ld; o[2] = a_cvt; o[3] = a_st; o[4] = a_sync; o[5] = a_issue; o[6] = t_end - t_start;
  }
  tc_teardown<4>(S.sh);
}

template <int NWG> void run(const uint8_t* dW, int wbytes, long long* d, long long* h) {
  const int rounds = 400;
  cudaFuncSetAttribute(k_chain<NWG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  k_chain<NWG><<<148, NWG * 128, 60000>>>(dW, wbytes, rounds, d);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, 148 * 4 * 8 * 8, cudaMemcpyDeviceToHost);
  double s[7] = {0};
  for (int b = 0; b < 148; ++b) for (int g = 0; g < NWG; ++g) for (int j = 0; j < 7; ++j) s[j] += (double)h[(b * 4 + g) * 8 + j];
  const double n = 148.0 * NWG * rounds;
  printf("%d chains/SM: round %.0f cyc = issue->wake %.0f + LDTM x2 %.0f + cvt+STTM issue %.0f + wait::st %.0f + bar.sync %.0f + elect/issue %.0f   (tensor pipe busy %.0f %%)  [%s]\n",
         NWG, s[6] / n, s[0] / n, s[1] / n, s[2] / n, s[3] / n, s[4] / n, s[5] / n, 100.0 * NWG * 128.0 / (s[6] / n), cudaGetErrorString(e));
}

int main() {
  const int wbytes = 32 * 64 * 2 + 2 * 64 * 64 * 2 + 16 * 64 * 2 + 256;
  uint8_t* dW; cudaMalloc(&dW, wbytes); cudaMemset(dW, 0, wbytes);
  long long *d, *h = (long long*)malloc(148 * 4 * 8 * 8);
  cudaMalloc(&d, 148 * 4 * 8 * 8);
  run<1>(dW, wbytes, d, h); run<2>(dW, wbytes, d, h); run<3>(dW, wbytes, d, h); run<4>(dW, wbytes, d, h);
  return 0;
}
