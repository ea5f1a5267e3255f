// fp32 CUDA-core evaluation of the tiny tcnn-style MLPs (BNV_MLP_FP32: the exact-parity mode).
//
// One thread evaluates one row: n_in -> 64 -> 64 -> 64 -> n_out, ReLU, no bias, ones-padded input
// (reference: tcnn FullyFusedMLP behind src/utils/pointnet_utils.py:274-286 and
// src/models/fusion/modules.py:171-176,249-253).  Weights sit in shared memory in k-major
// ("transposed") fp32 blocks so a warp reads them as broadcast LDS.128; the 64 accumulators live
// in registers with compile-time indices; the hidden activations of a thread live in a private,
// bank-conflict-free column of shared memory (element j at sH[j * stride]).
#pragma once
#include "bnv_common.cuh"

namespace bnv {

// k-major weight image built on the host by bnv_mlp_create:
//   T0 [NIN][64] | B0 [64] (= sum of the ones-padded input columns of W0) | T1 [64][64] | T2 [64][64]
//   | T3 [64][NOUT]
template <int NIN, int NOUT>
struct SimtMlp {
  static constexpr int kT0 = 0;
  static constexpr int kB0 = NIN * 64;
  static constexpr int kT1 = kB0 + 64;
  static constexpr int kT2 = kT1 + 64 * 64;
  static constexpr int kT3 = kT2 + 64 * 64;
  static constexpr int kFloats = kT3 + 64 * NOUT;

  __device__ static __forceinline__ void hidden_layer(const float* __restrict__ T, float* sH, int stride,
                                                      float (&acc)[64]) {
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
#pragma unroll 2
    for (int k = 0; k < 64; ++k) {
      const float hk = sH[k * stride];
      const float4* w = reinterpret_cast<const float4*>(T + k * 64);
#pragma unroll
      for (int j4 = 0; j4 < 16; ++j4) {
        const float4 v = w[j4];
        acc[4 * j4 + 0] = fmaf(v.x, hk, acc[4 * j4 + 0]);
        acc[4 * j4 + 1] = fmaf(v.y, hk, acc[4 * j4 + 1]);
        acc[4 * j4 + 2] = fmaf(v.z, hk, acc[4 * j4 + 2]);
        acc[4 * j4 + 3] = fmaf(v.w, hk, acc[4 * j4 + 3]);
      }
    }
  }

  // sW: shared weights (kFloats), sH: this thread's activation column, x: input row, y: output row
  __device__ static __forceinline__ void run(const float* __restrict__ sW, float* sH, int stride,
                                             const float (&x)[NIN], float (&y)[NOUT]) {
    float acc[64];
    {
      const float4* b = reinterpret_cast<const float4*>(sW + kB0);
#pragma unroll
      for (int j4 = 0; j4 < 16; ++j4) {
        const float4 v = b[j4];
        acc[4 * j4 + 0] = v.x; acc[4 * j4 + 1] = v.y; acc[4 * j4 + 2] = v.z; acc[4 * j4 + 3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < NIN; ++k) {
        const float4* w = reinterpret_cast<const float4*>(sW + kT0 + k * 64);
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const float4 v = w[j4];
          acc[4 * j4 + 0] = fmaf(v.x, x[k], acc[4 * j4 + 0]);
          acc[4 * j4 + 1] = fmaf(v.y, x[k], acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(v.z, x[k], acc[4 * j4 + 2]);
          acc[4 * j4 + 3] = fmaf(v.w, x[k], acc[4 * j4 + 3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 64; ++j) sH[j * stride] = fmaxf(acc[j], 0.f);
    hidden_layer(sW + kT1, sH, stride, acc);
#pragma unroll
    for (int j = 0; j < 64; ++j) sH[j * stride] = fmaxf(acc[j], 0.f);
    hidden_layer(sW + kT2, sH, stride, acc);
#pragma unroll
    for (int o = 0; o < NOUT; ++o) y[o] = 0.f;
#pragma unroll
    for (int k = 0; k < 64; ++k) {
      const float hk = fmaxf(acc[k], 0.f);
#pragma unroll
      for (int o = 0; o < NOUT; ++o) y[o] = fmaf(sW[kT3 + k * NOUT + o], hk, y[o]);
    }
  }
};

// cooperative copy of the weight image into shared memory
__device__ __forceinline__ void load_weights(float* sW, const float* __restrict__ gW, int n_floats) {
  const float4* g = reinterpret_cast<const float4*>(gW);
  float4* s = reinterpret_cast<float4*>(sW);
  for (int i = threadIdx.x; i < n_floats / 4; i += blockDim.x) s[i] = __ldg(g + i);
  __syncthreads();
}

}  // namespace bnv
