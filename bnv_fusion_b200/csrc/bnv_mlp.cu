// MLP weight handles (tcnn.NetworkWithInputEncoding replacement) and the plain batched forward.
// Reference: src/utils/pointnet_utils.py:269-294 (tcnnPointNetEncoder), src/models/fusion/modules.py:
// 169-176,249-253 (tcnnNeRFModel.geo_forward), src/models/tcnn_config.json:24-30.
#include <cuda_fp16.h>

#include <vector>

#include "bnv_common.cuh"
#include "bnv_mlp_simt.cuh"

using namespace bnv;

// bnv_tc.cu: builds the fp16 UMMA-layout weight image
int bnv_internal_pack_tc_weights(bnv_mlp_t* mlp, const float* params_host);
int bnv_internal_mlp_forward_tc(const bnv_mlp_t* mlp, const float* x, int64_t n, float* y, cudaStream_t s);

const float* bnv_internal_simt_weights(const bnv_mlp_t* mlp) { return mlp->w32; }

namespace bnv {

template <int NIN, int NOUT>
__global__ void __launch_bounds__(256) mlp_forward_simt_kernel(const float* __restrict__ gW, const float* __restrict__ x,
                                                               int64_t n, float* __restrict__ y) {
  using M = SimtMlp<NIN, NOUT>;
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* sH = smem + M::kFloats + threadIdx.x;
  load_weights(sW, gW, M::kFloats);
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float xi[NIN], yo[NOUT];
#pragma unroll
  for (int k = 0; k < NIN; ++k) xi[k] = x[i * NIN + k];
  M::run(sW, sH, 256, xi, yo);
#pragma unroll
  for (int o = 0; o < NOUT; ++o) y[i * NOUT + o] = yo[o];
}

template <int NIN, int NOUT>
static int launch_forward(const bnv_mlp_t* mlp, const float* x, int64_t n, float* y, cudaStream_t s) {
  using M = SimtMlp<NIN, NOUT>;
  const size_t smem = (size_t)(M::kFloats + 64 * 256) * sizeof(float);
  // per device and cheap: set on every call rather than cached per process
  BNV_CUDA(cudaFuncSetAttribute(mlp_forward_simt_kernel<NIN, NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mlp_forward_simt_kernel<NIN, NOUT><<<(unsigned)((n + 255) / 256), 256, smem, s>>>(mlp->w32, x, n, y);
  BNV_LAUNCH_CHECK("mlp_forward_simt_kernel");
  return BNV_OK;
}

}  // namespace bnv

extern "C" {

int bnv_mlp_create(bnv_mlp_t** out, const float* params, int64_t n_params, int n_in, int n_out, int device) {
  if (!out || !params) { set_error("bnv_mlp_create: null argument"); return BNV_E_ARG; }
  if (!((n_in == 6 && n_out == 8) || (n_in == 17 && n_out == 1))) {
    set_error("bnv_mlp_create: supported shapes are 6->8 (encoder) and 17->1 (decoder), got %d->%d", n_in, n_out);
    return BNV_E_UNSUPPORTED;
  }
  const int in_pad = (n_in + 15) / 16 * 16, out_pad = (n_out + 15) / 16 * 16;
  const int64_t expect = (int64_t)kWidth * in_pad + 2 * kWidth * kWidth + (int64_t)out_pad * kWidth;
  if (n_params != expect) {
    set_error("bnv_mlp_create: %d->%d needs %lld parameters, got %lld", n_in, n_out, (long long)expect, (long long)n_params);
    return BNV_E_ARG;
  }
  BNV_CUDA(cudaSetDevice(device));
  bnv_mlp* m = new bnv_mlp();
  memset(m, 0, sizeof(*m));
  m->n_in = n_in; m->n_out = n_out; m->in_pad = in_pad; m->out_pad = out_pad; m->device = device;
  m->n_params = n_params;
  // k-major fp32 image for the CUDA-core path (layout documented in bnv_mlp_simt.cuh)
  const float* W0 = params;
  const float* W1 = W0 + kWidth * in_pad;
  const float* W2 = W1 + kWidth * kWidth;
  const float* W3 = W2 + kWidth * kWidth;
  std::vector<float> img((size_t)n_in * 64 + 64 + 2 * 4096 + 64 * n_out);
  float* T0 = img.data();
  float* B0 = T0 + n_in * 64;
  float* T1 = B0 + 64;
  float* T2 = T1 + 4096;
  float* T3 = T2 + 4096;
  for (int j = 0; j < 64; ++j) {
    for (int k = 0; k < n_in; ++k) T0[k * 64 + j] = W0[j * in_pad + k];
    float b = 0.f;                       // ones-padded input columns act as a bias
    for (int k = n_in; k < in_pad; ++k) b += W0[j * in_pad + k];
    B0[j] = b;
    for (int k = 0; k < 64; ++k) {
      T1[k * 64 + j] = W1[j * 64 + k];
      T2[k * 64 + j] = W2[j * 64 + k];
    }
  }
  for (int o = 0; o < n_out; ++o)
    for (int k = 0; k < 64; ++k) T3[k * n_out + o] = W3[o * 64 + k];
  cudaError_t e = cudaMalloc((void**)&m->w32, img.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(m->w32, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    bnv_mlp_destroy(m);
    return cuda_fail(e, "bnv_mlp_create upload");
  }
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->wraw, n_params * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(m->wraw, params, n_params * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    bnv_mlp_destroy(m);
    return cuda_fail(e, "bnv_mlp_create upload (raw)");
  }
  int rc = bnv_internal_pack_tc_weights(m, params);
  if (rc != BNV_OK) {
    bnv_mlp_destroy(m);
    return rc;
  }
  *out = m;
  return BNV_OK;
}

int bnv_mlp_destroy(bnv_mlp_t* m) {
  if (!m) return BNV_OK;
  cudaSetDevice(m->device);
  if (m->w32) cudaFree(m->w32);
  if (m->wraw) cudaFree(m->wraw);
  if (m->w16) cudaFree(m->w16);
  delete m;
  return BNV_OK;
}

int bnv_mlp_forward(const bnv_mlp_t* mlp, const float* x, int64_t n, float* y, int mode, void* stream) {
  if (!mlp || n < 0 || (n > 0 && (!x || !y))) { set_error("bnv_mlp_forward: bad argument"); return BNV_E_ARG; }
  if (n == 0) return BNV_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == BNV_MLP_TC16) return bnv_internal_mlp_forward_tc(mlp, x, n, y, s);
  if (mode != BNV_MLP_FP32) { set_error("bnv_mlp_forward: unknown mode %d", mode); return BNV_E_ARG; }
  if (mlp->n_in == 6) return launch_forward<6, 8>(mlp, x, n, y, s);
  return launch_forward<17, 1>(mlp, x, n, y, s);
}

}  // extern "C"
