// tcgen05 tensor-core variants (BNV_MLP_TC16).  Placeholder until the UMMA kernels land: the entry
// points fail loudly instead of silently falling back.
#include "bnv_common.cuh"
#include "bnv_decode_common.cuh"

using namespace bnv;

int bnv_internal_pack_tc_weights(bnv_mlp_t* mlp, const float* params_host) {
  (void)mlp; (void)params_host;
  return BNV_OK;
}

int bnv_internal_mlp_forward_tc(const bnv_mlp_t*, const float*, int64_t, float*, cudaStream_t) {
  set_error("BNV_MLP_TC16 forward is not built yet");
  return BNV_E_UNSUPPORTED;
}

int bnv_internal_encode_tc(bnv_map_t*, const void*, int, int64_t, const bnv_mlp_t*, cudaStream_t) {
  set_error("BNV_MLP_TC16 encode is not built yet");
  return BNV_E_UNSUPPORTED;
}

int bnv_internal_decode_tc(bnv_map_t*, const bnv::DecArgs&, const bnv_mlp_t*, cudaStream_t) {
  set_error("BNV_MLP_TC16 decode is not built yet");
  return BNV_E_UNSUPPORTED;
}
