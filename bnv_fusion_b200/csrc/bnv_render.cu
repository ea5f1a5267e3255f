// Ray sampling and SDF loss of the global optimisation step.
//
// Reference being replaced: calculate_loss (src/utils/render_utils.py:559-594) = render_with_rays (:461-507:
// get_camera_params :426-458 + lift :405-423, hierarchical_sampling :190-233 over stratified_sampling :76-95,
// get_neighbors + count_optim, decode_pts) + compute_sdf_loss (:510-557) -- about a hundred small PyTorch kernels per
// inner step of NeuralMap.optimize (src/run_e2e.py:111-156).  Here one step is five launches:
//   ray_samples_kernel (this file)  ->  count_optim over the samples' corners (bnv_map.cu)  ->  fused decode
//   (bnv_decode_sdf)  ->  ray_sdf_loss_kernel (this file: loss + d loss / d sdf)  ->  decode backward (bnv_decode.cu).
// float32 with the reference's operation order (separately rounded multiplies and adds: PyTorch runs every op as its
// own kernel); the stratified random draws are inputs.  The reference sorts the 35 samples of a ray by distance; the
// loss is a sum over samples, so the order is immaterial and no sort is done.
#include "bnv_common.cuh"

namespace bnv {

struct RayCam {
  float K[9];
  float T[16];
};

// torch.linspace(0, 1, steps)[i] in float32
__device__ __forceinline__ float linspace01(int i, int steps) {
  const float step = __fdiv_rn(1.0f, (float)(steps - 1));
  return i < steps / 2 ? __fadd_rn(0.0f, __fmul_rn(step, (float)i)) : __fsub_rn(1.0f, __fmul_rn(step, (float)(steps - 1 - i)));
}

__device__ __forceinline__ float norm3(float x, float y, float z) {
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// get_camera_params (:426-458): unit ray direction of pixel (x, y)
__device__ __forceinline__ void ray_direction(const RayCam& c, float x, float y, float (&dir)[3]) {
  const float fx = c.K[0], fy = c.K[4], cx = c.K[2], cy = c.K[5], sk = c.K[1];
  const float z = __fadd_rn(__fmul_rn(x, 0.0f), 1.0f);
  const float xl = __fmul_rn(__fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(x, cx), __fdiv_rn(__fmul_rn(cy, sk), fy)),
                                                  __fdiv_rn(__fmul_rn(sk, y), fy)), fx), z);
  const float yl = __fmul_rn(__fdiv_rn(__fsub_rn(y, cy), fy), z);
  const float cam[4] = {xl, yl, z, 1.0f};
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float acc = __fmul_rn(c.T[r * 4], cam[0]);
#pragma unroll
    for (int k = 1; k < 4; ++k) acc = __fadd_rn(acc, __fmul_rn(c.T[r * 4 + k], cam[k]));
    d[r] = __fsub_rn(acc, c.T[r * 4 + 3]);
  }
  const float n = fmaxf(norm3(d[0], d[1], d[2]), 1e-12f);                              // F.normalize
#pragma unroll
  for (int r = 0; r < 3; ++r) dir[r] = __fdiv_rn(d[r], n);
}

// stratified_sampling (:76-95), sample j of `steps` over [0, distance]
__device__ __forceinline__ float stratified(int j, int steps, float distance, float t_rand) {
  const float b = __fmul_rn(linspace01(j, steps), distance);
  const float lower = j == 0 ? b : __fmul_rn(0.5f, __fadd_rn(b, __fmul_rn(linspace01(j - 1, steps), distance)));
  const float upper = j == steps - 1 ? b : __fmul_rn(0.5f, __fadd_rn(__fmul_rn(linspace01(j + 1, steps), distance), b));
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand));
}

// one thread per (ray, sample): hierarchical_sampling (:190-233) -> point on the ray
__global__ void __launch_bounds__(256) ray_samples_kernel(const float* __restrict__ uv, const float* __restrict__ gt_pts,
                                                          int64_t n_rays, RayCam cam, const float* __restrict__ t_fine,
                                                          int n_fine, const float* __restrict__ t_coarse, int n_coarse,
                                                          float off, float two_off, float* __restrict__ pts) {
  const int S = n_fine + n_coarse;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rays * S) return;
  const int64_t i = t / S;
  const int j = (int)(t - i * S);
  float dir[3];
  ray_direction(cam, uv[i * 2], uv[i * 2 + 1], dir);
  const float cl[3] = {cam.T[3], cam.T[7], cam.T[11]};
  const float g[3] = {gt_pts[i * 3], gt_pts[i * 3 + 1], gt_pts[i * 3 + 2]};
  const float gt_depth = norm3(__fsub_rn(g[0], cl[0]), __fsub_rn(g[1], cl[1]), __fsub_rn(g[2], cl[2]));
  float dist;
  if (j < n_fine) {
    const float half = __fadd_rn(0.0f, off);
    const float neg = __fsub_rn(gt_depth, off) < 0.f ? gt_depth : half;
    float sp[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) sp[r] = __fsub_rn(__fsub_rn(g[r], __fmul_rn(neg, dir[r])), cl[r]);
    const float start_depth = norm3(sp[0], sp[1], sp[2]);
    dist = __fadd_rn(stratified(j, n_fine, __fadd_rn(0.0f, two_off), t_fine[i * n_fine + j]), start_depth);
  } else {
    dist = stratified(j - n_fine, n_coarse, gt_depth, t_coarse[i * n_coarse + (j - n_fine)]);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) pts[t * 3 + r] = __fadd_rn(cl[r], __fmul_rn(dist, dir[r]));
}

// one thread per (ray, sample): compute_sdf_loss (:510-557) -> loss contribution and d loss / d pred_sdf
__global__ void __launch_bounds__(256) ray_sdf_loss_kernel(const float* __restrict__ pts, const float* __restrict__ pred,
                                                           int64_t n_rays, int S, const float* __restrict__ gt_pts,
                                                           float cx, float cy, float cz, const float* __restrict__ nbr,
                                                           const float* __restrict__ nbr_mask, int n_nbr,
                                                           const float* __restrict__ ray_mask,
                                                           const float* __restrict__ n_valid, float td, float valid_thr,
                                                           double* __restrict__ loss, float* __restrict__ grad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (t < n_rays * S) {
    const int64_t i = t / S;
    const float p[3] = {pts[t * 3], pts[t * 3 + 1], pts[t * 3 + 2]};
    const float gt_depth = norm3(__fsub_rn(gt_pts[i * 3], cx), __fsub_rn(gt_pts[i * 3 + 1], cy), __fsub_rn(gt_pts[i * 3 + 2], cz));
    const float depth = norm3(__fsub_rn(p[0], cx), __fsub_rn(p[1], cy), __fsub_rn(p[2], cz));
    const float gt_sdf = fminf(fmaxf(__fsub_rn(gt_depth, depth), -td), td);
    const bool valid = gt_sdf > valid_thr;
    float nearest = 3.0e38f;
    for (int k = 0; k < n_nbr; ++k) {
      const float* q = nbr + (i * n_nbr + k) * 3;
      const float d = nbr_mask[i * n_nbr + k] != 0.f ? norm3(__fsub_rn(q[0], p[0]), __fsub_rn(q[1], p[1]), __fsub_rn(q[2], p[2]))
                                                      : 10000.0f;
      nearest = fminf(nearest, d);
    }
    const float sign = gt_sdf > 0.f ? 1.0f : -1.0f;
    const float target = fminf(fmaxf(__fmul_rn(nearest, sign), -td), td);
    const float diff = __fsub_rn(pred[t], target);
    const float m = ray_mask[i], nv = *n_valid;
    if (valid) contrib = __fmul_rn(fabsf(diff), m);
    grad[t] = valid ? __fdiv_rn(__fmul_rn(diff > 0.f ? 1.0f : diff < 0.f ? -1.0f : 0.0f, m), nv) : 0.0f;
  }
  // block sum in double (the order of a float32 sum over 35 000 terms would show at the 1e-6 level)
  double v = (double)contrib;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s[w];
    if (tot != 0.0) atomicAdd(loss, tot / (double)*n_valid);
  }
}

}  // namespace bnv

using namespace bnv;

extern "C" {

int bnv_ray_samples(const float* uv, const float* gt_pts, int64_t n_rays, const float* K, const float* T_wc,
                    const float* t_fine, int n_fine, const float* t_coarse, int n_coarse, double truncated_dist,
                    float* pts, void* stream) {
  if (n_rays < 0 || n_fine < 2 || n_coarse < 2 || !K || !T_wc || (n_rays > 0 && (!uv || !gt_pts || !t_fine || !t_coarse || !pts))) {
    set_error("bnv_ray_samples: bad argument");
    return BNV_E_ARG;
  }
  if (n_rays == 0) return BNV_OK;
  RayCam cam;
  for (int i = 0; i < 9; ++i) cam.K[i] = K[i];
  for (int i = 0; i < 16; ++i) cam.T[i] = T_wc[i];
  const int64_t total = n_rays * (n_fine + n_coarse);
  // `offset_distance * 2` is a Python float product before it meets the float32 tensor (:215)
  ray_samples_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      uv, gt_pts, n_rays, cam, t_fine, n_fine, t_coarse, n_coarse, (float)truncated_dist, (float)(truncated_dist * 2), pts);
  BNV_LAUNCH_CHECK("ray_samples_kernel");
  return BNV_OK;
}

int bnv_ray_sdf_loss(const float* pts, const float* pred_sdf, int64_t n_rays, int n_samples, const float* gt_pts,
                     const float* T_wc, const float* nbr_pts, const float* nbr_mask, int n_nbr, const float* ray_mask,
                     const float* n_valid, double truncated_dist, double* loss, float* grad_pred, void* stream) {
  if (n_rays < 0 || n_samples <= 0 || n_nbr < 0 || !T_wc || !n_valid || !loss ||
      (n_rays > 0 && (!pts || !pred_sdf || !gt_pts || !ray_mask || !grad_pred || (n_nbr > 0 && (!nbr_pts || !nbr_mask))))) {
    set_error("bnv_ray_sdf_loss: bad argument");
    return BNV_E_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  BNV_CUDA(cudaMemsetAsync(loss, 0, sizeof(double), s));
  if (n_rays == 0) return BNV_OK;
  const int64_t total = n_rays * n_samples;
  const double thr = -truncated_dist * 0.5 > -0.05 ? -truncated_dist * 0.5 : -0.05;            // max(-td * 0.5, -0.05), :526
  ray_sdf_loss_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(pts, pred_sdf, n_rays, n_samples, gt_pts, T_wc[3], T_wc[7],
                                                                      T_wc[11], nbr_pts, nbr_mask, n_nbr, ray_mask, n_valid,
                                                                      (float)truncated_dist, (float)thr, loss, grad_pred);
  BNV_LAUNCH_CHECK("ray_sdf_loss_kernel");
  return BNV_OK;
}

}  // extern "C"
