// Coarse projective TSDF prior, resident on the device.
//
// Reference: TSDFVolume (third_parties/fusion.py:22-167), its CPU-mode integrate (:251-294, numba helpers
// :169-206) -- the path the reference runs when PyCUDA is absent and the CPU baseline BASELINE.json names --
// and the consumer NeuralMap.prepare_tsdf_volume (src/run_e2e.py:169-186).  The reference's own inline
// CUDA kernel (:68-141) re-uploads the depth and colour images and 5 small arrays on EVERY launch and
// visits all voxels in a 3-D grid of 1-D blocks; here the volumes, weights and colours stay in HBM, the
// camera goes in kernel parameters, one thread handles one voxel with z fastest (coalesced read-modify-write of
// 12 B per voxel), and whole bricks outside the camera frustum are skipped after 8 corner projections.
//
// Arithmetic follows the CPU mode step by step (float32 / float64 mix spelled out in
// oracle/tsdf_oracle.py), including its quirks: volume initialised to -trunc (:50-51), np.round
// (half-to-even) pixel rounding, colour packed as B*65536 + G*256 + R in a float32.
#include "bnv_common.cuh"

struct bnv_tsdf {
  int device;
  int32_t dim[3];
  float origin[3];
  double voxel_size, trunc;
  float* tsdf;
  float* weight;
  float* color;
  float* prior;      // scratch for bnv_tsdf_prior when the caller passes no output buffer
  int64_t n;
};

namespace bnv {

struct TsdfCam {
  float Ti[12];      // rows 0..2 of inv(T_wc), float32 like np.linalg.inv on a float32 pose
  float fx, fy, cx, cy;
  int H, W;
};

// camera-space position of voxel (x, y, z): vox2world (fusion.py:169-180: float32(origin) + float64(vs) * float32(coord)
// -> float32) followed by rigid_transform with inv(cam_pose) in float32 (fusion.py:343-348)
__device__ __forceinline__ void voxel_to_cam(int x, int y, int z, float ox, float oy, float oz, double vs, const TsdfCam& cam,
                                             float (&c)[3]) {
  const float wx = (float)((double)ox + vs * (double)(float)x);
  const float wy = (float)((double)oy + vs * (double)(float)y);
  const float wz = (float)((double)oz + vs * (double)(float)z);
#pragma unroll
  for (int r = 0; r < 3; ++r)
    c[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cam.Ti[r * 4], wx), __fmul_rn(cam.Ti[r * 4 + 1], wy)),
                               __fmul_rn(cam.Ti[r * 4 + 2], wz)),
                     cam.Ti[r * 4 + 3]);
}

// One CTA = one brick of kBx x kBy x kBz voxels (z fastest: a warp reads / writes 32 consecutive voxels).  Frustum
// culling per brick: the 8 corner voxels are projected first; if they all lie behind the camera, or all lie in front
// of it and all fall outside the same image border (with a one-pixel margin), no voxel of the (convex) brick can pass
// the reference's per-voxel test `0 <= pix < size and z > 0` (fusion.py:262-268) and the CTA exits after 8
// projections -- ~3/4 of the 206^3 volume of the headline workload.  The reference's own CUDA kernel launches a
// thread for every voxel of the volume (fusion.py:226-250).
constexpr int kBx = 2, kBy = 4, kBz = 32;

template <bool U16>
__global__ void __launch_bounds__(256) tsdf_integrate_kernel(float* __restrict__ tsdf, float* __restrict__ weight,
                                                             float* __restrict__ color, int nx, int ny, int nz,
                                                             float ox, float oy, float oz, double vs, double trunc,
                                                             TsdfCam cam, const void* __restrict__ depth_p,
                                                             const float* __restrict__ rgb, double obs) {
  const int gz = (nz + kBz - 1) / kBz, gy = (ny + kBy - 1) / kBy;
  const int bz = blockIdx.x % gz, by = (blockIdx.x / gz) % gy, bx = blockIdx.x / (gz * gy);
  __shared__ int s_out[8];
  if (threadIdx.x < 8) {
    const int cx = min(bx * kBx + ((threadIdx.x & 1) ? kBx - 1 : 0), nx - 1);
    const int cy = min(by * kBy + ((threadIdx.x & 2) ? kBy - 1 : 0), ny - 1);
    const int cz = min(bz * kBz + ((threadIdx.x & 4) ? kBz - 1 : 0), nz - 1);
    float c[3];
    voxel_to_cam(cx, cy, cz, ox, oy, oz, vs, cam, c);
    int code = 0;
    if (!(c[2] > 0.f)) {
      code = 16;                                        // behind the camera (or NaN)
    } else {
      const float u = c[0] * cam.fx / c[2] + cam.cx, v = c[1] * cam.fy / c[2] + cam.cy;
      code = (u < -1.f ? 1 : 0) | (u > (float)cam.W ? 2 : 0) | (v < -1.f ? 4 : 0) | (v > (float)cam.H ? 8 : 0);
    }
    s_out[threadIdx.x] = code;
  }
  __syncthreads();
  {
    int all_and = s_out[0], all_or = s_out[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      all_and &= s_out[j];
      all_or |= s_out[j];
    }
    // all behind, or all in front and beyond one common border
    if ((all_and & 16) || (!(all_or & 16) && (all_and & 15))) return;
  }
  const int x = bx * kBx + (threadIdx.x >> 7), y = by * kBy + ((threadIdx.x >> 5) & 3), z = bz * kBz + (threadIdx.x & 31);
  if (x >= nx || y >= ny || z >= nz) return;
  const int64_t i = ((int64_t)x * ny + y) * nz + z;
  float c[3];
  voxel_to_cam(x, y, z, ox, oy, oz, vs, cam, c);
  // cam2pix (fusion.py:182-194): float32, np.round = half-to-even
  const float fpx = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(c[0], cam.fx), c[2]), cam.cx));
  const float fpy = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(c[1], cam.fy), c[2]), cam.cy));
  if (!(fpx >= 0.f && fpx < (float)cam.W && fpy >= 0.f && fpy < (float)cam.H && c[2] > 0.f)) return;
  const int px = (int)fpx, py = (int)fpy;
  float dvf;
  if (U16) {       // load_depth (src/utils/common.py:93): uint16 mm / 1000. in float64, then .float()
    dvf = (float)((double)reinterpret_cast<const uint16_t*>(depth_p)[(size_t)py * cam.W + px] / 1000.0);
  } else {
    dvf = reinterpret_cast<const float*>(depth_p)[(size_t)py * cam.W + px];
  }
  const double dv = (double)dvf;
  const double diff = dv - (double)c[2];
  if (!(dv > 0.0 && diff >= -trunc)) return;
  const double dist = fmin(1.0, diff / trunc);
  // integrate_tsdf (fusion.py:196-206)
  const float w_old = weight[i], t_old = tsdf[i];
  const float w_new = (float)((double)w_old + obs);
  tsdf[i] = (float)(((double)__fmul_rn(w_old, t_old) + obs * dist) / (double)w_new);
  weight[i] = w_new;
  if (rgb) {       // colour running average (fusion.py:283-294), packed B*65536 + G*256 + R
    const float* p = rgb + ((size_t)py * cam.W + px) * 3;
    const float newc = floorf(p[2] * 65536.f + p[1] * 256.f + p[0]);
    const float old = color[i];
    const float ob = floorf(old / 65536.f), og = floorf((old - ob * 65536.f) / 256.f), orr = old - ob * 65536.f - og * 256.f;
    const float nb = floorf(newc / 65536.f), ng = floorf((newc - nb * 65536.f) / 256.f), nr = newc - nb * 65536.f - ng * 256.f;
    const double wn = (double)w_new;
    const double b = fmin(255.0, rint(((double)w_old * ob + obs * nb) / wn));
    const double g = fmin(255.0, rint(((double)w_old * og + obs * ng) / wn));
    const double r = fmin(255.0, rint(((double)w_old * orr + obs * nr) / wn));
    color[i] = (float)(b * 65536.0 + g * 256.0 + r);
  }
}

// prepare_tsdf_volume (src/run_e2e.py:169-186): clip(tsdf * (vs * 5), +-trunc_dist) * weight, float32
__global__ void tsdf_prior_kernel(const float* __restrict__ tsdf, int64_t n, double scale, float lim, float wgt,
                                  float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = __fmul_rn(tsdf[i], (float)scale);     // float32 array * python float: float32 multiply
  v = fminf(fmaxf(v, -lim), lim);
  out[i] = __fmul_rn(v, wgt);
}

__global__ void tsdf_fill_kernel(float* p, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace bnv

using namespace bnv;

extern "C" {

int bnv_tsdf_create(bnv_tsdf_t** out, const double* vol_bnds, double voxel_size, int device) {
  if (!out || !vol_bnds || !(voxel_size > 0)) { set_error("bnv_tsdf_create: bad argument"); return BNV_E_ARG; }
  BNV_CUDA(cudaSetDevice(device));
  bnv_tsdf* t = new bnv_tsdf();
  memset(t, 0, sizeof(*t));
  t->device = device;
  t->voxel_size = voxel_size;
  t->trunc = 5 * voxel_size;
  t->n = 1;
  for (int a = 0; a < 3; ++a) {
    t->dim[a] = (int32_t)ceil((vol_bnds[2 * a + 1] - vol_bnds[2 * a]) / voxel_size);   // fusion.py:39
    t->origin[a] = (float)vol_bnds[2 * a];
    t->n *= t->dim[a];
  }
  if (t->n <= 0 || t->n >= (1ll << 40)) { delete t; set_error("bnv_tsdf_create: bad volume size"); return BNV_E_ARG; }
  cudaError_t e = cudaMalloc((void**)&t->tsdf, t->n * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&t->weight, t->n * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&t->color, t->n * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&t->prior, t->n * 4);
  if (e != cudaSuccess) { bnv_tsdf_destroy(t); set_error("bnv_tsdf_create: cudaMalloc failed: %s", cudaGetErrorString(e)); return BNV_E_ALLOC; }
  tsdf_fill_kernel<<<(unsigned)((t->n + 255) / 256), 256>>>(t->tsdf, t->n, (float)(1.0 * 0 - t->trunc));   // fusion.py:50-51
  BNV_LAUNCH_CHECK("tsdf_fill_kernel");
  BNV_CUDA(cudaMemset(t->weight, 0, t->n * 4));
  BNV_CUDA(cudaMemset(t->color, 0, t->n * 4));
  BNV_CUDA(cudaDeviceSynchronize());
  *out = t;
  return BNV_OK;
}

int bnv_tsdf_destroy(bnv_tsdf_t* t) {
  if (!t) return BNV_OK;
  cudaSetDevice(t->device);
  if (t->tsdf) cudaFree(t->tsdf);
  if (t->weight) cudaFree(t->weight);
  if (t->color) cudaFree(t->color);
  if (t->prior) cudaFree(t->prior);
  delete t;
  return BNV_OK;
}

int bnv_tsdf_dims(const bnv_tsdf_t* t, int32_t* dims_host) {
  if (!t || !dims_host) { set_error("bnv_tsdf_dims: null argument"); return BNV_E_ARG; }
  for (int a = 0; a < 3; ++a) dims_host[a] = t->dim[a];
  return BNV_OK;
}

int bnv_tsdf_integrate(bnv_tsdf_t* t, const float* rgb_dev, const void* depth_dev, int depth_is_u16_mm, int H, int W,
                       const float* K, const float* Tinv, double obs_weight, void* stream) {
  if (!t || !depth_dev || !K || !Tinv || H <= 0 || W <= 0) { set_error("bnv_tsdf_integrate: bad argument"); return BNV_E_ARG; }
  TsdfCam cam;
  for (int i = 0; i < 12; ++i) cam.Ti[i] = Tinv[i];
  cam.fx = K[0]; cam.fy = K[4]; cam.cx = K[2]; cam.cy = K[5];
  cam.H = H; cam.W = W;
  const unsigned blocks = (unsigned)(((t->dim[0] + kBx - 1) / kBx) * ((t->dim[1] + kBy - 1) / kBy) * ((t->dim[2] + kBz - 1) / kBz));
  if (depth_is_u16_mm)
    tsdf_integrate_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(t->tsdf, t->weight, t->color, t->dim[0], t->dim[1], t->dim[2],
        t->origin[0], t->origin[1], t->origin[2], t->voxel_size, t->trunc, cam, depth_dev, rgb_dev, obs_weight);
  else
    tsdf_integrate_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(t->tsdf, t->weight, t->color, t->dim[0], t->dim[1], t->dim[2],
        t->origin[0], t->origin[1], t->origin[2], t->voxel_size, t->trunc, cam, depth_dev, rgb_dev, obs_weight);
  BNV_LAUNCH_CHECK("tsdf_integrate_kernel");
  return BNV_OK;
}

int bnv_tsdf_volume(bnv_tsdf_t* t, float** tsdf_dev, float** color_dev, float** weight_dev) {
  if (!t) { set_error("bnv_tsdf_volume: null handle"); return BNV_E_ARG; }
  if (tsdf_dev) *tsdf_dev = t->tsdf;
  if (color_dev) *color_dev = t->color;
  if (weight_dev) *weight_dev = t->weight;
  return BNV_OK;
}

int bnv_tsdf_copy(bnv_tsdf_t* t, int which, float* out_dev, void* stream) {
  if (!t || !out_dev || which < 0 || which > 2) { set_error("bnv_tsdf_copy: bad argument"); return BNV_E_ARG; }
  const float* src = which == 0 ? t->tsdf : which == 1 ? t->color : t->weight;
  BNV_CUDA(cudaMemcpyAsync(out_dev, src, t->n * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return BNV_OK;
}

int bnv_tsdf_prior(bnv_tsdf_t* t, double truncated_dist, double sdf_delta_weight, float* out_dev, void* stream) {
  if (!t) { set_error("bnv_tsdf_prior: null handle"); return BNV_E_ARG; }
  float* out = out_dev ? out_dev : t->prior;
  tsdf_prior_kernel<<<(unsigned)((t->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      t->tsdf, t->n, t->voxel_size * 5, (float)truncated_dist, (float)sdf_delta_weight, out);
  BNV_LAUNCH_CHECK("tsdf_prior_kernel");
  return BNV_OK;
}

}  // extern "C"
