// Per-frame device helpers shared by the CUDA-core and tcgen05 encode kernels: camera model,
// float64 back-projection of one pixel, scatter of one encoded row into the per-frame accumulators.
#pragma once
#include "bnv_common.cuh"

namespace bnv {

struct Camera {
  float fx, fy, cx, cy;   // float32 intrinsics as the dataset holds them
  float T[12];            // float32 T_wc rows 0..2 (row-major 3x4)
  double max_depth;
  int H, W;
};

// uint16 millimetres / 1000. in float64, correctly rounded, without the division routine: with r = RN(1/1000) and
// q0 = RN(d r), the residual d - 1000 q0 is exact in one fma and RN(q0 + residual r) is the correctly rounded
// quotient (Markstein); checked against d / 1000.0 for every uint16 d by tests/test_div1000.py.
__device__ __forceinline__ double mm_to_m(uint16_t d) {
  const double a = (double)d;
  const double q0 = __dmul_rn(a, 0.001);
  return __fma_rn(__fma_rn(-q0, 1000.0, a), 0.001, q0);
}

// masked metric depth of a pixel with replicate padding (load_depth, src/utils/common.py:93-112)
__device__ __forceinline__ double depth_at(const uint16_t* __restrict__ d, const Camera& cam, int u, int v) {
  u = min(max(u, 0), cam.W - 1);
  v = min(max(v, 0), cam.H - 1);
  const double z = mm_to_m(__ldg(d + (size_t)v * cam.W + u));
  return (z > 0.0 && z < cam.max_depth) ? z : 0.0;
}

// Second half of one pixel of FusionInferenceAbstractDataset.__getitem__ (fusion_inference_dataset.py:52-74)
// in float64, rounded to float32 like run_e2e.py:247-249: from the 3x3 neighbourhood's camera-space points
// (kornia depth_to_3d, replicate padding) -> Sobel/8 gradients -> cross product -> normalise -> world transform
// of the centre point and the normal.  Op order == oracle/bnv_oracle.py.
__device__ __forceinline__ void backproject_finish(const double (&X)[3][3], const double (&Y)[3][3],
                                                   const double (&Z)[3][3], double zc, int u, int v, const Camera& cam,
                                                   float (&out)[6]) {
  const double e = 0.125;
  // The six products of a Sobel sum are exact (powers of two times a double far from the subnormal range), so
  // fma(e, P, a) rounds exactly like the reference's separate multiply and add: same bits, half the instructions.
  auto sobx = [&](const double (&P)[3][3]) {
    double a = __dmul_rn(-e, P[0][0]);
    a = __fma_rn(e, P[0][2], a);
    a = __fma_rn(-2 * e, P[1][0], a);
    a = __fma_rn(2 * e, P[1][2], a);
    a = __fma_rn(-e, P[2][0], a);
    a = __fma_rn(e, P[2][2], a);
    return a;
  };
  auto soby = [&](const double (&P)[3][3]) {
    double a = __dmul_rn(-e, P[0][0]);
    a = __fma_rn(-2 * e, P[0][1], a);
    a = __fma_rn(-e, P[0][2], a);
    a = __fma_rn(e, P[2][0], a);
    a = __fma_rn(2 * e, P[2][1], a);
    a = __fma_rn(e, P[2][2], a);
    return a;
  };
  const double gx0 = sobx(X), gx1 = sobx(Y), gx2 = sobx(Z);
  const double gy0 = soby(X), gy1 = soby(Y), gy2 = soby(Z);
  double n0 = __dsub_rn(__dmul_rn(gx1, gy2), __dmul_rn(gx2, gy1));
  double n1 = __dsub_rn(__dmul_rn(gx2, gy0), __dmul_rn(gx0, gy2));
  double n2 = __dsub_rn(__dmul_rn(gx0, gy1), __dmul_rn(gx1, gy0));
  const double nn = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1)), __dmul_rn(n2, n2)));
  const double den = fmax(nn, 1e-12);
  n0 = __ddiv_rn(n0, den); n1 = __ddiv_rn(n1, den); n2 = __ddiv_rn(n2, den);
  // depth2xyz (src/utils/geometry.py:150-171): (u-cx)/fx in float32, then float64 * depth
  const double xc = __dmul_rn((double)__fdiv_rn(__fsub_rn((float)u, cam.cx), cam.fx), zc);
  const double yc = __dmul_rn((double)__fdiv_rn(__fsub_rn((float)v, cam.cy), cam.fy), zc);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double t0 = (double)cam.T[r * 4 + 0], t1 = (double)cam.T[r * 4 + 1], t2 = (double)cam.T[r * 4 + 2],
                 t3 = (double)cam.T[r * 4 + 3];
    const double p = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(xc, t0), __dmul_rn(yc, t1)), __dmul_rn(zc, t2)), t3);
    const double q = __dadd_rn(__dadd_rn(__dmul_rn(n0, t0), __dmul_rn(n1, t1)), __dmul_rn(n2, t2));
    out[r] = (float)p;
    out[3 + r] = (float)q;
  }
}

// one pixel straight from global memory (bnv_backproject's dense kernel)
__device__ __forceinline__ bool backproject_pixel(const uint16_t* __restrict__ depth, const Camera& cam,
                                                  int u, int v, float (&out)[6]) {
  const double zc = depth_at(depth, cam, u, v);
  if (!(zc > 0.0)) return false;
  const double fx = (double)cam.fx, fy = (double)cam.fy, cx = (double)cam.cx, cy = (double)cam.cy;
  double X[3][3], Y[3][3], Z[3][3];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int uu = min(max(u + dx, 0), cam.W - 1), vv = min(max(v + dy, 0), cam.H - 1);
      const double z = depth_at(depth, cam, uu, vv);
      X[dy + 1][dx + 1] = __dmul_rn(__ddiv_rn(__dsub_rn((double)uu, cx), fx), z);
      Y[dy + 1][dx + 1] = __dmul_rn(__ddiv_rn(__dsub_rn((double)vv, cy), fy), z);
      Z[dy + 1][dx + 1] = z;
    }
  backproject_finish(X, Y, Z, zc, u, v, cam, out);
  return true;
}

// Tile form used by the frame prepass (bnv_encode.cu): a CTA covers a kTileW x kTileH pixel tile and stages, once
// per tile, the masked float64 depths of the tile + its one-pixel replicate-padded apron and the float64 ratios
// (u - cx) / fx, (v - cy) / fy of the apron's (clamped) columns / rows in shared memory.  Bit-identical to
// backproject_pixel: every quotient and every masked depth is produced by the same correctly-rounded operations,
// only once per tile instead of once per use (30 float64 divisions per pixel -> 3).
constexpr int kTileW = 32, kTileH = 8;
struct FrameTile {
  double z[kTileH + 2][kTileW + 2];
  double ax[kTileW + 2];
  double ay[kTileH + 2];
};

__device__ __forceinline__ void stage_frame_tile(FrameTile& t, const uint16_t* __restrict__ depth, const Camera& cam,
                                                 int u0, int v0) {
  const double fx = (double)cam.fx, fy = (double)cam.fy, cx = (double)cam.cx, cy = (double)cam.cy;
  for (int i = threadIdx.x; i < (kTileH + 2) * (kTileW + 2); i += blockDim.x) {
    const int yy = i / (kTileW + 2), xx = i - yy * (kTileW + 2);
    const int uu = min(max(u0 - 1 + xx, 0), cam.W - 1), vv = min(max(v0 - 1 + yy, 0), cam.H - 1);
    // load_depth (src/utils/common.py:93): uint16 millimetres / 1000. in float64
    const double z = mm_to_m(__ldg(depth + (size_t)vv * cam.W + uu));
    t.z[yy][xx] = (z > 0.0 && z < cam.max_depth) ? z : 0.0;
  }
  if (threadIdx.x < kTileW + 2) {
    const int uu = min(max(u0 - 1 + (int)threadIdx.x, 0), cam.W - 1);
    t.ax[threadIdx.x] = __ddiv_rn(__dsub_rn((double)uu, cx), fx);
  } else if (threadIdx.x >= 64 && threadIdx.x < 64 + kTileH + 2) {
    const int j = threadIdx.x - 64;
    const int vv = min(max(v0 - 1 + j, 0), cam.H - 1);
    t.ay[j] = __ddiv_rn(__dsub_rn((double)vv, cy), fy);
  }
}

__device__ __forceinline__ bool backproject_tile_pixel(const FrameTile& t, const Camera& cam, int tx, int ty, int u,
                                                       int v, float (&out)[6]) {
  const double zc = t.z[ty + 1][tx + 1];
  if (!(zc > 0.0)) return false;
  double X[3][3], Y[3][3], Z[3][3];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const double z = t.z[ty + dy][tx + dx];
      X[dy][dx] = __dmul_rn(t.ax[tx + dx], z);
      Y[dy][dx] = __dmul_rn(t.ay[ty + dy], z);
      Z[dy][dx] = z;
    }
  backproject_finish(X, Y, Z, zc, u, v, cam, out);
  return true;
}

struct EncSrc {
  const uint16_t* depth;   // FROM_DEPTH
  Camera cam;
  const float* pts6;       // !FROM_DEPTH
  int64_t n_points;
};

// the frames of one bnv_fuse_frames call (a kernel parameter: 1.3 KB of the 4 KB parameter space)
struct FrameBatch {
  Camera cam[kMaxBatch];
  const uint16_t* depth[kMaxBatch];
};

// Programmatic dependent launch (sm_90+): the frame's kernels are launched with programmatic stream serialization,
// so a kernel's launch and prologue overlap the tail of the kernel before it; everything that depends on that
// kernel's results comes after this wait (a no-op for a normally launched kernel).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- per-frame scratch (MapDev::ftable / fkeys / fsum) ------------------------------------------------
__device__ __forceinline__ int32_t ft_count(unsigned long long e) { return (int32_t)(e & 0xffffffffull); }
__device__ __forceinline__ int32_t ft_row(unsigned long long e) { return (int32_t)(e >> 32); }
// dense scratch row of a voxel touched this frame (valid in kernels after the prepass)
__device__ __forceinline__ int32_t scratch_row_of(const MapDev& m, int32_t flat, int frame) {
  return __ldg(reinterpret_cast<const int32_t*>(ft_entry(m, flat, frame)) + 1);
}
// point records (MapDev::prec) carry the frame's position in its batch in the top byte of their last word
__device__ __forceinline__ int record_frame(float last_word) { return (int)((uint32_t)__float_as_int(last_word) >> 24); }

// exact-parity mode: 2^30 fixed-point int64 sums (integer addition is associative => bit-reproducible means)
__device__ __forceinline__ void add_row_fixed(const MapDev& m, int32_t row, const float (&y)[8]) {
  unsigned long long* s = reinterpret_cast<unsigned long long*>(m.fsum + (size_t)row * kFeat);
#pragma unroll
  for (int j = 0; j < kFeat; ++j) atomicAdd(s + j, (unsigned long long)__double2ll_rn((double)y[j] * kFixScale));
}

// tensor-core mode: fp32 partial sums, two 16-byte vector reductions (red.global.add.v4.f32, sm_90+) per row.
// The features of this mode already carry fp16 operand rounding (1e-3), so fp32 summation-order noise (1e-7) is
// immaterial; the row count is exact in both modes (ftable's low word).
__device__ __forceinline__ void add_row_f32(const MapDev& m, int32_t row, const float (&y)[8]) {
  float* s = reinterpret_cast<float*>(m.fsum) + (size_t)row * kFeat;
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(s), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]) : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(s + 4), "f"(y[4]), "f"(y[5]), "f"(y[6]), "f"(y[7]) : "memory");
}

}  // namespace bnv
