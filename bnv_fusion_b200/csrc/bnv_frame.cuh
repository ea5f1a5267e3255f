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

// masked metric depth of a pixel with replicate padding (load_depth, src/utils/common.py:93-112)
__device__ __forceinline__ double depth_at(const uint16_t* __restrict__ d, const Camera& cam, int u, int v) {
  u = min(max(u, 0), cam.W - 1);
  v = min(max(v, 0), cam.H - 1);
  const double z = (double)__ldg(d + (size_t)v * cam.W + u) / 1000.0;
  return (z > 0.0 && z < cam.max_depth) ? z : 0.0;
}

// One pixel of FusionInferenceAbstractDataset.__getitem__ (fusion_inference_dataset.py:52-74) in
// float64, rounded to float32 like run_e2e.py:247-249.  Op order == oracle/bnv_oracle.py.
__device__ __forceinline__ bool backproject_pixel(const uint16_t* __restrict__ depth, const Camera& cam,
                                                  int u, int v, float (&out)[6]) {
  const double zc = depth_at(depth, cam, u, v);
  if (!(zc > 0.0)) return false;
  const double fx = (double)cam.fx, fy = (double)cam.fy, cx = (double)cam.cx, cy = (double)cam.cy;
  // kornia depth_to_3d over the 3x3 neighbourhood (float64 (u-cx)/fx), Sobel/8, replicate pad
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
  const double e = 0.125;
  auto sobx = [&](double (&P)[3][3]) {
    double a = __dmul_rn(-e, P[0][0]);
    a = __dadd_rn(a, __dmul_rn(e, P[0][2]));
    a = __dadd_rn(a, __dmul_rn(-2 * e, P[1][0]));
    a = __dadd_rn(a, __dmul_rn(2 * e, P[1][2]));
    a = __dadd_rn(a, __dmul_rn(-e, P[2][0]));
    a = __dadd_rn(a, __dmul_rn(e, P[2][2]));
    return a;
  };
  auto soby = [&](double (&P)[3][3]) {
    double a = __dmul_rn(-e, P[0][0]);
    a = __dadd_rn(a, __dmul_rn(-2 * e, P[0][1]));
    a = __dadd_rn(a, __dmul_rn(-e, P[0][2]));
    a = __dadd_rn(a, __dmul_rn(e, P[2][0]));
    a = __dadd_rn(a, __dmul_rn(2 * e, P[2][1]));
    a = __dadd_rn(a, __dmul_rn(e, P[2][2]));
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
  return true;
}

// Same arithmetic with the three kinds of float64 divisions hoisted out of the per-pixel path (bit-identical:
// every quotient is produced by the same correctly-rounded division, only once instead of per use):
//   zlut[d]  = (double)d / 1000.0            for every uint16 depth value   (device table, built at map creation)
//   ax[u]    = ((double)u - cx) / fx         for every image column         (per CTA, shared memory)
//   ay[v]    = ((double)v - cy) / fy         for every image row
// 30 float64 divisions per pixel -> 3 (the normal's normalisation).
__device__ __forceinline__ void build_ratio_tables(const Camera& cam, double* __restrict__ ax, double* __restrict__ ay) {
  const double fx = (double)cam.fx, fy = (double)cam.fy, cx = (double)cam.cx, cy = (double)cam.cy;
  for (int i = threadIdx.x; i < cam.W; i += blockDim.x) ax[i] = __ddiv_rn(__dsub_rn((double)i, cx), fx);
  for (int i = threadIdx.x; i < cam.H; i += blockDim.x) ay[i] = __ddiv_rn(__dsub_rn((double)i, cy), fy);
}

// the nine raw uint16 depths of the 3x3 neighbourhood (replicate padding), row-major; issued as
// independent loads so that a caller can prefetch them a tile ahead
__device__ __forceinline__ void load_depth9(const uint16_t* __restrict__ depth, const Camera& cam, int u, int v,
                                            uint32_t (&raw)[9]) {
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int uu = min(max(u + dx, 0), cam.W - 1), vv = min(max(v + dy, 0), cam.H - 1);
      raw[(dy + 1) * 3 + dx + 1] = __ldg(depth + (size_t)vv * cam.W + uu);
    }
}

__device__ __forceinline__ bool backproject_raw_lut(const uint32_t (&raw)[9], const Camera& cam,
                                                    const double* __restrict__ zlut, const double* __restrict__ ax,
                                                    const double* __restrict__ ay, int u, int v, float (&out)[6]) {
  auto z_of = [&](uint32_t d) {
    const double z = __ldg(zlut + d);
    return (z > 0.0 && z < cam.max_depth) ? z : 0.0;
  };
  const double zc = z_of(raw[4]);
  if (!(zc > 0.0)) return false;
  double X[3][3], Y[3][3], Z[3][3];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int uu = min(max(u + dx, 0), cam.W - 1), vv = min(max(v + dy, 0), cam.H - 1);
      const double z = (dx == 0 && dy == 0) ? zc : z_of(raw[(dy + 1) * 3 + dx + 1]);
      X[dy + 1][dx + 1] = __dmul_rn(ax[uu], z);
      Y[dy + 1][dx + 1] = __dmul_rn(ay[vv], z);
      Z[dy + 1][dx + 1] = z;
    }
  const double e = 0.125;
  auto sobx = [&](double (&P)[3][3]) {
    double a = __dmul_rn(-e, P[0][0]);
    a = __dadd_rn(a, __dmul_rn(e, P[0][2]));
    a = __dadd_rn(a, __dmul_rn(-2 * e, P[1][0]));
    a = __dadd_rn(a, __dmul_rn(2 * e, P[1][2]));
    a = __dadd_rn(a, __dmul_rn(-e, P[2][0]));
    a = __dadd_rn(a, __dmul_rn(e, P[2][2]));
    return a;
  };
  auto soby = [&](double (&P)[3][3]) {
    double a = __dmul_rn(-e, P[0][0]);
    a = __dadd_rn(a, __dmul_rn(-2 * e, P[0][1]));
    a = __dadd_rn(a, __dmul_rn(-e, P[0][2]));
    a = __dadd_rn(a, __dmul_rn(e, P[2][0]));
    a = __dadd_rn(a, __dmul_rn(2 * e, P[2][1]));
    a = __dadd_rn(a, __dmul_rn(e, P[2][2]));
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
  return true;
}

__device__ __forceinline__ bool backproject_pixel_lut(const uint16_t* __restrict__ depth, const Camera& cam,
                                                      const double* __restrict__ zlut, const double* __restrict__ ax,
                                                      const double* __restrict__ ay, int u, int v, float (&out)[6]) {
  uint32_t raw[9];
  load_depth9(depth, cam, u, v, raw);
  return backproject_raw_lut(raw, cam, zlut, ax, ay, u, v, out);
}

struct EncSrc {
  const uint16_t* depth;   // FROM_DEPTH
  Camera cam;
  const float* pts6;       // !FROM_DEPTH
  int64_t n_points;
  const double* zlut;      // FROM_DEPTH: millimetres -> metres table (bnv_map::zlut)
};

// split form of scatter_row: claim the voxel's scratch row first (the CAS round trip can then overlap
// other work), add the encoded features later (fire-and-forget reductions)
__device__ __forceinline__ int32_t claim_row(const MapDev& m, int32_t flat, int32_t row) {
  const int32_t old = atomicCAS(&m.ftable[flat], kEmpty, row);
  if (old == kEmpty) {
    m.fkeys[row] = flat;
    const int32_t pos = atomicAdd(&m.ctr[1], 1);
    m.touched[pos] = row;
    return row;
  }
  return old;
}
__device__ __forceinline__ void add_row(const MapDev& m, int32_t slot, const float (&y)[8]) {
  atomicAdd(&m.fcnt[slot], 1);
  unsigned long long* s = reinterpret_cast<unsigned long long*>(m.fsum + (size_t)slot * kFeat);
#pragma unroll
  for (int j = 0; j < kFeat; ++j) atomicAdd(s + j, (unsigned long long)__double2ll_rn((double)y[j] * kFixScale));
}

// tensor-core mode: fp32 partial sums in the first 32 bytes of the scratch row, accumulated with two
// 16-byte vector reductions (red.global.add.v4.f32, sm_90+) + the count: 3 L2 operations per row instead
// of 9.  The encode kernel is bound by the number of L2 atomic operations (tools/encode_experiment.py);
// the features of this mode already carry fp16 operand rounding (1e-3), so fp32 summation-order noise
// (1e-7) is immaterial.  The exact-parity CUDA-core mode keeps the order-independent fixed-point sums.
__device__ __forceinline__ void add_row_f32(const MapDev& m, int32_t slot, const float (&y)[8]) {
  float* s = reinterpret_cast<float*>(m.fsum + (size_t)slot * kFeat);
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(s), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]) : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(s + 4), "f"(y[4]), "f"(y[5]), "f"(y[6]), "f"(y[7]) : "memory");
  atomicAdd(&m.fcnt[slot], 1);
}

// Warp-shuffle reduction per voxel before the L2 reductions: neighbouring pixels (lanes) mostly fall into
// the same voxel, so runs of equal scratch rows are summed with a segmented inclusive scan (5 shuffle
// steps) and only the last lane of a run issues the reductions, with the run length as the count.
// Must be called by all 32 lanes; slot < 0 = nothing to add for this lane.
__device__ __forceinline__ void add_row_f32_runs(const MapDev& m, int32_t slot, float (&y)[8]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int32_t prev = __shfl_up_sync(full, slot, 1);
  const unsigned heads = __ballot_sync(full, lane == 0 || prev != slot);
  const int my_head = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const bool take = lane - d >= my_head;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = __shfl_up_sync(full, y[j], d);
      if (take) y[j] += v;
    }
  }
  const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
  if (tail && slot >= 0) {
    float* s = reinterpret_cast<float*>(m.fsum + (size_t)slot * kFeat);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(s), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(s + 4), "f"(y[4]), "f"(y[5]), "f"(y[6]), "f"(y[7]) : "memory");
    atomicAdd(&m.fcnt[slot], lane - my_head + 1);
  }
}

__device__ __forceinline__ void scatter_row(const MapDev& m, int32_t flat, int32_t row, const float (&y)[8]) {
  const int32_t old = atomicCAS(&m.ftable[flat], kEmpty, row);
  const int32_t slot = old == kEmpty ? row : old;
  if (old == kEmpty) {
    m.fkeys[row] = flat;
    const int32_t pos = atomicAdd(&m.ctr[1], 1);
    m.touched[pos] = row;
  }
  atomicAdd(&m.fcnt[slot], 1);
  unsigned long long* s = reinterpret_cast<unsigned long long*>(m.fsum + (size_t)slot * kFeat);
#pragma unroll
  for (int j = 0; j < kFeat; ++j) {
    const long long q = __double2ll_rn((double)y[j] * kFixScale);
    atomicAdd(s + j, (unsigned long long)q);
  }
}

}  // namespace bnv
