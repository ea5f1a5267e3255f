"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's coarse TSDF prior in CPU mode
(third_parties/fusion.py:22-167,169-206,208-294; the path the reference takes when PyCUDA is absent,
and the CPU baseline BASELINE.json names).  Pinned against the reference's own module executed here
(tests/golden/make_golden_tsdf.py -> tests/golden/golden_tsdf.npz).

The reference mixes float32 and float64 (numba type inference on float32 arrays with Python-float
scalars); this file spells the precision of every step out, and the CUDA kernel follows the same
sequence.  The one step whose order is not fixed by the reference is the float32 BLAS product in
rigid_transform (np.dot on float32): here and on the GPU it is ((t0*x + t1*y) + t2*z) + t3.
"""
import numpy as np

F32 = np.float32


class TSDFOracle:
    def __init__(self, vol_bnds, voxel_size):
        vol_bnds = np.asarray(vol_bnds, np.float64).copy()
        self.voxel_size = float(voxel_size)
        self.trunc = 5 * self.voxel_size
        self.dim = np.ceil((vol_bnds[:, 1] - vol_bnds[:, 0]) / self.voxel_size).copy(order="C").astype(int)
        self.origin = vol_bnds[:, 0].copy(order="C").astype(F32)
        self.tsdf = np.ones(self.dim).astype(F32) * 0 - self.trunc      # fusion.py:50-51: -trunc, float64 -> array
        self.tsdf = self.tsdf.astype(F32)
        self.weight = np.zeros(self.dim, F32)
        self.color = np.zeros(self.dim, F32)
        xv, yv, zv = np.meshgrid(range(self.dim[0]), range(self.dim[1]), range(self.dim[2]), indexing="ij")
        self.vox = np.stack([xv.reshape(-1), yv.reshape(-1), zv.reshape(-1)], 1).astype(int)

    def integrate(self, color_im, depth_im, K, T_wc, obs_weight=1.0):
        h, w = depth_im.shape
        color_im = np.asarray(color_im, F32)
        color = np.floor(color_im[..., 2] * 65536 + color_im[..., 1] * 256 + color_im[..., 0])
        # vox2world: float32(origin) + float64(vs) * float32(coord) -> stored float32
        pts = (self.origin.astype(np.float64)[None] + self.voxel_size * self.vox.astype(F32).astype(np.float64)).astype(F32)
        Ti = np.linalg.inv(np.asarray(T_wc, F32)).astype(F32)
        cam = (((Ti[None, :3, 0] * pts[:, 0:1]).astype(F32) + (Ti[None, :3, 1] * pts[:, 1:2]).astype(F32)).astype(F32)
               + (Ti[None, :3, 2] * pts[:, 2:3]).astype(F32)).astype(F32) + Ti[None, :3, 3]
        cam = cam.astype(F32)
        K = np.asarray(K, F32)
        fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
        z = cam[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            px = np.rint(((cam[:, 0] * fx).astype(F32) / z).astype(F32) + cx)
            py = np.rint(((cam[:, 1] * fy).astype(F32) / z).astype(F32) + cy)
        ok = np.isfinite(px) & np.isfinite(py)
        px = np.where(ok, px, -1).astype(np.int64)
        py = np.where(ok, py, -1).astype(np.int64)
        valid = (px >= 0) & (px < w) & (py >= 0) & (py < h) & (z > 0)
        dv = np.zeros(px.shape)
        dv[valid] = depth_im[py[valid], px[valid]]
        diff = dv - z
        vp = (dv > 0) & (diff >= -self.trunc)
        dist = np.minimum(1, diff / self.trunc)
        vx, vy, vz = self.vox[vp, 0], self.vox[vp, 1], self.vox[vp, 2]
        w_old = self.weight[vx, vy, vz]
        t_old = self.tsdf[vx, vy, vz]
        w_new = (w_old.astype(np.float64) + obs_weight).astype(F32)
        t_new = (((w_old * t_old).astype(F32).astype(np.float64) + obs_weight * dist[vp]) / w_new.astype(np.float64)).astype(F32)
        self.weight[vx, vy, vz] = w_new
        self.tsdf[vx, vy, vz] = t_new
        old = self.color[vx, vy, vz]
        ob = np.floor(old / 65536)
        og = np.floor((old - ob * 65536) / 256)
        orr = old - ob * 65536 - og * 256
        new = color[py[vp], px[vp]]
        nb = np.floor(new / 65536)
        ng = np.floor((new - nb * 65536) / 256)
        nr = new - nb * 65536 - ng * 256
        nb = np.minimum(255., np.round((w_old * ob + obs_weight * nb) / w_new))
        ng = np.minimum(255., np.round((w_old * og + obs_weight * ng) / w_new))
        nr = np.minimum(255., np.round((w_old * orr + obs_weight * nr) / w_new))
        self.color[vx, vy, vz] = nb * 65536 + ng * 256 + nr

    def get_volume(self):
        return self.tsdf, self.color


def prepare_tsdf_volume(tsdf, tsdf_voxel_size, truncated_dist, sdf_delta_weight):
    """NeuralMap.prepare_tsdf_volume (src/run_e2e.py:169-186): tsdf * (vs*5) -> clip -> * weight, float32."""
    v = np.asarray(tsdf) * (tsdf_voxel_size * 5)
    v = v.astype(F32)
    v = np.clip(v, F32(-truncated_dist), F32(truncated_dist)).astype(F32)
    return (v * F32(sdf_delta_weight)).astype(F32)
