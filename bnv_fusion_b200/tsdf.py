"""TSDFVolume: drop-in for the reference's coarse TSDF prior (third_parties/fusion.py:19-300) with the
volume resident on the B200 (libbnv_b200 bnv_tsdf_*).  Same constructor and methods NeuralMap uses
(src/run_e2e.py:62-71,99-109,169-186): TSDFVolume(vol_bnds, voxel_size), integrate(color_im, depth_im,
cam_intr, cam_pose, obs_weight), get_volume().  integrate also accepts CUDA tensors (no host round trip);
get_volume() returns numpy arrays like the reference, get_volume_torch() / prior() stay on the device."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class TSDFVolume:
    def __init__(self, vol_bnds, voxel_size, use_gpu=True, device="cuda:0", verbose=True):
        vol_bnds = np.asarray(vol_bnds, dtype=np.float64)
        assert vol_bnds.shape == (3, 2), "[!] `vol_bnds` should be of shape (3, 2)."
        self._lib = _lib.load()
        self.device = torch.device(device)
        self._voxel_size = float(voxel_size)
        self._trunc_margin = 5 * self._voxel_size
        self._handle = C.c_void_p()
        b = np.ascontiguousarray(vol_bnds.reshape(6))
        with torch.cuda.device(self.device):
            _lib.check(self._lib.bnv_tsdf_create(C.byref(self._handle), _lib.ptr(b), self._voxel_size,
                                                 self.device.index or 0), "bnv_tsdf_create")
        dims = (C.c_int32 * 3)()
        _lib.check(self._lib.bnv_tsdf_dims(self._handle, dims), "bnv_tsdf_dims")
        self._vol_dim = np.asarray(list(dims), dtype=int)
        self._vol_bnds = vol_bnds.copy()
        self._vol_bnds[:, 1] = self._vol_bnds[:, 0] + self._vol_dim * self._voxel_size
        self._vol_origin = self._vol_bnds[:, 0].astype(np.float32)
        self.gpu_mode = 1
        if verbose:
            print("Voxel volume size: {} x {} x {} - # points: {:,}".format(*self._vol_dim, int(np.prod(self._vol_dim))))

    def __del__(self):
        try:
            h, self._handle = self._handle, None
            if h:
                self._lib.bnv_tsdf_destroy(h)
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.):
        """color_im [H,W,3] in 0..255 (or None), depth_im [H,W] metres (float) or uint16 millimetres."""
        def dev(a, dt):
            if a is None:
                return None
            t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
            return t.to(self.device, dtype=dt, non_blocking=True).contiguous()
        u16 = (torch.is_tensor(depth_im) and depth_im.dtype in (torch.uint16, torch.int16)) or \
              (isinstance(depth_im, np.ndarray) and depth_im.dtype == np.uint16)
        if u16:
            d = depth_im if torch.is_tensor(depth_im) else torch.from_numpy(depth_im.view(np.int16))
            d = d.to(self.device).contiguous()
        else:
            d = dev(depth_im, torch.float32)
        rgb = dev(color_im, torch.float32)
        H, W = d.shape
        K = np.ascontiguousarray(np.asarray(cam_intr.cpu() if torch.is_tensor(cam_intr) else cam_intr, np.float32).reshape(-1)[:9])
        pose = np.asarray(cam_pose.cpu() if torch.is_tensor(cam_pose) else cam_pose, np.float32).reshape(4, 4)
        Tinv = np.ascontiguousarray(np.linalg.inv(pose).astype(np.float32)[:3].reshape(12))    # fusion.py:254
        _lib.check(self._lib.bnv_tsdf_integrate(self._handle, _lib.ptr(rgb), _lib.ptr(d), 1 if u16 else 0, H, W,
                                                _lib.ptr(K), _lib.ptr(Tinv), float(obs_weight), self._stream()),
                   "bnv_tsdf_integrate")

    def _view(self, which):
        out = torch.empty(tuple(int(v) for v in self._vol_dim), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.bnv_tsdf_copy(self._handle, which, _lib.ptr(out), self._stream()), "bnv_tsdf_copy")
        return out

    def get_volume_torch(self):
        return self._view(0), self._view(1)

    def get_volume(self):
        t, c = self.get_volume_torch()
        return t.cpu().numpy(), c.cpu().numpy()

    def prior(self, truncated_dist, sdf_delta_weight):
        """NeuralMap.prepare_tsdf_volume (src/run_e2e.py:169-186) on the device: [1,1,Tx,Ty,Tz] float32."""
        out = torch.empty((1, 1) + tuple(int(v) for v in self._vol_dim), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.bnv_tsdf_prior(self._handle, float(truncated_dist), float(sdf_delta_weight),
                                            _lib.ptr(out), self._stream()), "bnv_tsdf_prior")
        return out
