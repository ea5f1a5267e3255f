"""LitFusionPointNet: host-side mirror of the reference's fusion model API
(/root/reference/src/models/fusion/local_point_fusion.py:21-165,647-673) with the tcnn encoder
(src/utils/pointnet_utils.py:269-294) and decoder (src/models/fusion/modules.py:136-285) modules.

Keeps the reference's constructor (a Hydra-style cfg), `load_state_dict` keys
(`pointnet_backbone.model.params` [10240], `nerf.model.params` [11264]), `forward`,
`encode_pointcloud`, `get_relative_xyz`, `_update`, `_integrate`, `.nerf.{xyz_encoding, geo_forward,
get_neighbors}`, `.eval()/.cuda()/.freeze()/.device/.dense_volume`, so src/run_e2e.py runs on it
unchanged.  All arithmetic of the hot path happens in libbnv_b200 (CUDA, sm_100a); this module is
plumbing.  Additional fused entry points (`fuse_depth_frame`, `fuse_points`) do the whole of
NeuralMap.integrate's local-fusion half in two kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import config
from .volume import SparseVolume, _REGISTRY, geometry_key

WIDTH = 64


def _get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


class _TcnnNetwork(nn.Module):
    """Parameter holder with tinycudann's module layout (`.params`, flat float32)."""

    def __init__(self, n_in, n_out):
        super().__init__()
        self.n_in, self.n_out = n_in, n_out
        in_pad = (n_in + 15) // 16 * 16
        out_pad = (n_out + 15) // 16 * 16
        n = WIDTH * in_pad + 2 * WIDTH * WIDTH + out_pad * WIDTH
        # tcnn initialises uniformly (Xavier-like); real runs load a checkpoint
        g = torch.Generator().manual_seed(1337 + n_in)
        self.params = nn.Parameter((torch.rand(n, generator=g) * 2 - 1) * (6.0 / (WIDTH + WIDTH)) ** 0.5)
        self._handle = None
        self._handle_key = None

    def handle(self):
        p = self.params
        if not p.is_cuda:
            raise RuntimeError("BNV-Fusion B200 modules run on CUDA only (call .cuda(); no CPU fallback)")
        key = (p.data_ptr(), p._version, p.device.index)
        if self._handle is None or key != self._handle_key:
            lib = _lib.load()
            if self._handle is not None:
                lib.bnv_mlp_destroy(self._handle)
            host = p.detach().float().cpu().contiguous().numpy()
            h = C.c_void_p()
            with torch.cuda.device(p.device):
                _lib.check(lib.bnv_mlp_create(C.byref(h), _lib.ptr(host), host.size, self.n_in, self.n_out,
                                              p.device.index or 0), "bnv_mlp_create")
            self._handle, self._handle_key = h, key
        return self._handle

    def forward(self, x):
        """[n, n_in] -> [n, n_out] (fp32; the reference's tcnn returns fp16)."""
        x = x.detach().reshape(-1, self.n_in).float().contiguous()
        y = torch.empty((x.shape[0], self.n_out), dtype=torch.float32, device=x.device)
        s = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(_lib.load().bnv_mlp_forward(self.handle(), _lib.ptr(x), x.shape[0], _lib.ptr(y),
                                               config.mlp_mode(), s), "bnv_mlp_forward")
        return y

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().bnv_mlp_destroy(self._handle)
        except Exception:
            pass


class tcnnPointNetEncoder(nn.Module):
    """src/utils/pointnet_utils.py:269-294."""

    def __init__(self, feat_dims, in_channels, **kwargs):
        super().__init__()
        self.model = _TcnnNetwork(in_channels, feat_dims)
        self.feat_dims = feat_dims
        self.in_channels = in_channels

    def _mlp_handle(self):
        return self.model.handle()

    def forward(self, x, global_feat):
        x = x.transpose(2, 1)
        B, N, D = x.size()
        x = self.model(x.reshape(-1, self.in_channels))
        x = x.reshape(B, N, self.feat_dims).permute(0, 2, 1)
        if global_feat:
            return torch.mean(x, 2, keepdim=True).view(-1, self.feat_dims)
        return x


def positional_encoding(tensor, num_encoding_functions=1):
    """src/models/fusion/modules.py:81-123 (include_input=True, log_sampling=True)."""
    enc = [tensor]
    for i in range(num_encoding_functions):
        f = 2.0 ** i
        enc += [torch.sin(tensor * f), torch.cos(tensor * f)]
    return enc[0] if len(enc) == 1 else torch.cat(enc, dim=-1)


class tcnnNeRFModel(nn.Module):
    """src/models/fusion/modules.py:136-285 (decoder)."""

    def __init__(self, feat_dims, hidden_size=256, num_layers=4, num_encoding_fn_xyz=1, num_encoding_fn_dir=4,
                 include_input_xyz=True, include_input_dir=True, xyz_agnostic=False, interpolate_decode=True,
                 global_coords=False, **kwargs):
        super().__init__()
        self.dim_xyz = (3 if include_input_xyz else 0) + 2 * 3 * num_encoding_fn_xyz
        self.num_encoding_fn_xyz = num_encoding_fn_xyz
        self.interpolate_decode = interpolate_decode
        self.global_coords = global_coords
        if self.dim_xyz + feat_dims != 17:
            raise NotImplementedError("the B200 decoder kernels are built for the configured 17 -> 1 network "
                                      "(num_encoding_fn_xyz=1, feature_vector_size=8)")
        self.model = _TcnnNetwork(self.dim_xyz + feat_dims, 1)

    def _mlp_handle(self):
        return self.model.handle()

    def xyz_encoding(self, x):
        return positional_encoding(x, self.num_encoding_fn_xyz)

    def geo_forward(self, xyz):
        shapes = list(xyz.shape)
        out = self.model(xyz.reshape(-1, shapes[-1]))
        return out.reshape(shapes[:-1] + [1])

    def get_neighbors(self, points):
        """[b, n_steps, n_samples, 3] -> [b, 8, n_steps, n_samples, 3] int32 (modules.py:178-247)."""
        fl, ce = torch.floor(points), torch.ceil(points)
        sel = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
        out = [torch.stack([(ce if s[a] else fl)[..., a] for a in range(3)], dim=-1) for s in sel]
        return torch.stack(out, dim=1).int()

    def forward(self, x, feats, mask=None, test=False):
        xyz = x[..., :3]
        enc = self.xyz_encoding(xyz)
        if not test:
            feats = feats.unsqueeze(1).repeat(1, xyz.shape[1], 1)
        return self.geo_forward(torch.cat([enc, feats], dim=-1))


class LitFusionPointNet(nn.Module):
    """src/models/fusion/local_point_fusion.py:21-49 (tiny_cuda configuration)."""

    def __init__(self, cfg, **kwargs):
        super().__init__()
        self.cfg = cfg
        model = _get(cfg, "model")
        self.dense_volume = _get(_get(cfg, "trainer"), "dense_volume", False)
        if self.dense_volume:
            raise NotImplementedError("trainer.dense_volume=True (dense training grids) is outside the "
                                      "per-frame hot path")
        self.feat_dims = _get(model, "feature_vector_size", 8)
        nerf_cfg = dict(_get(model, "nerf", {}) or {})
        self.interpolate_decode = nerf_cfg.get("interpolate_decode", True)
        pn = dict(_get(model, "point_net", {"in_channels": 6}) or {})
        self.pointnet_backbone = tcnnPointNetEncoder(self.feat_dims, **pn)
        self.nerf = tcnnNeRFModel(self.feat_dims, **nerf_cfg)
        self.voxel_size = _get(model, "voxel_size", 0.01)
        self.min_pts_in_grid = _get(model, "min_pts_in_grid", 8)
        self._scratch = {}

    # ---- Lightning-module surface used by run_e2e.py:231-236 -----------------------------------
    @property
    def device(self):
        return self.pointnet_backbone.model.params.device

    def freeze(self):
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    # ---- forward (local_point_fusion.py:51-65) ---------------------------------------------------
    def forward(self, input_feats, normalize, voxel_size=None, global_feats=True):
        if normalize:
            inv = float(np.float32(1.0) / np.float32(voxel_size))    # tensor / python scalar on CUDA
            input_feats[:, :, :3] = input_feats[:, :, :3] * inv
            assert torch.min(input_feats[:, :, :3]) >= -1
            assert torch.max(input_feats[:, :, :3]) <= 1
        return self.pointnet_backbone(input_feats.permute(0, 2, 1), global_feats)

    def get_relative_xyz(self, xyz, bound_min, voxel_size):
        """local_point_fusion.py:153-165 (torch ops; the fused kernels do not call this)."""
        inv = float(np.float32(1.0) / np.float32(voxel_size))
        xyz_normalized = (xyz - bound_min) * inv
        grid_id = self.nerf.get_neighbors(xyz_normalized.unsqueeze(1)).squeeze(2)
        rel = (xyz_normalized.unsqueeze(1) - grid_id) * voxel_size
        return rel, grid_id

    # ---- encode ------------------------------------------------------------------------------------
    def _volume_for(self, n_xyz, bound_min, voxel_size):
        key = geometry_key(n_xyz, bound_min, voxel_size) + (self.device.index or 0,)
        vol = _REGISTRY.get(key)
        if vol is None:
            vol = self._scratch.get(key)
        if vol is None:
            # no SparseVolume with this geometry exists yet: build one to hold the per-frame scratch
            n = [int(v) for v in n_xyz]
            dims = (np.asarray(n, np.float64) - 2) * float(voxel_size)
            vol = SparseVolume(self.feat_dims, float(voxel_size), dims, self.min_pts_in_grid,
                               device=str(self.device), pool_capacity=1024)
            if geometry_key(vol._n_xyz_host, vol.min_coords, voxel_size) + (vol._dev_index,) != key:
                raise RuntimeError("encode_pointcloud: no SparseVolume matches the given grid geometry")
            self._scratch[key] = vol
        return vol

    def encode_pointcloud(self, input_pts, n_xyz, bound_min, bound_max, voxel_size, return_dense=True):
        """local_point_fusion.py:81-151 with return_dense=False semantics:
        (feats [M,8] f32, counts [M,1] i64, flat_ids [M] i64 ascending, coords [M,3] i64, n_avg)."""
        if return_dense:
            raise NotImplementedError("return_dense=True (dense training grids) is outside the hot path")
        assert input_pts.shape[0] == 1 and input_pts.shape[-1] == 6
        vol = self._volume_for(n_xyz, bound_min, voxel_size)
        pts = input_pts[0].detach().float().contiguous()
        n = pts.shape[0]
        dev = pts.device
        cap = min(8 * max(n, 1), int(np.prod(vol._n_xyz_host)))
        feats = torch.empty((cap, self.feat_dims), dtype=torch.float32, device=dev)
        counts = torch.empty(cap, dtype=torch.int64, device=dev)
        flat = torch.empty(cap, dtype=torch.int64, device=dev)
        coords = torch.empty((cap, 3), dtype=torch.int64, device=dev)
        stats = torch.zeros(2, dtype=torch.int64, device=dev)
        navg = torch.zeros(1, dtype=torch.float32, device=dev)
        _lib.check(vol._lib.bnv_encode_points(vol._handle, _lib.ptr(pts), n, self.pointnet_backbone._mlp_handle(),
                                              int(self.min_pts_in_grid), config.mlp_mode(), _lib.ptr(feats),
                                              _lib.ptr(counts), _lib.ptr(flat), _lib.ptr(coords), cap,
                                              _lib.ptr(stats), _lib.ptr(navg), vol._stream()), "bnv_encode_points")
        M, Mt = [int(v) for v in stats.tolist()]
        if Mt == 0:
            return None, None, None, None, None           # local_point_fusion.py:101-102
        return feats[:M], counts[:M].unsqueeze(-1), flat[:M], coords[:M], navg[0]

    # ---- integrate (local_point_fusion.py:647-673) ----------------------------------------------
    def _update(self, new_feats, new_weights, old_feats, old_weights):
        updated_weights = old_weights + new_weights
        new_feats = (old_feats * old_weights + new_feats * new_weights) / updated_weights
        return new_feats, updated_weights

    def _integrate(self, volume_object, fine_coords, fine_feats, fine_weights):
        if fine_coords is None or len(fine_coords) == 0:
            return
        coords = fine_coords.reshape(-1, 3).long().contiguous()
        n = coords.shape[0]
        feats = fine_feats.detach().reshape(n, self.feat_dims).float().contiguous()
        counts = fine_weights.reshape(n).long().contiguous()
        _lib.check(volume_object._lib.bnv_integrate(volume_object._handle, _lib.ptr(coords), _lib.ptr(feats),
                                                    _lib.ptr(counts), n, volume_object._stream()), "bnv_integrate")

    # ---- fused fast paths (no reference equivalent: the whole local-fusion step in 2 kernels) ----
    def fuse_depth_frame(self, volume, depth_mm, K, T_wc, max_depth=3.0, stats=None, navg=None):
        """depth_mm: uint16 [H,W] CUDA tensor (millimetres); K [3,3], T_wc [4,4] host float32."""
        assert depth_mm.dtype == torch.uint16 or depth_mm.dtype == torch.int16
        H, W = depth_mm.shape
        K = np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
        T = np.ascontiguousarray(np.asarray(T_wc, np.float32).reshape(16))
        _lib.check(volume._lib.bnv_fuse_frame(volume._handle, _lib.ptr(depth_mm), H, W, _lib.ptr(K), _lib.ptr(T),
                                              float(max_depth), self.pointnet_backbone._mlp_handle(),
                                              int(self.min_pts_in_grid), config.mlp_mode(), _lib.ptr(stats),
                                              _lib.ptr(navg), volume._stream()), "bnv_fuse_frame")

    def fuse_depth_frame_host(self, volume, depth_mm_host, K, T_wc, max_depth=3.0, stats_host=None, next_depth_mm_host=None):
        """Host-buffer form of `fuse_depth_frame`: depth_mm_host is a (pinned) CPU uint16/int16 [H,W] tensor,
        stats_host a (pinned) CPU int64[4] tensor or None.  H2D copy, fusion and the D2H of the frame statistics
        are enqueued on the current stream by ONE library call; synchronise the stream before reading stats_host.
        next_depth_mm_host (optional, pinned) is the frame the following call will pass: its copy is started on
        the map's copy stream and overlaps this frame's kernels."""
        assert not depth_mm_host.is_cuda and depth_mm_host.dtype in (torch.uint16, torch.int16)
        H, W = depth_mm_host.shape
        K = np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
        T = np.ascontiguousarray(np.asarray(T_wc, np.float32).reshape(16))
        _lib.check(volume._lib.bnv_fuse_frame_host(volume._handle, C.c_void_p(depth_mm_host.data_ptr()), H, W,
                                                   _lib.ptr(K), _lib.ptr(T), float(max_depth),
                                                   self.pointnet_backbone._mlp_handle(), int(self.min_pts_in_grid),
                                                   config.mlp_mode(),
                                                   C.c_void_p(stats_host.data_ptr()) if stats_host is not None else None,
                                                   C.c_void_p(next_depth_mm_host.data_ptr()) if next_depth_mm_host is not None else None,
                                                   volume._stream()), "bnv_fuse_frame_host")

    # ---- frame batches: n frames through ONE pass of the three kernels (include/bnv_b200.h "frame batches") ----------
    @staticmethod
    def _batch_cameras(n, Ks, Ts):
        K = np.asarray(Ks, np.float32)
        K = np.broadcast_to(K.reshape(-1, 9), (n, 9)) if K.size == 9 else K.reshape(n, 9)
        T = np.asarray(Ts, np.float32).reshape(n, 16)
        return np.ascontiguousarray(K), np.ascontiguousarray(T)

    def fuse_depth_frames(self, volume, depths_mm, Ks, Ts_wc, max_depth=3.0, stats=None, navg=None):
        """Fuse a batch of depth frames: the map ends up exactly as after `fuse_depth_frame` on each of them in order
        (the per-voxel running averages are applied in frame order by one thread per voxel).  depths_mm: sequence of
        uint16 [H,W] CUDA tensors or one [n,H,W] tensor; Ks [3,3] (shared) or [n,3,3]; Ts_wc [n,4,4].  The volume must
        have been laid out for batches: `SparseVolume(..., frame_batch=n)` or `volume.set_frame_batch(n)`.
        stats (optional CUDA int64[4]) receives the frame statistics summed over the batch."""
        frames = list(depths_mm)
        n = len(frames)
        H, W = frames[0].shape
        for d in frames:
            assert d.is_cuda and d.dtype in (torch.uint16, torch.int16) and tuple(d.shape) == (H, W) and d.is_contiguous()
        K, T = self._batch_cameras(n, Ks, Ts_wc)
        ptrs = (C.c_void_p * n)(*[d.data_ptr() for d in frames])
        _lib.check(volume._lib.bnv_fuse_frames(volume._handle, ptrs, n, H, W, _lib.ptr(K), _lib.ptr(T), float(max_depth),
                                               self.pointnet_backbone._mlp_handle(), int(self.min_pts_in_grid),
                                               config.mlp_mode(), _lib.ptr(stats), _lib.ptr(navg), volume._stream()),
                   "bnv_fuse_frames")

    def fuse_depth_frames_host(self, volume, depths_mm_host, Ks, Ts_wc, max_depth=3.0, stats_host=None,
                               next_depths_mm_host=None):
        """Host-buffer form of `fuse_depth_frames` (pinned CPU uint16/int16 [H,W] tensors), with the prefetch hint of
        `fuse_depth_frame_host` for the frames of the next call."""
        frames = list(depths_mm_host)
        n = len(frames)
        H, W = frames[0].shape
        for d in frames:
            assert not d.is_cuda and d.dtype in (torch.uint16, torch.int16) and tuple(d.shape) == (H, W) and d.is_contiguous()
        K, T = self._batch_cameras(n, Ks, Ts_wc)
        ptrs = (C.c_void_p * n)(*[d.data_ptr() for d in frames])
        nxt = list(next_depths_mm_host) if next_depths_mm_host is not None else []
        nptrs = (C.c_void_p * max(1, len(nxt)))(*[d.data_ptr() for d in nxt])
        _lib.check(volume._lib.bnv_fuse_frames_host(volume._handle, ptrs, n, H, W, _lib.ptr(K), _lib.ptr(T), float(max_depth),
                                                    self.pointnet_backbone._mlp_handle(), int(self.min_pts_in_grid),
                                                    config.mlp_mode(),
                                                    C.c_void_p(stats_host.data_ptr()) if stats_host is not None else None,
                                                    nptrs if nxt else None, len(nxt), volume._stream()),
                   "bnv_fuse_frames_host")

    def fuse_points(self, volume, input_pts, stats=None, navg=None):
        pts = input_pts.reshape(-1, 6).detach().float().contiguous()
        _lib.check(volume._lib.bnv_fuse_points(volume._handle, _lib.ptr(pts), pts.shape[0],
                                               self.pointnet_backbone._mlp_handle(), int(self.min_pts_in_grid),
                                               config.mlp_mode(), _lib.ptr(stats), _lib.ptr(navg),
                                               volume._stream()), "bnv_fuse_points")


def backproject(volume, depth_mm, K, T_wc, max_depth=3.0):
    """Dataset-side arithmetic of FusionInferenceAbstractDataset.__getitem__
    (src/datasets/fusion_inference_dataset.py:52-74) on the device: uint16 depth [H,W] ->
    input_pts [N,6] float32 in row-major pixel order."""
    H, W = depth_mm.shape
    K = np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
    T = np.ascontiguousarray(np.asarray(T_wc, np.float32).reshape(16))
    pts = torch.empty((H * W, 6), dtype=torch.float32, device=depth_mm.device)
    n = torch.zeros(1, dtype=torch.int32, device=depth_mm.device)
    _lib.check(volume._lib.bnv_backproject(volume._handle, _lib.ptr(depth_mm), H, W, _lib.ptr(K), _lib.ptr(T),
                                           float(max_depth), _lib.ptr(pts), _lib.ptr(n), volume._stream()),
               "bnv_backproject")
    return pts[: int(n.item())]
