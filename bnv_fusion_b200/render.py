"""calculate_loss: host-side mirror of the reference's global-optimisation loss
(/root/reference/src/utils/render_utils.py:559-594, called by NeuralMap.optimize, src/run_e2e.py:139-146) over
libbnv_b200.  Same signature and return value ({"depth_bce_loss": scalar tensor whose backward reaches
volume.features}); the ~100 small PyTorch kernels of render_with_rays + compute_sdf_loss become five launches:
ray samples -> count_optim over the samples' corners -> fused decode -> loss + d loss / d sdf -> decode backward.
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import config


class _RayLossFn(torch.autograd.Function):
    """loss as a function of the exported features; the gradient is computed in the forward pass (everything it needs
    is at hand there) and scaled by the incoming gradient in backward."""

    @staticmethod
    def forward(ctx, feats, vol, nerf, pts, weights, tsdf, dims, rays, T, n_valid, truncated_dist):
        lib = vol._lib
        n, S = pts.shape[0], pts.shape[1]
        dev = pts.device
        f = feats.detach().contiguous()
        q = pts.reshape(-1, 3)
        pred = torch.empty(n * S, dtype=torch.float32, device=dev)
        _lib.check(lib.bnv_decode_sdf(vol._handle, _lib.ptr(q), n * S, 0, _lib.ptr(f), _lib.ptr(weights), f.shape[0],
                                      nerf._mlp_handle(), int(vol.min_pts_in_grid), config.mlp_mode(), _lib.ptr(tsdf), dims,
                                      _lib.ptr(pred), None, vol._stream()), "bnv_decode_sdf")
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        grad_pred = torch.empty(n * S, dtype=torch.float32, device=dev)
        nbr, nbm = rays["neighbor_pts"], rays["neighbor_masks"]
        _lib.check(lib.bnv_ray_sdf_loss(_lib.ptr(q), _lib.ptr(pred), n, S, _lib.ptr(rays["gt_pts"]), _lib.ptr(T), _lib.ptr(nbr),
                                        _lib.ptr(nbm), int(nbr.shape[1]), _lib.ptr(rays["mask"]), _lib.ptr(n_valid),
                                        float(truncated_dist), _lib.ptr(loss), _lib.ptr(grad_pred), vol._stream()),
                   "bnv_ray_sdf_loss")
        grad = None
        if feats.requires_grad:
            grad = torch.zeros_like(f)
            _lib.check(lib.bnv_decode_sdf_backward(vol._handle, _lib.ptr(q), n * S, 0, _lib.ptr(f), _lib.ptr(weights), f.shape[0],
                                                   nerf._mlp_handle(), int(vol.min_pts_in_grid), _lib.ptr(grad_pred),
                                                   _lib.ptr(grad), vol._stream()), "bnv_decode_sdf_backward")
        ctx.grad = grad
        ctx.sdf_on_rays = pred.reshape(n, S)
        return loss[0].float()

    @staticmethod
    def backward(ctx, g):
        return (None if ctx.grad is None else ctx.grad * g,) + (None,) * 10


_CAM_CACHE = {}


def _camera_host(rays):
    """K [9] and T_wc [16] as host float32 arrays.  NeuralMap.optimize passes the same (CUDA) camera tensors with every
    ray split of a view (run_e2e.py:127-139): the device-to-host read -- a stream synchronisation -- happens once per
    tensor version, not once per call."""
    Kt, Tt = rays["intr_mat"], rays["T_wc"]
    key = (Kt.data_ptr(), Kt._version, Tt.data_ptr(), Tt._version, str(Kt.device))
    hit = _CAM_CACHE.get(key)
    if hit is None:
        K = np.ascontiguousarray(Kt.reshape(-1, 3, 3)[0].detach().cpu().numpy().astype(np.float32).reshape(9))
        T = np.ascontiguousarray(Tt.reshape(-1, 4, 4)[0].detach().cpu().numpy().astype(np.float32).reshape(16))
        if len(_CAM_CACHE) > 64:
            _CAM_CACHE.clear()
        # the tensors are kept alive with the entry, so their addresses cannot be recycled while it is cached
        hit = _CAM_CACHE[key] = (K, T, Kt, Tt)
    return hit[0], hit[1]


def sample_rays(volume, rays, truncated_units, truncated_dist, ray_max_dist, t_rand=None):
    """get_camera_params + hierarchical_sampling of render_with_rays (render_utils.py:461-492): world points on the rays
    [n, S, 3] with S = 2 * truncated_units fine + int(5 * ray_max_dist) coarse samples per ray (fine first, unsorted).
    t_rand = (fine [n, S_f], coarse [n, S_c]) uniform draws; drawn with torch.rand when None like the reference."""
    lib = volume._lib
    dev = volume.device
    uv = rays["uv"].reshape(-1, 2).to(dev).float().contiguous()
    gt = rays["gt_pts"].reshape(-1, 3).to(dev).float().contiguous()
    n = uv.shape[0]
    n_fine, n_coarse = int(truncated_units * 2), int(ray_max_dist * 5)
    if t_rand is None:
        t_rand = (torch.rand(n, n_fine, device=dev), torch.rand(n, n_coarse, device=dev))
    tf = t_rand[0].reshape(n, n_fine).to(dev).float().contiguous()
    tc = t_rand[1].reshape(n, n_coarse).to(dev).float().contiguous()
    K, T = _camera_host(rays)
    pts = torch.empty((n, n_fine + n_coarse, 3), dtype=torch.float32, device=dev)
    _lib.check(lib.bnv_ray_samples(_lib.ptr(uv), _lib.ptr(gt), n, _lib.ptr(K), _lib.ptr(T), _lib.ptr(tf), n_fine, _lib.ptr(tc),
                                   n_coarse, float(truncated_dist), _lib.ptr(pts), volume._stream()), "bnv_ray_samples")
    return pts, T


def calculate_loss(volume, rays, nerf, truncated_units, truncated_dist, ray_max_dist, sdf_delta=None, t_rand=None):
    """render_utils.calculate_loss: rays = {"uv" [1,n,2], "gt_pts" [1,n,3], "T_wc" [1,4,4], "intr_mat" [1,3,3],
    "mask" [1,n], "neighbor_pts" [1,n,Nn,3], "neighbor_masks" [1,n,Nn]} (one view per call, like the reference's
    IterableInferenceDataset).  Side effect, like the reference: volume.count_optim on the samples' corners."""
    assert volume.features is not None, "call volume.to_tensor() first."
    if rays["T_wc"].reshape(-1, 4, 4).shape[0] != 1:
        raise NotImplementedError("calculate_loss: one view (T_wc [1,4,4]) per call, as NeuralMap.optimize issues them")
    dev = volume.device
    pts, T = sample_rays(volume, rays, truncated_units, truncated_dist, ray_max_dist, t_rand)
    n, S = pts.shape[0], pts.shape[1]
    lib = volume._lib
    # count_optim BEFORE the decode, like render_with_rays (:494-498): the validity mask sees the incremented weights
    _lib.check(lib.bnv_map_count_optim_queries(volume._handle, _lib.ptr(pts), n * S, 0, _lib.ptr(volume.weights),
                                               volume.weights.shape[0], volume._stream()), "bnv_map_count_optim_queries")
    feats, weights, tsdf, dims = volume._decode_inputs(nerf, sdf_delta, True)
    r = {"gt_pts": rays["gt_pts"].reshape(-1, 3).to(dev).float().contiguous(),
         "mask": rays["mask"].reshape(-1).to(dev).float().contiguous(),
         "neighbor_pts": rays["neighbor_pts"].reshape(n, -1, 3).to(dev).float().contiguous(),
         "neighbor_masks": rays["neighbor_masks"].reshape(n, -1).to(dev).float().contiguous()}
    n_valid = (torch.sum(r["mask"]) + 1e-4).reshape(1).float()                     # render_utils.py:576
    loss = _RayLossFn.apply(volume.features, volume, nerf, pts, weights, tsdf, dims, r, T, n_valid, truncated_dist)
    return {"depth_bce_loss": loss}
