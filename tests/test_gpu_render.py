"""GPU (B200): one inner step of the global optimisation (calculate_loss) through libbnv_b200 against
the golden minted from the reference's own calculate_loss + torch autograd (tests/golden/make_golden_loss.py): sampled
points, count_optim side effect, loss value and the gradient w.r.t. volume.features."""
import os

import numpy as np
import pytest
import torch

from bnv_fusion_b200 import config, synth
from test_gpu_parity import dev, model  # noqa: F401

pytestmark = pytest.mark.gpu


def _setup(golden_dir, dev):
    from bnv_fusion_b200.volume import SparseVolume
    g = np.load(os.path.join(golden_dir, "golden_loss.npz"))
    spec = synth.stream_spec("parity64")
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, pool_capacity=1 << 16)
    n = len(g["coords"])
    vol.insert(torch.from_numpy(g["coords"]).to(dev), torch.from_numpy(g["feats"]).to(dev),
               torch.from_numpy(g["weights_before"]).to(dev), torch.zeros(n, 1, device=dev))
    vol.to_tensor()
    nx = vol._n_xyz_host
    flat = lambda c: c[:, 0] * nx[1] * nx[2] + c[:, 1] * nx[2] + c[:, 2]
    order = np.argsort(flat(g["coords"]))
    rows = torch.argsort(flat(vol.active_coordinates)).cpu().numpy()          # exported row of the i-th sorted key
    perm = np.empty(n, np.int64)
    perm[rows] = order                                                          # exported row r holds golden row perm[r]
    rays = {k: torch.from_numpy(g[k]).to(dev) for k in ("uv", "gt_pts", "intr_mat", "T_wc", "mask", "neighbor_pts", "neighbor_masks")}
    t_rand = (torch.from_numpy(g["t_rand_fine"][0]).to(dev), torch.from_numpy(g["t_rand_coarse"][0]).to(dev))
    delta = torch.from_numpy(g["tsdf_delta"]).to(dev)[None, None]
    return g, vol, perm, rays, t_rand, delta


def test_sample_rays_matches_reference(golden_dir, dev):
    from bnv_fusion_b200.render import sample_rays
    g, vol, perm, rays, t_rand, delta = _setup(golden_dir, dev)
    pts, _ = sample_rays(vol, rays, 10, 0.05, 3, t_rand)
    p = pts.cpu().numpy()
    assert p.shape == (300, 35, 3)
    d = np.linalg.norm(p - g["cam_loc"][0][None, None], axis=-1)
    p = np.take_along_axis(p, np.argsort(d, axis=1)[:, :, None], axis=1)       # the reference sorts by distance
    assert np.abs(p - g["pts_on_rays"][0]).max() <= 1e-6


@pytest.mark.parametrize("mode", ["fp32", "tc16"])
def test_calculate_loss_matches_reference_autograd(golden_dir, model, dev, mode):
    from bnv_fusion_b200.render import calculate_loss
    config.set_mlp_mode(mode)
    g, vol, perm, rays, t_rand, delta = _setup(golden_dir, dev)
    vol.features = torch.nn.Parameter(vol.features)
    out = calculate_loss(vol, rays, model.nerf, truncated_units=10, truncated_dist=0.05, ray_max_dist=3, sdf_delta=delta,
                         t_rand=t_rand)
    loss = out["depth_bce_loss"]
    loss.backward()
    ref_loss, ref_grad = float(g["loss"]), g["grad"][perm]
    tol = 1e-5 if mode == "fp32" else 2e-3
    assert abs(float(loss) - ref_loss) <= tol * ref_loss, (float(loss), ref_loss)
    # count_optim side effect: the same rows gained +1
    assert np.array_equal(vol.weights.cpu().numpy(), g["weights_after"][perm])
    grad = vol.features.grad.cpu().numpy()
    scale = np.abs(ref_grad).max()
    # a sample whose |pred - target| is at float noise level may flip its L1 sign: allow a handful of rows
    bad = np.abs(grad - ref_grad).max(1) > (1e-4 if mode == "fp32" else 2e-2) * scale
    assert bad.mean() <= 0.002, (bad.sum(), np.abs(grad - ref_grad).max(), scale)
    assert (np.abs(grad).sum(1) > 0).sum() > 1000
    # a small step against the gradient lowers the loss of the same draws (weights reset first: count_optim is a side
    # effect of every call and would change the validity masks)
    with torch.no_grad():
        vol.features -= 2e-3 * vol.features.grad / vol.features.grad.abs().max()
        vol.weights.copy_(torch.from_numpy(g["weights_before"][perm]).to(dev))
    vol.features.grad = None
    again = calculate_loss(vol, rays, model.nerf, 10, 0.05, 3, sdf_delta=delta, t_rand=t_rand)["depth_bce_loss"]
    assert float(again) < float(loss), (float(again), float(loss))
    config.set_mlp_mode("tc16")
