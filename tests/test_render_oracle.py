"""The ray-sampling / SDF-loss oracle (oracle/render_oracle.py) against the golden minted from the reference's own
calculate_loss + torch autograd (tests/golden/make_golden_loss.py)."""
import os

import numpy as np

from oracle import bnv_oracle as O
from oracle import render_oracle as RO


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_loss.npz"))


def test_sampling_matches_reference(golden_dir):
    g = _g(golden_dir)
    pts, dists, gt_depth, dirs, cam = RO.sample_rays(g["uv"][0], g["gt_pts"][0], g["T_wc"][0], g["intr_mat"][0],
                                                     g["t_rand_fine"][0], g["t_rand_coarse"][0], 0.05)
    assert np.array_equal(cam, g["cam_loc"][0])
    assert np.abs(dirs - g["ray_dirs"][0]).max() <= 2e-7
    assert pts.shape == g["pts_on_rays"][0].shape == (300, 35, 3)
    assert np.abs(pts - g["pts_on_rays"][0]).max() <= 1e-6                     # metres; float32 ulp at 1 m is 1.2e-7


def test_loss_and_gradient_weights_match_reference(golden_dir, tcnn_params):
    g = _g(golden_dir)
    cam = g["cam_loc"][0]
    loss, grad_out, target, valid = RO.sdf_loss(g["pts_on_rays"][0], g["sdf_on_rays"][0], g["gt_pts"][0], cam,
                                                g["neighbor_pts"][0], g["neighbor_masks"][0], g["mask"][0], 0.05)
    assert abs(loss - float(g["loss"])) <= 2e-6 * abs(float(g["loss"])), (loss, float(g["loss"]))
    assert valid.mean() > 0.3 and (grad_out != 0).mean() > 0.2
    # the decode the reference ran inside (world coordinates, prior added, count_optim BEFORE the decode)
    grid = O.Grid.from_dimensions([0.3, 0.3, 0.3], 0.01)
    vm = O.VoxelMap(grid)
    flat = O.flatten_i32(g["coords"], grid.n_xyz)
    vm.insert(flat, g["feats"], g["weights_after"][:, 0], np.zeros(len(flat), np.float32))
    pts = g["pts_on_rays"][0].reshape(-1, 3)
    sdf = O.decode_pts(vm, pts, tcnn_params["decoder"], 8, sdf_delta=g["tsdf_delta"], is_coords=False).reshape(300, 35)
    assert np.abs(sdf - g["sdf_on_rays"][0]).max() <= 2e-6
    # ... and the whole step: d loss / d features through the oracle's decode backward == torch autograd's gradient
    grads = O.decode_pts_backward(vm, pts, tcnn_params["decoder"], grad_out.reshape(-1), 8, is_coords=False)
    got = np.zeros_like(g["grad"], dtype=np.float64)
    row_of = {int(f): i for i, f in enumerate(flat)}
    for f, gv in grads.items():
        got[row_of[int(f)]] = gv
    scale = np.abs(g["grad"]).max()
    assert scale > 0 and np.abs(got - g["grad"]).max() <= 2e-5 * scale, (np.abs(got - g["grad"]).max(), scale)
    # count_optim: +1 once for every row that is a corner of some sample
    assert set(np.unique(g["weights_after"] - g["weights_before"]).tolist()) <= {0.0, 1.0}
