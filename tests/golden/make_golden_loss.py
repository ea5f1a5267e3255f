"""Mint tests/golden/golden_loss.npz: the reference's own calculate_loss (/root/reference/src/utils/render_utils.py:
461-594: get_camera_params, hierarchical_sampling, count_optim, decode_pts, compute_sdf_loss) and torch autograd's
gradient of it w.r.t. volume.features -- what one inner step of NeuralMap.optimize (src/run_e2e.py:111-156) computes --
over the fakes of oracle/ref_stubs.py.  The random stratified draws (torch.rand inside stratified_sampling) are
recorded so that the B200 path can be fed the same numbers.  Build-container only."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import bnv_oracle as O     # noqa: E402
from oracle import ref_stubs as R      # noqa: E402
from bnv_fusion_b200 import synth      # noqa: E402


def make_rays(spec, fi, n_rays, rng):
    """IterableInferenceDataset._sample_key_frame (src/datasets/fusion_inference_dataset.py:357-404) on a synthetic frame"""
    import src.utils.geometry as geometry
    d, K, T = synth.make_frame(spec, fi, seed=0)
    depth = d.astype(np.float64) / 1000.0
    mask = (depth > 0) * (depth < 3.0)
    depth = depth * mask
    pts_c = geometry.depth2xyz(depth, K).reshape(-1, 3)
    pts_w = (T @ geometry.get_homogeneous(pts_c).T)[:3, :].T
    h, w = depth.shape
    idx = rng.permutation(h * w)[:n_rays]
    uv = np.mgrid[0:h, 0:w].astype(np.int32)
    uv = np.flip(uv, axis=0).copy().astype(np.float32).reshape(2, -1).T[idx]
    xyz_map = pts_w.reshape(h, w, 3)
    half = 1
    rng_ = np.arange(-half, half + 1)
    ind = np.stack(np.meshgrid(rng_, rng_), axis=-1).reshape(-1, 2)
    ind = np.tile(ind, (n_rays, 1, 1)) + uv.astype(np.int32)[:, None, :]
    ind[:, :, 0] = np.clip(ind[:, :, 0], 0, w - 1)
    ind[:, :, 1] = np.clip(ind[:, :, 1], 0, h - 1)
    nb = xyz_map[ind[:, :, 1].reshape(-1), ind[:, :, 0].reshape(-1)].reshape(n_rays, 9, 3)
    nbm = mask.astype(np.float32)[ind[:, :, 1].reshape(-1), ind[:, :, 0].reshape(-1)].reshape(n_rays, 9)
    ray_mask = mask.reshape(-1)[idx].astype(np.float32)
    ray_mask[rng.random(n_rays) < 0.1] = 0.0
    return {"uv": torch.from_numpy(uv)[None], "gt_pts": torch.from_numpy(pts_w[idx])[None].float(),
            "intr_mat": torch.from_numpy(K)[None].float(), "T_wc": torch.from_numpy(T)[None].float(),
            "mask": torch.from_numpy(ray_mask)[None], "neighbor_pts": torch.from_numpy(nb)[None].float(),
            "neighbor_masks": torch.from_numpy(nbm)[None].float(),
            "rgb": torch.zeros(1, n_rays, 3, dtype=torch.float64)}


def main():
    model, SparseVolume = R.build_reference(tempfile.mkdtemp())
    from src.utils.render_utils import calculate_loss
    g = np.load(os.path.join(HERE, "golden_parity64.npz"))
    spec = synth.stream_spec("parity64")
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device="cpu")
    w0 = g["recip/map_weights"].copy()
    w0 += 7.5
    w0[::10] -= 7.5                                       # most voxels above min_pts_in_grid = 8, every tenth below
    vol.insert(torch.from_numpy(g["recip/map_coords"]), torch.from_numpy(g["recip/map_feats"]),
               torch.from_numpy(w0), torch.from_numpy(g["recip/map_hits"]))
    vol.to_tensor()
    vol.features = torch.nn.Parameter(vol.features)
    rng = np.random.default_rng(33)
    rays = make_rays(spec, 3, 300, rng)
    delta = torch.from_numpy(g["recip/tsdf_delta"])[None, None]
    draws = []
    orig_rand = torch.rand

    def rec_rand(*a, **k):
        t = orig_rand(*a, **k)
        draws.append(t.clone())
        return t

    out = {"coords": vol.active_coordinates.numpy().copy(), "feats": vol.features.detach().numpy().copy(),
           "weights_before": vol.weights.numpy().copy(), "tsdf_delta": g["recip/tsdf_delta"]}
    torch.manual_seed(7)
    torch.rand = rec_rand
    try:
        with R.cuda_div_semantics():
            loss = calculate_loss(vol, rays, model.nerf, truncated_units=10, truncated_dist=0.05, ray_max_dist=3,
                                  sdf_delta=delta.clone())["depth_bce_loss"]
    finally:
        torch.rand = orig_rand
    loss.backward()
    assert len(draws) == 2 and draws[0].shape[-1] == 20 and draws[1].shape[-1] == 15, [d.shape for d in draws]
    out.update({k: v.numpy() for k, v in rays.items() if k != "rgb"})
    out.update({"t_rand_fine": draws[0].numpy(), "t_rand_coarse": draws[1].numpy(), "loss": np.asarray(float(loss), np.float64),
                "grad": vol.features.grad.numpy().copy(), "weights_after": vol.weights.numpy().copy()})
    # intermediate values for the oracle: points on rays and predicted SDF
    from src.utils.render_utils import render_with_rays
    draws.clear()
    torch.manual_seed(7)
    torch.rand = rec_rand
    try:
        vol2 = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device="cpu")
        vol2.insert(torch.from_numpy(g["recip/map_coords"]), torch.from_numpy(g["recip/map_feats"]),
                    torch.from_numpy(w0), torch.from_numpy(g["recip/map_hits"]))
        vol2.to_tensor()
        with torch.no_grad(), R.cuda_div_semantics():
            ro = render_with_rays(vol2, rays, model.nerf, delta.clone(), 10, 0.05, 3)
    finally:
        torch.rand = orig_rand
    out["pts_on_rays"] = ro["pts_on_rays"].numpy()
    out["sdf_on_rays"] = ro["sdf_on_rays"].numpy()
    out["ray_dirs"] = ro["ray_dirs"].numpy()
    out["cam_loc"] = ro["cam_loc"].numpy()
    np.savez_compressed(os.path.join(HERE, "golden_loss.npz"), **out)
    print("loss", float(loss), "grad nnz rows", int((np.abs(out["grad"]).sum(1) > 0).sum()), "of", len(out["grad"]),
          "weights +1:", int((out["weights_after"] != out["weights_before"]).sum()),
          "valid sdf frac", float((out["sdf_on_rays"] != np.float32(0.01)).mean()), os.path.getsize(os.path.join(HERE, "golden_loss.npz")))


if __name__ == "__main__":
    main()
