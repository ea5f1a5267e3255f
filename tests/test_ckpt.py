"""Map checkpoint compatibility: a `*_sparse_volume.pth` written by the reference's own
SparseVolume.save (tests/golden/make_golden_ckpt.py, unmodified reference code) loads into the B200 SparseVolume and
decodes to the SDF the reference decodes from it; the B200 writer produces the same dict layout."""
import os

import numpy as np
import pytest
import torch

REF_KEYS = {"25%", "50%", "75%", "dimensions", "voxel_size", "mean", "min", "active_keys", "active_vals", "features",
            "weights", "num_hits", "active_coordinates"}


def _ref_ckpt(golden_dir):
    return torch.load(os.path.join(golden_dir, "golden_ckpt_sparse_volume.pth"), weights_only=False)


def test_reference_checkpoint_layout(golden_dir):
    ck = _ref_ckpt(golden_dir)
    assert set(ck) == REF_KEYS                                    # sparse_volume.py:846-860
    n = ck["active_coordinates"].shape[0]
    assert n > 3000 and ck["features"].shape == (n, 8) and ck["weights"].shape == (n, 1) and ck["num_hits"].shape == (n, 1)
    assert ck["active_keys"].dtype == torch.int64 and ck["active_vals"].shape == (n, 1)
    # the tensor indexer maps key -> row of the exported tensors
    assert torch.equal(ck["active_keys"][torch.argsort(ck["active_vals"][:, 0])], ck["active_coordinates"])


@pytest.mark.gpu
def test_load_reference_checkpoint_and_roundtrip(golden_dir, tcnn_params, tmp_path):
    from bnv_fusion_b200 import config, synth
    from bnv_fusion_b200.volume import SparseVolume
    from test_gpu_parity import _map_sorted
    from bnv_fusion_b200.model import LitFusionPointNet
    dev = "cuda:0"
    cfg = {"trainer": {"dense_volume": False}, "model": {"feature_vector_size": 8, "voxel_size": 0.01, "min_pts_in_grid": 8,
           "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
    model = LitFusionPointNet(cfg)
    model.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(tcnn_params["encoder"]),
                           "nerf.model.params": torch.from_numpy(tcnn_params["decoder"])})
    model.eval(); model.cuda(); model.freeze()
    spec = synth.stream_spec("parity64")
    ck = _ref_ckpt(golden_dir)
    g = np.load(os.path.join(golden_dir, "golden_ckpt.npz"))
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, pool_capacity=1 << 16)
    vol.load(os.path.join(golden_dir, "golden_ckpt_sparse_volume.pth"))
    n = ck["active_coordinates"].shape[0]
    assert len(vol) == n and vol.features.shape == (n, 8)
    flat, feats, w, h = _map_sorted(vol)
    nx = vol._n_xyz_host
    rflat = (ck["active_coordinates"][:, 0] * nx[1] * nx[2] + ck["active_coordinates"][:, 1] * nx[2] + ck["active_coordinates"][:, 2]).numpy()
    o = np.argsort(rflat)
    assert np.array_equal(flat, rflat[o]) and np.array_equal(feats, ck["features"].numpy()[o])
    assert np.array_equal(w, ck["weights"].numpy()[o, 0]) and np.array_equal(h, ck["num_hits"].numpy()[o, 0])
    q = torch.from_numpy(g["q_mesh"]).to(dev)[None]
    for mode, tol in (("fp32", 2e-6), ("tc16", 1e-4)):
        config.set_mlp_mode(mode)
        sdf = vol.decode_pts(q, model.nerf, None, is_coords=True)[0, :, :, 0].cpu().numpy()
        assert np.abs(sdf - g["sdf_mesh"]).max() <= tol, (mode, np.abs(sdf - g["sdf_mesh"]).max())
    config.set_mlp_mode("tc16")
    # the map is live after a load (the reference only rebuilds the tensor indexer): query works
    f, ww, hh = vol.query(ck["active_coordinates"][:100].to(dev))
    assert torch.equal(f.cpu(), ck["features"][:100]) and torch.equal(ww.cpu(), ck["weights"][:100])
    # our writer: same dict layout, loadable again, identical content
    vol.track_n_pts(8.5)
    vol.save(str(tmp_path / "mine"))
    mine = torch.load(str(tmp_path / "mine_sparse_volume.pth"), weights_only=False)
    assert set(mine) == REF_KEYS
    for k in ("active_keys", "active_vals", "features", "weights", "num_hits", "active_coordinates"):
        assert mine[k].dtype == ck[k].dtype and mine[k].shape == ck[k].shape, k
    assert torch.equal(mine["active_keys"][torch.argsort(mine["active_vals"][:, 0])].cpu(), mine["active_coordinates"].cpu())
    vol2 = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, pool_capacity=1 << 16)
    vol2.load(str(tmp_path / "mine_sparse_volume.pth"))
    for x, y in zip(_map_sorted(vol), _map_sorted(vol2)):
        assert np.array_equal(x, y)
    with pytest.raises(RuntimeError):                            # a checkpoint larger than the pool fails loudly
        small = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, pool_capacity=1024)
        small.load(os.path.join(golden_dir, "golden_ckpt_sparse_volume.pth"))
