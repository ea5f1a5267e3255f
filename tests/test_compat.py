"""The reference's import paths resolve to the B200 modules (CPU), and NeuralMap's integrate /
extract sequence (src/run_e2e.py:78-109,164-186) runs through them (GPU)."""
import os
import sys

import numpy as np
import pytest


def test_install_registers_reference_import_paths():
    import bnv_fusion_b200.compat as compat
    mods = compat.install(None)
    from src.models.fusion.local_point_fusion import LitFusionPointNet
    from src.models.sparse_volume import SparseVolume
    import third_parties.fusion as fusion
    import src.utils.voxel_utils as voxel_utils
    import bnv_fusion_b200 as b
    assert LitFusionPointNet is b.LitFusionPointNet and SparseVolume is b.SparseVolume
    assert hasattr(fusion, "TSDFVolume")
    mn, mx, n = voxel_utils.get_world_range(np.asarray([5.1] * 3), 0.01)
    assert n == [512, 512, 512]
    ijk = np.array([[3, 4, 5], [511, 0, 7]])
    flat = voxel_utils.flatten(ijk, n)
    assert np.array_equal(voxel_utils.unflatten(flat, n), ijk)
    for k in mods:
        sys.modules.pop(k, None)


@pytest.mark.gpu
def test_neural_map_sequence_through_reference_names(tcnn_params):
    """What NeuralMap.__init__/integrate/prepare_tsdf_volume/extract_mesh do, with the reference's names."""
    import torch
    import bnv_fusion_b200.compat as compat
    from bnv_fusion_b200 import synth
    compat.install(None)
    from src.models.fusion.local_point_fusion import LitFusionPointNet
    from src.models.sparse_volume import SparseVolume
    import src.utils.voxel_utils as voxel_utils
    import third_parties.fusion as fusion
    from oracle import bnv_oracle as O

    class Cfg(dict):
        __getattr__ = dict.get
    cfg = Cfg(trainer=Cfg(dense_volume=False), device_type="cuda",
              model=Cfg(feature_vector_size=8, voxel_size=0.01, min_pts_in_grid=8, tiny_cuda=True,
                        point_net=Cfg(in_channels=6), nerf=Cfg(num_encoding_fn_xyz=1, interpolate_decode=True),
                        sdf_delta_weight=0.1, ray_tracer=Cfg(truncated_units=10)))
    pointnet = LitFusionPointNet(cfg)
    pointnet.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(tcnn_params["encoder"]),
                              "nerf.model.params": torch.from_numpy(tcnn_params["decoder"])})
    pointnet.eval(); pointnet.cuda(); pointnet.freeze()
    spec = synth.stream_spec("parity64")
    dimensions = spec.dimensions
    volume = SparseVolume(cfg.model.feature_vector_size, cfg.model.voxel_size, dimensions, cfg.model.min_pts_in_grid)
    mn, mx, n = voxel_utils.get_world_range(dimensions, 0.025)
    bnds = np.zeros((3, 2)); bnds[:, 0] = mn; bnds[:, 1] = mx
    tsdf_vol = fusion.TSDFVolume(bnds, voxel_size=0.025)
    truncated_dist = min(cfg.model.ray_tracer.truncated_units * cfg.model.voxel_size * 0.5, 0.1)
    for fi in range(12):
        d, K, T = synth.make_frame(spec, fi, seed=0)
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        frame = {"input_pts": torch.from_numpy(O.backproject(depth, mask, K, T))[None].cuda().float(),
                 "rgbd": torch.from_numpy(np.concatenate([np.zeros((3,) + depth.shape), depth[None]], 0))[None].cuda().float(),
                 "intr_mat": torch.from_numpy(K)[None].cuda(), "T_wc": torch.from_numpy(T)[None].cuda()}
        with torch.no_grad():                                          # NeuralMap.integrate, run_e2e.py:78-109
            fine_feats, fine_weights, _, fine_coords, fine_n_pts = pointnet.encode_pointcloud(
                frame["input_pts"], volume.n_xyz, volume.min_coords, volume.max_coords, volume.voxel_size,
                return_dense=pointnet.dense_volume)
            assert fine_feats is not None
            volume.track_n_pts(fine_n_pts)
            pointnet._integrate(volume, fine_coords, fine_feats, fine_weights)
            rgbd = frame["rgbd"].cpu().numpy()
            rgb = (rgbd[0, :3].transpose(1, 2, 0) + 0.5) * 255.
            tsdf_vol.integrate(rgb, rgbd[0, -1], frame["intr_mat"].cpu().numpy()[0], frame["T_wc"].cpu().numpy()[0], obs_weight=1.)
    volume.to_tensor()
    tsdf_volume, _ = tsdf_vol.get_volume()                              # prepare_tsdf_volume, run_e2e.py:169-186
    tsdf_volume = tsdf_volume * (0.025 * 5)
    delta = torch.from_numpy(tsdf_volume).to(pointnet.device).float().unsqueeze(0).unsqueeze(0)
    delta = torch.clip(delta, min=-truncated_dist, max=truncated_dist) * cfg.model.sdf_delta_weight
    assert torch.allclose(delta, tsdf_vol.prior(truncated_dist, cfg.model.sdf_delta_weight), atol=1e-7)
    out = volume.meshlize(pointnet.nerf, delta)                         # extract_mesh, run_e2e.py:164-167
    assert out is not None
    volume.print_statistic()
    assert len(volume) > 1000 and volume.n_frames == 12
