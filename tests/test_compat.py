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
    # 12 frames of weight <= 1 never reach min_pts_in_grid = 8: every sample takes the fallback value, no block
    # changes sign and meshlize returns None exactly like the reference (sparse_volume.py:757-758) ...
    assert volume.meshlize(pointnet.nerf, delta) is None
    # ... NeuralMap.optimize's count_optim rounds raise the weights (run_e2e.py:111-162); emulate that, then extract
    volume.weights += 8.0
    out = volume.meshlize(pointnet.nerf, delta)                         # extract_mesh, run_e2e.py:164-167
    assert out is not None
    surface_pts, mesh = out
    assert len(mesh.faces) > 100 and np.asarray(mesh.vertices).shape[1] == 3
    volume.print_statistic()
    assert len(volume) > 1000 and volume.n_frames == 12


REF_RUN_E2E = "/root/reference/src/run_e2e.py"


def _self_attrs(cls):
    """attribute names a class provides: class-level members + `self.X = ...` assignments anywhere in its body"""
    import ast
    import inspect
    import textwrap
    names = set(dir(cls))
    tree = ast.parse(textwrap.dedent(inspect.getsource(cls)))
    for node in ast.walk(tree):
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "self" \
                and isinstance(node.ctx, ast.Store):
            names.add(node.attr)
    return names


@pytest.mark.skipif(not os.path.exists(REF_RUN_E2E), reason="reference checkout not present (GPU box)")
def test_every_member_run_e2e_uses_exists_on_the_shim():
    """Mechanical form of "run_e2e.py drops onto it unchanged": parse the reference's src/run_e2e.py and check every
    attribute / method / keyword argument NeuralMap and main() use on the pointnet, the volume and the TSDF volume
    against the B200 classes (names and call signatures; src/run_e2e.py:27-194,231-296)."""
    import ast
    import inspect
    from bnv_fusion_b200.model import LitFusionPointNet, tcnnNeRFModel
    from bnv_fusion_b200.volume import SparseVolume
    from bnv_fusion_b200.tsdf import TSDFVolume
    tree = ast.parse(open(REF_RUN_E2E).read())
    owners = {("self", "pointnet"): LitFusionPointNet, ("self", "volume"): SparseVolume, ("self", "tsdf_vol"): TSDFVolume,
              ("neural_map", "volume"): SparseVolume}
    roots = {"pointnet_model": LitFusionPointNet}
    used = []                                   # (class, member, call node or None)

    def owner_of(node):
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name):
            return owners.get((node.value.id, node.attr))
        if isinstance(node, ast.Name):
            return roots.get(node.id)
        if isinstance(node, ast.Attribute) and owner_of(node.value) is LitFusionPointNet and node.attr == "nerf":
            return tcnnNeRFModel
        return None

    calls = {id(n.func): n for n in ast.walk(tree) if isinstance(n, ast.Call)}
    for node in ast.walk(tree):
        if isinstance(node, ast.Attribute):
            cls = owner_of(node.value)
            if cls is not None:
                used.append((cls, node.attr, calls.get(id(node))))
    assert len(used) >= 25, len(used)
    members = {c: _self_attrs(c) for c in (LitFusionPointNet, SparseVolume, TSDFVolume, tcnnNeRFModel)}
    for cls, name, call in used:
        assert name in members[cls], f"run_e2e.py uses {cls.__name__}.{name}, which the shim lacks"
        if call is not None and hasattr(cls, name) and callable(getattr(cls, name)) and name not in ("cuda", "eval", "load_state_dict"):
            sig = inspect.signature(getattr(cls, name))
            args = [None] * (len(call.args) + 1)                       # + self
            kwargs = {k.arg: None for k in call.keywords if k.arg}
            sig.bind(*args, **kwargs)                                   # raises TypeError on an incompatible call
    # constructors as NeuralMap.__init__ / main() call them (run_e2e.py:44-48,69-71,231)
    inspect.signature(SparseVolume.__init__).bind(None, 8, 0.01, [1, 1, 1], 8)
    inspect.signature(TSDFVolume.__init__).bind(None, np.zeros((3, 2)), voxel_size=0.025)
    inspect.signature(LitFusionPointNet.__init__).bind(None, {})
    # calculate_loss(volume, rays, nerf, ...) reaches the volume through these (src/utils/render_utils.py:461-594)
    for name in ("decode_pts", "count_optim", "_query_tensor", "voxel_size", "min_coords", "max_coords", "n_xyz"):
        assert name in members[SparseVolume], name


@pytest.mark.skipif(not os.path.exists("/root/reference/src/utils/voxel_utils.py"), reason="reference checkout not present")
def test_compat_voxel_utils_carries_the_reference_helpers():
    """With a reference root, src.utils.voxel_utils also offers the reference's remaining helpers (reference modules
    import them); the three hot-path functions stay the B200 ones."""
    pytest.importorskip("kornia")            # the reference file imports src.utils.geometry, which imports kornia
    import bnv_fusion_b200.compat as compat
    mods = compat.install("/root/reference")
    import src.utils.voxel_utils as vu
    from bnv_fusion_b200.volume import get_world_range
    assert vu.get_world_range is get_world_range
    import ast
    ref_funcs = [n.name for n in ast.parse(open("/root/reference/src/utils/voxel_utils.py").read()).body
                 if isinstance(n, ast.FunctionDef)]
    assert len(ref_funcs) > 3
    for name in ref_funcs:
        assert hasattr(vu, name), name
    for k in list(mods) + ["src", "src.models", "src.models.fusion", "src.utils", "third_parties"]:
        sys.modules.pop(k, None)
    if "/root/reference" in sys.path:
        sys.path.remove("/root/reference")
