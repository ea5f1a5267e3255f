"""GPU (B200): mesh extraction kernels (bnv_mesh_count / bnv_mesh_emit, SparseVolume.extract_triangles / meshlize)
against the numpy oracle (oracle/mesh_oracle.py, pinned on analytic SDFs by tests/test_mesh_cpu.py): triangle order,
vertex ids bit-exact, positions bit-exact (same float32 sequence), and the welded mesh of a fused map is a 2-manifold
with boundary only where the map ends."""
import numpy as np
import pytest
import torch

from oracle import mesh_oracle as MO
from bnv_fusion_b200 import synth
from test_gpu_parity import _depth_to_dev, _volume, dev, model  # noqa: F401
from test_mesh_cpu import _sample_blocks

pytestmark = pytest.mark.gpu


def _vol_with_voxels(dev, ijk, dims=(0.22, 0.22, 0.22)):
    from bnv_fusion_b200.volume import SparseVolume
    vol = SparseVolume(8, 0.01, list(dims), 8, device=dev, pool_capacity=1 << 16)
    k = torch.from_numpy(ijk).to(dev)
    n = len(ijk)
    vol.insert(k, torch.zeros(n, 8, device=dev), torch.ones(n, 1, device=dev), torch.zeros(n, 1, device=dev))
    vol.to_tensor()
    return vol


def test_mesh_kernels_match_oracle_on_analytic_sdf(dev):
    n = 24
    ijk = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(3)
    ijk = ijk[rng.permutation(len(ijk))]                       # slot order != lexicographic order
    vol = _vol_with_voxels(dev, ijk)
    assert vol._n_xyz_host == (24, 24, 24)
    coords = vol.active_coordinates.cpu().numpy()
    c, r = np.array([11.3, 12.1, 11.8], np.float32), 7.4
    sdf = _sample_blocks(lambda p: np.linalg.norm(p - c, axis=-1) - r, coords)
    sdf[5] = -1.0; sdf[5, 1, 1, 1] = 0.0                        # block rule: exact zero without a sign change
    mn = vol.min_coords.cpu().numpy()
    rv, rk, _ = MO.marching_blocks(sdf, coords, vol.voxel_size, mn, vol._n_xyz_host)
    verts, faces = vol.extract_triangles(torch.from_numpy(sdf).to(dev), weld=False)
    assert verts.shape[0] == len(rv) and len(rv) > 3000
    assert np.array_equal(verts.cpu().numpy(), rv)              # same float32 op sequence: bit-exact
    assert np.array_equal(faces.cpu().numpy(), np.arange(len(rv)).reshape(-1, 3))
    wv, wf = vol.extract_triangles(torch.from_numpy(sdf).to(dev), weld=True)
    ov, of = MO.weld(rv, rk)
    assert wv.shape[0] == len(ov) and np.array_equal(wf.cpu().numpy(), of)
    assert np.array_equal(wv.cpu().numpy(), ov)
    # closed, consistently oriented sphere
    f = wf.cpu().numpy()
    half = {(a, b) for tri in f.tolist() for a, b in ((tri[0], tri[1]), (tri[1], tri[2]), (tri[2], tri[0]))}
    assert len(half) == 3 * len(f) and all((b, a) in half for a, b in half)
    assert len(ov) - 3 * len(f) // 2 + len(f) == 2
    # empty input and all-positive input
    assert vol.extract_triangles(torch.ones(len(coords), 27, device=dev))[1].shape[0] == 0


def test_meshlize_returns_a_mesh_of_the_fused_surface(model, dev, tmp_path):
    """meshlize end to end on a fused map (decode + marching cubes on the GPU): a mesh object with vertices near the
    synthetic surface, written as PLY; equals the oracle's marching cubes over the same sampled blocks."""
    from bnv_fusion_b200 import config
    config.set_mlp_mode("tc16")
    spec = synth.stream_spec("parity64")
    vol = _volume(spec, dev, pool_capacity=1 << 16)
    for fi in range(24):
        d, K, T = synth.make_frame(spec, fi, seed=0)
        model.fuse_depth_frame(vol, _depth_to_dev(d, dev), K, T, spec.max_depth)
    vol.to_tensor()
    vol.min_pts_in_grid = 1          # validity threshold on the accumulated fusion weight (rule D3): keep every voxel
    out = vol.meshlize(model.nerf, None, path=str(tmp_path / "m.ply"))
    assert out is not None
    active_pts, mesh = out
    assert active_pts.shape == (vol.active_coordinates.shape[0], 3)
    v, f = np.asarray(mesh.vertices), np.asarray(mesh.faces)
    assert len(f) > 500 and v.shape[1] == 3 and f.max() == len(v) - 1
    assert (tmp_path / "m.ply").stat().st_size > 12 * len(v)
    sdf = vol.decode_voxel_blocks(model.nerf).cpu().numpy()
    rv, rk, _ = MO.marching_blocks(sdf, vol.active_coordinates.cpu().numpy(), vol.voxel_size,
                                   vol.min_coords.cpu().numpy(), vol._n_xyz_host)
    assert np.array_equal(v.astype(np.float32), rv)
    lo, hi = vol.min_coords.cpu().numpy(), vol.max_coords.cpu().numpy()
    assert (v >= lo - 1e-6).all() and (v <= hi + 1e-6).all()
