"""CPU: pin oracle/bnv_oracle.py against the golden vectors minted from the reference's own
sources (tests/golden/make_golden.py).  Integer outputs bit-exact; floats within fp32 noise."""
import os

import numpy as np
import pytest

from oracle import bnv_oracle as O
from bnv_fusion_b200 import synth

FEAT_ATOL = 2e-6      # fp32 BLAS accumulation order in the reference run vs float64 here
SDF_ATOL = 1e-7


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_param_checksums(tcnn_params):
    # SURVEY.md Appendix B
    assert tcnn_params["encoder"].shape == (10240,)
    assert tcnn_params["decoder"].shape == (11264,)
    assert abs(tcnn_params["encoder"].astype(np.float64).sum() - (-349.92647505)) < 1e-6
    assert abs(tcnn_params["decoder"].astype(np.float64).sum() - (-399.78298002)) < 1e-6


def test_world_range():
    for dims, vs, n in (([0.30] * 3, 0.01, 32), ([5.1] * 3, 0.01, 512), ([5.1] * 3, 0.025, 206)):
        _, _, n_xyz = O.get_world_range(np.asarray(dims), vs)
        assert n_xyz == [n, n, n]


def test_mlp_kat(golden_dir, tcnn_params):
    g = _load(golden_dir, "golden_edge.npz")
    ye = O.mlp_forward(tcnn_params["encoder"], g["kat_enc_x"], 6, 8)
    np.testing.assert_allclose(ye, g["kat_enc_y"], atol=2e-6, rtol=0)
    xd = g["kat_dec_x"]
    x17 = np.concatenate([O.positional_encoding(xd[:, :3]), xd[:, 3:]], axis=1)
    yd = O.mlp_forward(tcnn_params["decoder"], x17, 17, 1)
    np.testing.assert_allclose(yd, g["kat_dec_y"], atol=2e-6, rtol=0)
    # published KAT of SURVEY.md Appendix B
    y = O.mlp_forward(tcnn_params["encoder"], np.array([[0.25, -0.5, 0.75, 0, 0.6, 0.8]], np.float32), 6, 8)
    np.testing.assert_allclose(y[0], [-1.11670, 1.04352, 1.34597, -1.35804, -0.26037, 1.07424, 2.57634, 2.12760], atol=2e-5)
    # fp16 operand emulation stays within the 1e-4 m SDF budget (x voxel_size 0.01)
    yd16 = O.mlp_forward(tcnn_params["decoder"], x17, 17, 1, mode="fp16")
    assert np.abs(yd16 - yd).max() * 0.01 < 1e-4


@pytest.mark.parametrize("mode", ["true", "recip"])
def test_parity64_stream(golden_dir, tcnn_params, mode):
    g = _load(golden_dir, "golden_parity64.npz")
    spec = synth.stream_spec("parity64")
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    assert tuple(g[f"{mode}/n_xyz"]) == grid.n_xyz
    assert np.array_equal(g[f"{mode}/bmin"], grid.bmin) and np.array_equal(g[f"{mode}/bmax"], grid.bmax)
    vm = O.VoxelMap(grid)
    n_frames = g["depth"].shape[0]
    for fi in range(n_frames):
        # the synthetic generator is deterministic: regenerate and compare with the stored frames
        d, K, T = synth.make_frame(spec, fi, seed=0)
        assert np.array_equal(d, g["depth"][fi]) and np.array_equal(T, g["T_wc"][fi])
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        pts6 = O.backproject(depth, mask, K, T)
        feats, counts, flat, coords, navg, _ = O.encode_pointcloud(pts6, grid, tcnn_params["encoder"], 8, mode)
        key = f"{mode}/f{fi}_"
        if key + "flat" in g.files:
            ref = g[key + "pts6"]
            assert pts6.shape == ref.shape
            # xyz: reference geometry code (pinned); normals: kornia restatement (self-consistency)
            assert np.array_equal(pts6[:, :3], ref[:, :3])
            np.testing.assert_allclose(pts6[:, 3:], ref[:, 3:], atol=1e-7, rtol=0)
            assert np.array_equal(flat, g[key + "flat"])
            assert np.array_equal(counts, g[key + "counts"])
            assert np.array_equal(coords, g[key + "coords"])
            assert np.all(np.diff(flat) > 0)
            np.testing.assert_allclose(feats, g[key + "feats"], atol=FEAT_ATOL, rtol=0)
            np.testing.assert_allclose(navg, g[key + "navg"], rtol=1e-6)
        O.integrate(vm, flat, feats, counts)
    ref_flat = O.flatten_i32(g[f"{mode}/map_coords"], grid.n_xyz)
    assert len(vm) == ref_flat.size == np.unique(ref_flat).size
    f, w, h, found = vm.query(ref_flat)
    assert found.all()
    np.testing.assert_allclose(f, g[f"{mode}/map_feats"], atol=FEAT_ATOL, rtol=0)
    np.testing.assert_allclose(w, g[f"{mode}/map_weights"][:, 0], atol=1e-6, rtol=0)
    assert np.all(h == 0) and np.all(g[f"{mode}/map_hits"] == 0)       # quirk A8
    assert (g[f"{mode}/map_weights"] >= 8).sum() > 100                 # mask branch is exercised
    dec = tcnn_params["decoder"]
    qm = g[f"{mode}/q_mesh"].reshape(-1, 3)
    sdf = O.decode_pts(vm, qm, dec, 8, None, True, mode)
    ref = g[f"{mode}/sdf_mesh"].reshape(-1)
    assert 0.05 < (ref != np.float32(0.01)).mean() < 0.95
    np.testing.assert_allclose(sdf, ref, atol=SDF_ATOL, rtol=0)
    sdf = O.decode_pts(vm, qm, dec, 8, g[f"{mode}/tsdf_delta"], True, mode)
    np.testing.assert_allclose(sdf, g[f"{mode}/sdf_mesh_prior"].reshape(-1), atol=SDF_ATOL, rtol=0)
    sdf = O.decode_pts(vm, g[f"{mode}/q_rand"], dec, 8, g[f"{mode}/tsdf_delta"], True, mode)
    np.testing.assert_allclose(sdf, g[f"{mode}/sdf_rand_prior"], atol=SDF_ATOL, rtol=0)
    sdf = O.decode_pts(vm, g[f"{mode}/q_world"], dec, 8, None, False, mode)
    np.testing.assert_allclose(sdf, g[f"{mode}/sdf_world"], atol=SDF_ATOL, rtol=0)


def test_lounge_crop(golden_dir, tcnn_params):
    g = _load(golden_dir, "golden_lounge_crop.npz")
    spec = synth.stream_spec("lounge")
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    assert grid.n_xyz == (512, 512, 512) == tuple(g["recip/n_xyz"])
    vm = O.VoxelMap(grid)
    for fi in range(2):
        depth, mask = O.load_depth_u16(g["depth"][fi], spec.max_depth)
        pts6 = O.backproject(depth, mask, g["K"][fi], g["T_wc"][fi])
        if fi == 0:
            assert np.array_equal(pts6[:, :3], g["recip/f0_pts6"][:, :3])
        feats, counts, flat, coords, navg, _ = O.encode_pointcloud(pts6, grid, tcnn_params["encoder"], 8, "recip")
        assert np.array_equal(flat, g[f"recip/f{fi}_flat"])
        assert np.array_equal(counts, g[f"recip/f{fi}_counts"])
        assert np.array_equal(coords, g[f"recip/f{fi}_coords"])
        np.testing.assert_allclose(feats, g[f"recip/f{fi}_feats"], atol=FEAT_ATOL, rtol=0)
        O.integrate(vm, flat, feats, counts)
    ref_flat = O.flatten_i32(g["recip/map_coords"], grid.n_xyz)
    f, w, h, found = vm.query(ref_flat)
    assert found.all() and len(vm) == ref_flat.size
    np.testing.assert_allclose(f, g["recip/map_feats"], atol=FEAT_ATOL, rtol=0)
    np.testing.assert_allclose(w, g["recip/map_weights"][:, 0], atol=1e-6, rtol=0)
    vm.weights *= g["recip/weight_scale"]
    for qk, sk, prior in (("q_mesh", "sdf_mesh", None), ("q_mesh", "sdf_mesh_prior", g["recip/tsdf_delta"]),
                          ("q_rand", "sdf_rand_prior", g["recip/tsdf_delta"])):
        sdf = O.decode_pts(vm, g["recip/" + qk].reshape(-1, 3), tcnn_params["decoder"], 8, prior, True, "recip")
        np.testing.assert_allclose(sdf, g["recip/" + sk].reshape(-1), atol=SDF_ATOL, rtol=0)
    assert (g["recip/sdf_mesh"] != np.float32(0.01)).mean() > 0.05


@pytest.mark.parametrize("mode", ["true", "recip"])
def test_edge_cases(golden_dir, tcnn_params, mode):
    g = _load(golden_dir, "golden_edge.npz")
    spec = synth.stream_spec("parity64")
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    feats, counts, flat, coords, navg, ex = O.encode_pointcloud(g["pts6"], grid, tcnn_params["encoder"], 8, mode)
    assert np.array_equal(flat, g[f"{mode}/flat"])
    assert np.array_equal(counts, g[f"{mode}/counts"])
    assert np.array_equal(coords, g[f"{mode}/coords"])
    np.testing.assert_allclose(feats, g[f"{mode}/feats"], atol=FEAT_ATOL, rtol=0)
    np.testing.assert_allclose(navg, g[f"{mode}/navg"], rtol=1e-6)
    # all points outside the volume -> 5 x None (rule A1)
    r = O.encode_pointcloud(g["far_pts6"], grid, tcnn_params["encoder"], 8, mode)
    assert all(v is None for v in r[:5])
    # empty input
    r = O.encode_pointcloud(np.zeros((0, 6), np.float32), grid, tcnn_params["encoder"], 8, mode)
    assert all(v is None for v in r[:5])


def test_recip_vs_true_differ_rarely():
    """Rule A2 sanity: the two division forms disagree on floor() for ~5e-6 of coordinates."""
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 5.12, 2_000_000).astype(np.float32)
    a = np.floor(O._scalar_div(x, 0.01, "recip"))
    b = np.floor(O._scalar_div(x, 0.01, "true"))
    frac = (a != b).mean()
    assert 0 < frac < 1e-4


def test_decode_backward_oracle(golden_dir, tcnn_params):
    """Gradient of decode_pts w.r.t. the voxel features vs the reference's own autograd."""
    g = _load(golden_dir, "golden_parity64.npz")
    gg = _load(golden_dir, "golden_decode_grad.npz")
    spec = synth.stream_spec("parity64")
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    vm = O.VoxelMap(grid)
    vm.insert(O.flatten_i32(g["recip/map_coords"], grid.n_xyz), g["recip/map_feats"], g["recip/map_weights"][:, 0],
              g["recip/map_hits"][:, 0])
    pos = {int(k): i for i, k in enumerate(O.flatten_i32(gg["coords"], grid.n_xyz))}
    for name in ("mesh", "rand"):
        gr = O.decode_pts_backward(vm, gg[f"{name}_q"].reshape(-1, 3), tcnn_params["decoder"], gg[f"{name}_r"].reshape(-1), 8, True)
        mine = np.zeros(gg[f"{name}_grad"].shape)
        for k, v in gr.items():
            mine[pos[k]] = v
        assert np.abs(gg[f"{name}_grad"]).max() > 1e-4
        np.testing.assert_allclose(mine, gg[f"{name}_grad"], atol=2e-8, rtol=0)
