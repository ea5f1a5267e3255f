"""GPU (B200) parity tests: the CUDA path through the C ABI vs the golden vectors minted from the
reference's own sources, and vs the numpy oracle on seeded inputs.

Bars (BASELINE.json north_star): voxel ids / keys / counts / order bit-exact; SDF <= 1e-4 abs.
In BNV_MLP_FP32 mode the features and SDFs additionally agree with the fp32 oracle to ~1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import bnv_oracle as O
from bnv_fusion_b200 import synth

pytestmark = pytest.mark.gpu

FEAT_ATOL = 2e-5
SDF_ATOL = 1e-4          # the contract
SDF_ATOL_FP32 = 2e-6     # what the fp32 path actually achieves


@pytest.fixture(autouse=True)
def fp32_mode():
    """This file pins the exact-parity CUDA-core arithmetic; test_gpu_tc.py covers the tensor cores."""
    from bnv_fusion_b200 import config
    config.set_mlp_mode("fp32")
    yield
    config.set_mlp_mode("tc16")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return "cuda:0"


@pytest.fixture(scope="module")
def model(dev, tcnn_params):
    from bnv_fusion_b200.model import LitFusionPointNet
    cfg = {"trainer": {"dense_volume": False},
           "model": {"feature_vector_size": 8, "voxel_size": 0.01, "min_pts_in_grid": 8,
                     "point_net": {"in_channels": 6},
                     "nerf": {"hidden_size": 256, "num_layers": 4, "num_encoding_fn_xyz": 1,
                              "num_encoding_fn_dir": 6, "include_input_xyz": True, "include_input_dir": True,
                              "interpolate_decode": True, "global_coords": False, "xyz_agnostic": False}}}
    m = LitFusionPointNet(cfg)
    sd = {"pointnet_backbone.model.params": torch.from_numpy(tcnn_params["encoder"]),
          "nerf.model.params": torch.from_numpy(tcnn_params["decoder"])}
    r = m.load_state_dict(sd)
    assert not r.missing_keys and not r.unexpected_keys
    m.eval()
    m.cuda()
    m.freeze()
    return m


def _volume(spec, dev, **kw):
    from bnv_fusion_b200.volume import SparseVolume
    return SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, **kw)


def _depth_to_dev(d, dev):
    return torch.from_numpy(d.view(np.int16)).to(dev).view(torch.uint16)


def _map_sorted(vol):
    coords, feats, weights, hits = vol.to_tensor()
    n = vol._n_xyz_host
    flat = (coords[:, 0] * (n[1] * n[2]) + coords[:, 1] * n[2] + coords[:, 2]).cpu().numpy()
    o = np.argsort(flat, kind="stable")
    return flat[o], feats.cpu().numpy()[o], weights.cpu().numpy()[o, 0], hits.cpu().numpy()[o, 0]


def _golden_map_sorted(g, mode, n_xyz):
    flat = O.flatten_i32(g[f"{mode}/map_coords"], n_xyz)
    o = np.argsort(flat, kind="stable")
    return flat[o], g[f"{mode}/map_feats"][o], g[f"{mode}/map_weights"][o, 0], g[f"{mode}/map_hits"][o, 0]


def test_pytorch_cuda_true_div_fast_path(dev):
    """Rule A2: on CUDA `tensor / python_scalar` is x * (1.0f/(float)s)."""
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(4_000_000, generator=g) * 5.12).to(dev)
    a = x / 0.01
    b = x * float(np.float32(1.0) / np.float32(0.01))
    assert torch.equal(a, b)
    c = (x.cpu() / 0.01).to(dev)
    assert (a != c).float().mean().item() > 0      # and it differs from the CPU's IEEE divide


def test_mlp_known_answers(model, golden_dir, dev):
    g = np.load(os.path.join(golden_dir, "golden_edge.npz"))
    ye = model.pointnet_backbone.model(torch.from_numpy(g["kat_enc_x"]).to(dev)).cpu().numpy()
    np.testing.assert_allclose(ye, g["kat_enc_y"], atol=FEAT_ATOL, rtol=0)
    xd = torch.from_numpy(g["kat_dec_x"]).to(dev)
    geo_in = torch.cat([model.nerf.xyz_encoding(xd[:, :3]), xd[:, 3:]], dim=-1)
    yd = model.nerf.geo_forward(geo_in).cpu().numpy()
    np.testing.assert_allclose(yd, g["kat_dec_y"], atol=FEAT_ATOL, rtol=0)
    # the module-level forward of the reference API ([B,N,6] -> [B,F,N])
    x = torch.from_numpy(g["kat_enc_x"]).to(dev)[None].clone()
    x[:, :, :3] *= 0.01
    out = model(x, normalize=True, voxel_size=0.01, global_feats=False)
    assert out.shape == (1, 8, 64)
    np.testing.assert_allclose(out[0].t().cpu().numpy(), g["kat_enc_y"], atol=1e-4, rtol=0)


@pytest.mark.parametrize("workload,frames", [("parity64", [0, 3, 23]), ("lounge", [0, 7])])
def test_backproject_bit_exact(workload, frames, dev):
    from bnv_fusion_b200.model import backproject
    spec = synth.stream_spec(workload)
    vol = _volume(spec, dev, pool_capacity=1024)
    for fi in frames:
        d, K, T = synth.make_frame(spec, fi, seed=0)
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        ref = O.backproject(depth, mask, K, T)
        got = backproject(vol, _depth_to_dev(d, dev), K, T, spec.max_depth).cpu().numpy()
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), np.abs(got - ref).max()


@pytest.mark.parametrize("mode", ["recip"])
def test_encode_points_vs_golden(model, golden_dir, dev, mode):
    g = np.load(os.path.join(golden_dir, "golden_parity64.npz"))
    spec = synth.stream_spec("parity64")
    vol = _volume(spec, dev, pool_capacity=1 << 16)
    for fi in (0, 3, 23):
        pts6 = torch.from_numpy(g[f"{mode}/f{fi}_pts6"]).to(dev)[None]
        feats, counts, flat, coords, navg = model.encode_pointcloud(
            pts6, vol.n_xyz, vol.min_coords, vol.max_coords, vol.voxel_size, return_dense=False)
        assert feats.dtype == torch.float32 and counts.dtype == torch.int64 and flat.dtype == torch.int64
        assert counts.shape == (flat.shape[0], 1) and coords.shape == (flat.shape[0], 3)
        assert np.array_equal(flat.cpu().numpy(), g[f"{mode}/f{fi}_flat"])
        assert np.array_equal(counts.cpu().numpy(), g[f"{mode}/f{fi}_counts"])
        assert np.array_equal(coords.cpu().numpy(), g[f"{mode}/f{fi}_coords"])
        np.testing.assert_allclose(feats.cpu().numpy(), g[f"{mode}/f{fi}_feats"], atol=FEAT_ATOL, rtol=0)
        np.testing.assert_allclose(float(navg), float(g[f"{mode}/f{fi}_navg"]), rtol=1e-6)
    vol.check_status()
    assert len(vol) == 0            # encode alone never touches the persistent map


def test_edge_cases(model, golden_dir, dev):
    g = np.load(os.path.join(golden_dir, "golden_edge.npz"))
    spec = synth.stream_spec("parity64")
    vol = _volume(spec, dev, pool_capacity=1 << 16)
    args = (vol.n_xyz, vol.min_coords, vol.max_coords, vol.voxel_size)
    feats, counts, flat, coords, navg = model.encode_pointcloud(torch.from_numpy(g["pts6"]).to(dev)[None], *args,
                                                                return_dense=False)
    assert np.array_equal(flat.cpu().numpy(), g["recip/flat"])        # integral coords: duplicates counted
    assert np.array_equal(counts.cpu().numpy(), g["recip/counts"])
    assert np.array_equal(coords.cpu().numpy(), g["recip/coords"])
    np.testing.assert_allclose(feats.cpu().numpy(), g["recip/feats"], atol=FEAT_ATOL, rtol=0)
    np.testing.assert_allclose(float(navg), float(g["recip/navg"]), rtol=1e-6)
    # everything outside the volume -> 5 x None (rule A1); and again a normal frame afterwards
    r = model.encode_pointcloud(torch.from_numpy(g["far_pts6"]).to(dev)[None], *args, return_dense=False)
    assert all(v is None for v in r)
    r = model.encode_pointcloud(torch.zeros((1, 0, 6), device=dev), *args, return_dense=False)
    assert all(v is None for v in r)
    feats2, counts2, flat2, _, _ = model.encode_pointcloud(torch.from_numpy(g["pts6"]).to(dev)[None], *args,
                                                           return_dense=False)
    assert torch.equal(flat2, flat) and torch.equal(counts2, counts) and torch.equal(feats2, feats)
    # empty integrate / insert are no-ops
    model._integrate(vol, coords[:0], feats[:0], counts[:0])
    assert vol.insert(coords[:0], feats[:0], counts[:0].float(), counts[:0].float()) is None
    assert len(vol) == 0
    vol.check_status()


@pytest.mark.parametrize("path", ["api", "fused_points", "fused_depth"])
def test_parity64_stream(model, golden_dir, dev, path):
    """24 frames: encode -> integrate -> to_tensor -> decode, three ways through the C ABI."""
    g = np.load(os.path.join(golden_dir, "golden_parity64.npz"))
    spec = synth.stream_spec("parity64")
    vol = _volume(spec, dev, pool_capacity=1 << 16)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    for fi in range(g["depth"].shape[0]):
        d, K, T = g["depth"][fi], g["K"][fi], g["T_wc"][fi]
        if path == "fused_depth":
            model.fuse_depth_frame(vol, _depth_to_dev(d, dev), K, T, spec.max_depth, stats=stats)
            continue
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        pts6 = torch.from_numpy(O.backproject(depth, mask, K, T)).to(dev)
        if path == "fused_points":
            model.fuse_points(vol, pts6, stats=stats)
            continue
        feats, counts, flat, coords, navg = model.encode_pointcloud(
            pts6[None], vol.n_xyz, vol.min_coords, vol.max_coords, vol.voxel_size, return_dense=False)
        vol.track_n_pts(navg)
        model._integrate(vol, coords, feats, counts)
    vol.check_status()
    flat, feats, w, h = _map_sorted(vol)
    rflat, rfeats, rw, rh = _golden_map_sorted(g, "recip", vol._n_xyz_host)
    assert np.array_equal(flat, rflat)                         # same set of voxels, bit-exact keys
    np.testing.assert_allclose(w, rw, atol=1e-6, rtol=0)
    np.testing.assert_allclose(feats, rfeats, atol=FEAT_ATOL, rtol=0)
    assert np.all(h == 0)                                       # quirk A8: num_hits never incremented
    if path != "api":
        s = stats.cpu().numpy()
        assert s[0] > 3000 and s[1] % 8 == 0 and 0 < s[1] <= 8 * s[0] and 0 < s[3] <= s[2]
    # decode: meshlize samples and random points, with and without the TSDF prior
    nerf = model.nerf
    prior = torch.from_numpy(g["recip/tsdf_delta"]).to(dev)[None, None]
    qm = torch.from_numpy(g["recip/q_mesh"]).to(dev)[None]
    sdf = vol.decode_pts(qm, nerf, None, is_coords=True)
    assert sdf.shape == (1, qm.shape[1], 27, 1)
    ref = g["recip/sdf_mesh"]
    err = np.abs(sdf[0, :, :, 0].cpu().numpy() - ref)
    assert err.max() <= SDF_ATOL_FP32, err.max()
    assert ((sdf[0, :, :, 0].cpu().numpy() == np.float32(0.01)) == (ref == np.float32(0.01))).all()
    sdf = vol.decode_pts(qm, nerf, prior, is_coords=True)
    np.testing.assert_allclose(sdf[0, :, :, 0].cpu().numpy(), g["recip/sdf_mesh_prior"], atol=SDF_ATOL_FP32, rtol=0)
    qr = torch.from_numpy(g["recip/q_rand"]).to(dev)[None, :, None, :]
    sdf = vol.decode_pts(qr, nerf, prior, is_coords=True)
    np.testing.assert_allclose(sdf[0, :, 0, 0].cpu().numpy(), g["recip/sdf_rand_prior"], atol=SDF_ATOL_FP32, rtol=0)
    qw = torch.from_numpy(g["recip/q_world"]).to(dev)[None, :, None, :]
    sdf = vol.decode_pts(qw, nerf, None, is_coords=False)
    np.testing.assert_allclose(sdf[0, :, 0, 0].cpu().numpy(), g["recip/sdf_world"], atol=SDF_ATOL_FP32, rtol=0)


def test_lounge_crop_512_grid(model, golden_dir, dev):
    g = np.load(os.path.join(golden_dir, "golden_lounge_crop.npz"))
    spec = synth.stream_spec("lounge")
    vol = _volume(spec, dev, pool_capacity=1 << 18)
    assert vol._n_xyz_host == (512, 512, 512)
    for fi in range(2):
        model.fuse_depth_frame(vol, _depth_to_dev(g["depth"][fi], dev), g["K"][fi], g["T_wc"][fi], spec.max_depth)
    vol.check_status()
    flat, feats, w, h = _map_sorted(vol)
    rflat, rfeats, rw, rh = _golden_map_sorted(g, "recip", vol._n_xyz_host)
    assert np.array_equal(flat, rflat)
    np.testing.assert_allclose(w, rw, atol=1e-6, rtol=0)
    np.testing.assert_allclose(feats, rfeats, atol=FEAT_ATOL, rtol=0)
    vol.to_tensor()
    vol.weights *= float(g["recip/weight_scale"])
    prior = torch.from_numpy(g["recip/tsdf_delta"]).to(dev)[None, None]
    for qk, sk, p in (("q_mesh", "sdf_mesh", None), ("q_mesh", "sdf_mesh_prior", prior), ("q_rand", "sdf_rand_prior", prior)):
        q = torch.from_numpy(g["recip/" + qk].reshape(-1, 3)).to(dev)[None, :, None, :]
        sdf = vol.decode_pts(q, model.nerf, p, is_coords=True)[0, :, 0, 0].cpu().numpy()
        np.testing.assert_allclose(sdf, g["recip/" + sk].reshape(-1), atol=SDF_ATOL_FP32, rtol=0)


def test_map_insert_query_roundtrip(dev):
    spec = synth.stream_spec("lounge")
    vol = _volume(spec, dev, pool_capacity=1 << 20)
    rng = np.random.default_rng(5)
    n = 300_000
    flat = rng.choice(512 ** 3, size=n, replace=False)
    keys = torch.from_numpy(np.stack([flat // (512 * 512), (flat // 512) % 512, flat % 512], 1)).to(dev)
    feats = torch.randn(n, 8, device=dev)
    w = torch.rand(n, 1, device=dev) * 20
    h = torch.rand(n, 1, device=dev)
    vol.insert(keys, feats, w, h)
    assert len(vol) == n
    f2, w2, h2 = vol.query(keys)
    assert torch.equal(f2, feats) and torch.equal(w2, w) and torch.equal(h2, h)
    # upsert overwrites, misses give zeros, shapes follow the keys' leading dims
    vol.insert(keys[:1000], feats[:1000] * 2, w[:1000] + 1, h[:1000])
    assert len(vol) == n
    f3, w3, _ = vol.query(keys[:1000].reshape(10, 100, 3))
    assert f3.shape == (10, 100, 8) and torch.equal(f3.reshape(-1, 8), feats[:1000] * 2)
    assert torch.equal(w3.reshape(-1, 1), w[:1000] + 1)
    miss = torch.tensor([[0, 0, 0], [511, 511, 511], [-1, 5, 5], [512, 0, 0]], device=dev)
    present = set(flat[np.isin(flat, [0, 512 ** 3 - 1])].tolist())
    fm, wm, hm = vol.query(miss)
    if not present:
        assert float(fm.abs().sum()) == 0 and float(wm.abs().sum()) == 0
    assert float(fm[2:].abs().sum()) == 0
    # to_tensor: a set-equal snapshot whose row r is slot r
    coords, ft, wt, ht = vol.to_tensor()
    assert coords.shape == (n, 3)
    cf = (coords[:, 0] * 512 * 512 + coords[:, 1] * 512 + coords[:, 2]).cpu().numpy()
    assert np.array_equal(np.sort(cf), np.sort(flat))
    f4, w4, h4 = vol.query(coords)
    assert torch.equal(f4, ft) and torch.equal(w4, wt)
    fq, wq, hq = vol._query_tensor(coords[:5000].reshape(1, 50, 100, 3))
    assert torch.equal(fq.reshape(-1, 8), ft[:5000])
    # out-of-grid insert is latched as an error
    vol.insert(miss[2:3], feats[:1], w[:1], h[:1])
    with pytest.raises(RuntimeError):
        vol.check_status()
    with pytest.raises(RuntimeError):       # every host-side reader fails loudly once a fault is latched
        vol.to_tensor()


def test_full_frame_paths_agree_and_properties(model, dev):
    """640x480 into the 512^3 grid (BASELINE.json configs[1] shape): size-independent properties.
    The fused depth path, the fused points path and the reference-shaped API path must leave
    bit-identical maps; per-voxel means are order-independent (fixed-point accumulation), so the
    result is deterministic across repeats."""
    from bnv_fusion_b200.model import backproject
    spec = synth.stream_spec("lounge")
    vols = [_volume(spec, dev, pool_capacity=1 << 21) for _ in range(3)]
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    for fi in range(3):
        d, K, T = synth.make_frame(spec, fi, seed=1)
        dd = _depth_to_dev(d, dev)
        model.fuse_depth_frame(vols[0], dd, K, T, spec.max_depth, stats=stats)
        pts = backproject(vols[1], dd, K, T, spec.max_depth)
        assert pts.shape[0] == int(stats[0])
        model.fuse_points(vols[1], pts)
        feats, counts, flat, coords, navg = model.encode_pointcloud(
            pts[None], vols[2].n_xyz, vols[2].min_coords, vols[2].max_coords, vols[2].voxel_size, return_dense=False)
        assert bool((flat[1:] > flat[:-1]).all())                 # ascending, unique
        assert int(counts.min()) >= 8
        assert int(stats[3]) == flat.shape[0]
        model._integrate(vols[2], coords, feats, counts)
    maps = [_map_sorted(v) for v in vols]
    for m in maps[1:]:
        assert np.array_equal(m[0], maps[0][0])
        assert np.array_equal(m[1], maps[0][1])
        assert np.array_equal(m[2], maps[0][2])
    for v in vols:
        v.check_status()
    # decode properties on the real map: masked queries return voxel_size exactly; duplicates
    # from integral coordinates are normalised away (weights sum to 1)
    vol = vols[0]
    coords, feats, weights, _ = vol.to_tensor()
    vol.weights += 8.0                                             # make every voxel "valid"
    blocks = vol.decode_voxel_blocks(model.nerf)
    assert blocks.shape == (coords.shape[0], 3, 3, 3)
    q = O.meshlize_samples(coords[:2000].cpu().numpy())
    ref = vol.decode_pts(torch.from_numpy(q).to(dev)[None], model.nerf, None, is_coords=True)
    assert torch.equal(ref[0, :, :, 0], blocks[:2000].reshape(2000, 27))
    centre = blocks[:, 1, 1, 1]
    assert torch.isfinite(blocks).all() and float(centre.abs().max()) < 0.05
    far = torch.full((1, 10, 1, 3), 3.25, device=dev)
    out, mask = vol.decode_pts(far, model.nerf, None, is_coords=True, return_mask=True)
    assert not mask.any() and torch.all(out == np.float32(0.01))


def test_count_optim_and_query_tensor(model, dev):
    """SparseVolume.count_optim (sparse_volume.py:602-622): weights[rows(keys)] += 1 once per distinct row."""
    spec = synth.stream_spec("parity64")
    vol = _volume(spec, dev, pool_capacity=1 << 16)
    for fi in range(3):
        d, K, T = synth.make_frame(spec, fi, seed=0)
        model.fuse_depth_frame(vol, _depth_to_dev(d, dev), K, T, spec.max_depth)
    coords, feats, weights, hits = vol.to_tensor()
    w0 = weights.clone()
    n = coords.shape[0]
    # keys shaped like decode_pts' neighbour tensor [1, 8, B, S, 3]; rows 0..99 three times, plus misses
    sel = torch.arange(100, device=dev)
    keys = torch.cat([coords[sel], coords[sel], coords[sel], torch.full((40, 3), 31, device=dev)]).float()
    keys = keys.reshape(1, 1, -1, 1, 3)
    vol.count_optim(keys)
    expect = w0.clone()
    expect[sel] += 1                                            # non-accumulating for duplicates
    hit31 = ((coords == 31).all(1)).nonzero().flatten()
    expect[hit31] += 1
    assert torch.equal(vol.weights, expect)
    f, w, h = vol._query_tensor(coords[:50].reshape(1, 50, 1, 3))
    assert torch.equal(w.reshape(-1), vol.weights[:50, 0]) and torch.equal(f.reshape(-1, 8), feats[:50])


@pytest.mark.parametrize("mode", ["fp32", "tc16"])
def test_decode_backward_vs_reference_autograd(model, golden_dir, dev, mode):
    """d decode_pts / d volume.features (NeuralMap.optimize) vs the gradients the reference's own code +
    torch autograd produce (tests/golden/make_golden_grad.py)."""
    from bnv_fusion_b200 import config
    config.set_mlp_mode(mode)
    g = np.load(os.path.join(golden_dir, "golden_parity64.npz"))
    gg = np.load(os.path.join(golden_dir, "golden_decode_grad.npz"))
    spec = synth.stream_spec("parity64")
    vol = _volume(spec, dev, pool_capacity=1 << 16)
    vol.insert(torch.from_numpy(g["recip/map_coords"]).to(dev), torch.from_numpy(g["recip/map_feats"]).to(dev),
               torch.from_numpy(g["recip/map_weights"]).to(dev), torch.from_numpy(g["recip/map_hits"]).to(dev))
    coords, _, _, _ = vol.to_tensor()
    n = vol._n_xyz_host
    flat = (coords[:, 0] * n[1] * n[2] + coords[:, 1] * n[2] + coords[:, 2]).cpu().numpy()
    ref_flat = O.flatten_i32(gg["coords"], n)
    perm = np.array([dict(zip(ref_flat.tolist(), range(len(ref_flat))))[int(k)] for k in flat])
    vol.features = torch.nn.Parameter(vol.features)
    prior = torch.from_numpy(g["recip/tsdf_delta"]).to(dev)[None, None]
    for name in ("mesh", "rand"):
        q = torch.from_numpy(gg[f"{name}_q"]).to(dev)[None]
        r = torch.from_numpy(gg[f"{name}_r"]).to(dev)
        sdf = vol.decode_pts(q, model.nerf, prior, is_coords=True)
        assert sdf.requires_grad
        np.testing.assert_allclose(sdf.detach()[0, :, :, 0].cpu().numpy(), gg[f"{name}_sdf"], atol=1e-4 if mode == "tc16" else 2e-6, rtol=0)
        vol.features.grad = None
        (sdf[0, :, :, 0] * r).sum().backward()
        grad = vol.features.grad.cpu().numpy()
        ref = gg[f"{name}_grad"][perm]
        assert np.abs(ref).max() > 1e-4
        np.testing.assert_allclose(grad, ref, atol=3e-7, rtol=2e-4)
    # one Adam step through the public objects, as NeuralMap.optimize does, then write back
    opt = torch.optim.Adam([vol.features], lr=0.001)
    before = vol.features.detach().clone()
    opt.zero_grad()
    q = torch.from_numpy(gg["mesh_q"]).to(dev)[None]
    loss = vol.decode_pts(q, model.nerf, prior, is_coords=True).abs().mean()
    loss.backward()
    opt.step()
    assert float((vol.features.detach() - before).abs().max()) > 0
    vol.insert(vol.active_coordinates, vol.features, vol.weights, vol.num_hits)
    f2, _, _ = vol.query(vol.active_coordinates)
    assert torch.equal(f2, vol.features.detach())
    config.set_mlp_mode("fp32")


@pytest.mark.parametrize("mode", ["fp32", "tc16"])
def test_arkit_shape_stream_vs_oracle(model, tcnn_params, dev, mode):
    """BASELINE.json configs[4] shape: 256x192 depth, 2 cm voxels, 10 % invalidated pixels -- the CUDA path
    against the numpy oracle on the same seeded frames (no golden file: the oracle itself is pinned by
    tests/test_oracle_golden.py).  ids / weights exact; features and SDF within the mode's tolerance."""
    from bnv_fusion_b200 import config
    config.set_mlp_mode(mode)
    spec = synth.stream_spec("arkit")
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    vol = _volume(spec, dev, pool_capacity=1 << 18)
    assert tuple(int(v) for v in vol._n_xyz_host) == tuple(int(v) for v in grid.n_xyz)
    vm = O.VoxelMap(grid)
    for fi in range(3):
        d, K, T = synth.make_frame(spec, fi, seed=5)
        model.fuse_depth_frame(vol, _depth_to_dev(d, dev), K, T, spec.max_depth)
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        feats, counts, flat, coords, _, _ = O.encode_pointcloud(O.backproject(depth, mask, K, T), grid,
                                                                tcnn_params["encoder"], 8)
        O.integrate(vm, flat, feats, counts)
    vol.check_status()
    flat, feats, w, h = _map_sorted(vol)
    rflat = np.sort(np.fromiter(vm.index.keys(), dtype=np.int64))
    assert len(flat) > 5000 and np.array_equal(flat, rflat)
    rfeats, rw, _, _ = vm.query(flat)
    np.testing.assert_allclose(w, rw, atol=1e-6, rtol=0)
    np.testing.assert_allclose(feats, rfeats, atol=FEAT_ATOL if mode == "fp32" else 5e-3, rtol=0)
    coords = vol.to_tensor()[0]
    vol.weights += 8.0
    vm.weights[: len(vm)] += 8.0
    q = O.meshlize_samples(coords[:1500].cpu().numpy())
    sdf = vol.decode_pts(torch.from_numpy(q).to(dev)[None], model.nerf, None, is_coords=True)[0, :, :, 0].cpu().numpy()
    ref = O.decode_pts(vm, q.reshape(-1, 3), tcnn_params["decoder"], 8).reshape(-1, 27)
    assert np.abs(sdf - ref).max() <= (SDF_ATOL_FP32 * 2 if mode == "fp32" else SDF_ATOL), np.abs(sdf - ref).max()
    assert (ref != np.float32(spec.voxel_size)).mean() > 0.3          # blended values, not the fallback


def test_host_buffer_call_matches_device_call(model, dev):
    """bnv_fuse_frame_host (pinned host depth in, frame statistics out to pinned host memory, one call) leaves the
    map and the statistics of bnv_fuse_frame on a device-resident frame, bit for bit."""
    spec = synth.stream_spec("parity64")
    va, vb = _volume(spec, dev, pool_capacity=1 << 16), _volume(spec, dev, pool_capacity=1 << 16)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    stats_host = torch.zeros(4, dtype=torch.int64).pin_memory()
    fr = [synth.make_frame(spec, fi, seed=3) for fi in range(6)]
    hosts = [torch.from_numpy(d.view(np.int16).copy()).pin_memory() for d, _, _ in fr]
    for fi in range(6):
        d, K, T = fr[fi]
        model.fuse_depth_frame(va, _depth_to_dev(d, dev), K, T, spec.max_depth, stats=stats)
        # frames 0..2 hint the next frame (prefetched on the copy stream), 3..5 do not (copy on the call's stream)
        nxt = hosts[fi + 1] if fi < 3 else None
        model.fuse_depth_frame_host(vb, hosts[fi], K, T, spec.max_depth, stats_host=stats_host, next_depth_mm_host=nxt)
        torch.cuda.synchronize()
        assert stats.cpu().tolist() == stats_host.tolist() and int(stats_host[3]) > 0
    a, b = _map_sorted(va), _map_sorted(vb)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    with pytest.raises(RuntimeError):       # a frame larger than the map's max_points fails loudly
        big = torch.zeros((4096, 4096), dtype=torch.int16).pin_memory()
        model.fuse_depth_frame_host(vb, big, K, T, spec.max_depth)


@pytest.mark.parametrize("mode", ["fp32", "tc16"])
@pytest.mark.parametrize("hw", [(45, 70), (8, 32), (61, 33), (1, 1)])
def test_ragged_frame_sizes_vs_oracle(model, tcnn_params, dev, mode, hw):
    """Frame sizes that are not multiples of the prepass' 32 x 8 pixel tiles (partial tiles, replicate padding at the
    image border inside a tile, a single pixel): fused depth path vs the oracle, ids / weights exact."""
    from bnv_fusion_b200 import config
    config.set_mlp_mode(mode)
    H, W = hw
    spec = synth.stream_spec("parity64")
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    vol = _volume(spec, dev, pool_capacity=1 << 16)
    vm = O.VoxelMap(grid)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    n_total = 0
    for fi in range(3):
        d, K, T = synth.make_frame(spec, fi, seed=9)
        d = np.ascontiguousarray(d[10:10 + H, 5:5 + W])
        K = K.copy()
        K[0, 2] -= 5
        K[1, 2] -= 10
        model.fuse_depth_frame(vol, _depth_to_dev(d, dev), K, T, spec.max_depth, stats=stats)
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        feats, counts, flat, coords, _, _ = O.encode_pointcloud(O.backproject(depth, mask, K, T), grid, tcnn_params["encoder"], 8)
        if flat is not None:
            O.integrate(vm, flat, feats, counts)
            n_total += len(flat)
        st = stats.tolist()
        assert st[0] == int(mask.sum()) and st[3] == (0 if flat is None else len(flat))
    vol.check_status()
    flat, feats, w, h = _map_sorted(vol)
    rflat = np.sort(np.fromiter(vm.index.keys(), dtype=np.int64)) if len(vm) else np.zeros(0, np.int64)
    assert np.array_equal(flat, rflat)
    if len(flat):
        rfeats, rw, _, _ = vm.query(flat)
        np.testing.assert_allclose(w, rw, atol=1e-6, rtol=0)
        np.testing.assert_allclose(feats, rfeats, atol=FEAT_ATOL if mode == "fp32" else 5e-3, rtol=0)
    if hw == (45, 70):
        assert len(flat) > 300
    config.set_mlp_mode("tc16")
