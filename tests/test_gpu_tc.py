"""GPU (B200): the tcgen05 tensor-core mode (BNV_MLP_TC16: fp16 operands, fp32 accumulation in TMEM).

Bars: voxel ids / keys / counts bit-exact (index maths never touches the tensor cores); SDF within the
1e-4 abs budget of BASELINE.json against the fp32 goldens; MLP outputs additionally within fp16
operand-rounding noise of the fp16-emulating oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import bnv_oracle as O
from bnv_fusion_b200 import config, synth
from test_gpu_parity import _depth_to_dev, _golden_map_sorted, _map_sorted, _volume, dev, model  # noqa: F401

pytestmark = pytest.mark.gpu

SDF_ATOL = 1e-4


@pytest.fixture(autouse=True)
def tc_mode():
    config.set_mlp_mode("tc16")
    yield


def _close_fp16(a, ref, rel=4e-3, abs_=4e-3):
    err = np.abs(a - ref)
    assert (err <= abs_ + rel * np.abs(ref)).all(), (err.max(), np.abs(ref).max())


def test_tc_mlp_forward(model, golden_dir, tcnn_params, dev):
    g = np.load(os.path.join(golden_dir, "golden_edge.npz"))
    rng = np.random.default_rng(0)
    for n in (64, 1, 127, 128, 129, 5000, 200_000):
        xe = rng.uniform(-1, 1, size=(n, 6)).astype(np.float32)
        ye = model.pointnet_backbone.model(torch.from_numpy(xe).to(dev)).cpu().numpy()
        ref16 = O.mlp_forward(tcnn_params["encoder"], xe, 6, 8, mode="fp16")
        _close_fp16(ye, ref16, rel=2e-3, abs_=2e-3)
        xd = np.concatenate([rng.uniform(-1, 1, size=(n, 9)), rng.normal(0, 0.8, size=(n, 8))], 1).astype(np.float32)
        yd = model.nerf.model(torch.from_numpy(xd).to(dev)).cpu().numpy()
        ref16 = O.mlp_forward(tcnn_params["decoder"], xd, 17, 1, mode="fp16")
        _close_fp16(yd, ref16, rel=2e-3, abs_=2e-3)
        ref32 = O.mlp_forward(tcnn_params["decoder"], xd, 17, 1, mode="fp32")
        assert np.abs(yd - ref32).max() * 0.01 < SDF_ATOL          # SDF = y * voxel_size
    ye = model.pointnet_backbone.model(torch.from_numpy(g["kat_enc_x"]).to(dev)).cpu().numpy()
    _close_fp16(ye, g["kat_enc_y"])


@pytest.mark.parametrize("path", ["api", "fused_depth"])
def test_tc_parity64_stream(model, golden_dir, dev, path):
    g = np.load(os.path.join(golden_dir, "golden_parity64.npz"))
    spec = synth.stream_spec("parity64")
    vol = _volume(spec, dev, pool_capacity=1 << 16)
    for fi in range(g["depth"].shape[0]):
        d, K, T = g["depth"][fi], g["K"][fi], g["T_wc"][fi]
        if path == "fused_depth":
            model.fuse_depth_frame(vol, _depth_to_dev(d, dev), K, T, spec.max_depth)
            continue
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        pts6 = torch.from_numpy(O.backproject(depth, mask, K, T)).to(dev)
        feats, counts, flat, coords, navg = model.encode_pointcloud(
            pts6[None], vol.n_xyz, vol.min_coords, vol.max_coords, vol.voxel_size, return_dense=False)
        if f"recip/f{fi}_flat" in g.files:
            assert np.array_equal(flat.cpu().numpy(), g[f"recip/f{fi}_flat"])
            assert np.array_equal(counts.cpu().numpy(), g[f"recip/f{fi}_counts"])
            _close_fp16(feats.cpu().numpy(), g[f"recip/f{fi}_feats"])
        model._integrate(vol, coords, feats, counts)
    vol.check_status()
    flat, feats, w, h = _map_sorted(vol)
    rflat, rfeats, rw, rh = _golden_map_sorted(g, "recip", vol._n_xyz_host)
    assert np.array_equal(flat, rflat)                      # bit-exact voxel set
    np.testing.assert_allclose(w, rw, atol=1e-6, rtol=0)    # weights depend on counts only
    _close_fp16(feats, rfeats)
    prior = torch.from_numpy(g["recip/tsdf_delta"]).to(dev)[None, None]
    qm = torch.from_numpy(g["recip/q_mesh"]).to(dev)[None]
    for p, key in ((None, "sdf_mesh"), (prior, "sdf_mesh_prior")):
        sdf = vol.decode_pts(qm, model.nerf, p, is_coords=True)[0, :, :, 0].cpu().numpy()
        assert np.abs(sdf - g["recip/" + key]).max() <= SDF_ATOL, np.abs(sdf - g["recip/" + key]).max()
    qr = torch.from_numpy(g["recip/q_rand"]).to(dev)[None, :, None, :]
    sdf = vol.decode_pts(qr, model.nerf, prior, is_coords=True)[0, :, 0, 0].cpu().numpy()
    assert np.abs(sdf - g["recip/sdf_rand_prior"]).max() <= SDF_ATOL


def test_tc_full_frame_paths_agree(model, dev):
    from bnv_fusion_b200.model import backproject
    spec = synth.stream_spec("lounge")
    vols = [_volume(spec, dev, pool_capacity=1 << 21) for _ in range(2)]
    for fi in range(2):
        d, K, T = synth.make_frame(spec, fi, seed=1)
        dd = _depth_to_dev(d, dev)
        model.fuse_depth_frame(vols[0], dd, K, T, spec.max_depth)
        pts = backproject(vols[1], dd, K, T, spec.max_depth)
        feats, counts, flat, coords, navg = model.encode_pointcloud(
            pts[None], vols[1].n_xyz, vols[1].min_coords, vols[1].max_coords, vols[1].voxel_size, return_dense=False)
        model._integrate(vols[1], coords, feats, counts)
    a, b = _map_sorted(vols[0]), _map_sorted(vols[1])
    # same voxels and weights; features equal up to fp32 summation order (vector fp32 reductions)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    np.testing.assert_allclose(a[1], b[1], atol=2e-5, rtol=0)
    # fp32 CUDA-core mode on the same frames: same voxels, features within fp16 noise
    config.set_mlp_mode("fp32")
    v32 = _volume(spec, dev, pool_capacity=1 << 21)
    for fi in range(2):
        d, K, T = synth.make_frame(spec, fi, seed=1)
        model.fuse_depth_frame(v32, _depth_to_dev(d, dev), K, T, spec.max_depth)
    c = _map_sorted(v32)
    config.set_mlp_mode("tc16")
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[2], c[2])
    _close_fp16(a[1], c[1])
    # decode on the 512^3 map: tensor-core vs CUDA-core SDF within budget
    vol = vols[0]
    coords, _, _, _ = vol.to_tensor()
    vol.weights += 8.0
    tc = vol.decode_voxel_blocks(model.nerf)
    # the factored block decode (G table + blend) is bit-identical to the generic per-query kernel
    prior = (torch.randn(1, 1, 31, 33, 35, device=dev) * 0.004)
    tcp = vol.decode_voxel_blocks(model.nerf, prior)
    n_chk = 20000
    q = torch.from_numpy(O.meshlize_samples(coords[:n_chk].cpu().numpy())).to(dev)[None]
    assert torch.equal(vol.decode_pts(q, model.nerf, None, is_coords=True)[0, :, :, 0], tc[:n_chk].reshape(n_chk, 27))
    assert torch.equal(vol.decode_pts(q, model.nerf, prior, is_coords=True)[0, :, :, 0], tcp[:n_chk].reshape(n_chk, 27))
    part = vol.decode_voxel_blocks(model.nerf, prior, first=1234, count=777)
    assert torch.equal(part, tcp[1234:1234 + 777])
    config.set_mlp_mode("fp32")
    ref = vol.decode_voxel_blocks(model.nerf)
    config.set_mlp_mode("tc16")
    assert float((tc - ref).abs().max()) <= SDF_ATOL


def test_tc_lounge_crop_512_grid(model, golden_dir, dev):
    """The 512^3 lounge-crop goldens (minted from the unmodified reference sources) in the BENCHMARKED mode: voxel ids
    exact, weights exact, features within fp16 operand noise, SDF <= 1e-4."""
    g = np.load(os.path.join(golden_dir, "golden_lounge_crop.npz"))
    spec = synth.stream_spec("lounge")
    vol = _volume(spec, dev, pool_capacity=1 << 18)
    for fi in range(2):
        model.fuse_depth_frame(vol, _depth_to_dev(g["depth"][fi], dev), g["K"][fi], g["T_wc"][fi], spec.max_depth)
    vol.check_status()
    flat, feats, w, h = _map_sorted(vol)
    rflat, rfeats, rw, rh = _golden_map_sorted(g, "recip", vol._n_xyz_host)
    assert np.array_equal(flat, rflat)
    np.testing.assert_allclose(w, rw, atol=1e-6, rtol=0)
    _close_fp16(feats, rfeats)
    vol.to_tensor()
    vol.weights *= float(g["recip/weight_scale"])
    prior = torch.from_numpy(g["recip/tsdf_delta"]).to(dev)[None, None]
    for qk, sk, p in (("q_mesh", "sdf_mesh", None), ("q_mesh", "sdf_mesh_prior", prior), ("q_rand", "sdf_rand_prior", prior)):
        q = torch.from_numpy(g["recip/" + qk].reshape(-1, 3)).to(dev)[None, :, None, :]
        sdf = vol.decode_pts(q, model.nerf, p, is_coords=True)[0, :, 0, 0].cpu().numpy()
        assert np.abs(sdf - g["recip/" + sk].reshape(-1)).max() <= SDF_ATOL


def test_tc_lounge_full_frame_vs_oracle(model, tcnn_params, dev):
    """BASELINE.json configs[1] at full size in the benchmarked mode, DIRECTLY against the oracle: two 640x480 lounge
    frames fused into the 512^3 grid by the tensor-core path vs oracle.encode_pointcloud / integrate on the same frames
    (local_point_fusion.py:81-151,647-673): voxel ids, counts and weights exact, features within fp16 operand noise,
    27-sample SDF of 3000 voxels <= 1e-4."""
    spec = synth.stream_spec("lounge")
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    vol = _volume(spec, dev, pool_capacity=1 << 21)
    assert vol._n_xyz_host == (512, 512, 512)
    vm = O.VoxelMap(grid)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    for fi in range(2):
        d, K, T = synth.make_frame(spec, fi, seed=2)
        assert d.shape == (480, 640)
        model.fuse_depth_frame(vol, _depth_to_dev(d, dev), K, T, spec.max_depth, stats=stats)
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        pts6 = O.backproject(depth, mask, K, T)
        feats, counts, flat, coords, _, _ = O.encode_pointcloud(pts6, grid, tcnn_params["encoder"], 8)
        O.integrate(vm, flat, feats, counts)
        st = stats.tolist()
        assert st[0] == int(mask.sum()) and st[3] == len(flat)          # valid pixels, voxels with count >= 8
    vol.check_status()
    flat, feats, w, h = _map_sorted(vol)
    rflat = np.sort(np.fromiter(vm.index.keys(), dtype=np.int64))
    assert len(flat) > 50_000 and np.array_equal(flat, rflat)
    rfeats, rw, _, _ = vm.query(flat)
    assert np.array_equal(w, rw.astype(np.float32))                     # clip(count / 32, 1) sums: exact counts
    _close_fp16(feats, rfeats)
    coords = vol.to_tensor()[0]
    vol.weights += 8.0
    vm.weights[: len(vm)] += 8.0
    sel = np.linspace(0, len(flat) - 1, 3000).astype(np.int64)
    q = O.meshlize_samples(coords[torch.from_numpy(sel).to(dev)].cpu().numpy())
    sdf = vol.decode_pts(torch.from_numpy(q).to(dev)[None], model.nerf, None, is_coords=True)[0, :, :, 0].cpu().numpy()
    ref = O.decode_pts(vm, q.reshape(-1, 3), tcnn_params["decoder"], 8).reshape(-1, 27)
    assert np.abs(sdf - ref).max() <= SDF_ATOL, np.abs(sdf - ref).max()
    assert (ref != np.float32(spec.voxel_size)).mean() > 0.3


def test_host_buffer_call_with_stale_prefetch_hint(model, dev):
    """A prefetch hint that is NOT followed (frame X hinted, frame Y passed): the call must wait for the hinted copy
    before overwriting the staging buffer and fuse frame Y (advisor finding, round 1)."""
    spec = synth.stream_spec("parity64")
    va, vb = _volume(spec, dev, pool_capacity=1 << 16), _volume(spec, dev, pool_capacity=1 << 16)
    fr = [synth.make_frame(spec, fi, seed=4) for fi in range(4)]
    hosts = [torch.from_numpy(d.view(np.int16).copy()).pin_memory() for d, _, _ in fr]
    order = [0, 2, 1, 3]
    for j, fi in enumerate(order):
        d, K, T = fr[fi]
        model.fuse_depth_frame(va, _depth_to_dev(d, dev), K, T, spec.max_depth)
        wrong = hosts[(fi + 1) % 4]                     # never the frame the next call passes (order is 0,2,1,3)
        model.fuse_depth_frame_host(vb, hosts[fi], K, T, spec.max_depth, next_depth_mm_host=wrong)
    torch.cuda.synchronize()
    for x, y in zip(_map_sorted(va), _map_sorted(vb)):
        assert np.array_equal(x, y) or np.allclose(x, y, atol=2e-5, rtol=0)
