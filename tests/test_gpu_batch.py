"""GPU tests of the frame-batch path (bnv_fuse_frames): n frames through ONE pass of prepass / encoder MLP / finalize.

The contract (include/bnv_b200.h "frame batches"): the map ends up exactly as after n bnv_fuse_frame calls in the same
order -- which tests/test_gpu_parity.py and tests/test_gpu_tc.py pin against the oracle and the reference's golden
vectors -- bit for bit in the exact-parity arithmetic (integer sums), within fp32 summation-order noise on the tensor
cores; one case goes against the oracle directly."""
import numpy as np
import pytest
import torch

from oracle import bnv_oracle as O
from bnv_fusion_b200 import synth

from test_gpu_parity import _depth_to_dev, _map_sorted, _volume, dev, model  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def restore_mode():
    from bnv_fusion_b200 import config
    yield
    config.set_mlp_mode("tc16")


def _frames(spec, n, seed):
    fr = [synth.make_frame(spec, fi, seed=seed) for fi in range(n)]
    return fr, np.stack([K for _, K, _ in fr]), np.stack([T for _, _, T in fr])


def _assert_maps_equal(a, b, exact):
    assert np.array_equal(a[0], b[0]), "voxel ids differ"
    assert np.array_equal(a[2], b[2]), "weights differ"          # sums of exact multiples of 1/32, same order
    assert np.array_equal(a[3], b[3])
    if exact:
        assert np.array_equal(a[1], b[1]), np.abs(a[1] - b[1]).max()
    else:
        np.testing.assert_allclose(a[1], b[1], atol=1e-5, rtol=0)


@pytest.mark.parametrize("mode", ["fp32", "tc16"])
@pytest.mark.parametrize("n", [1, 2, 4, 7])
def test_batch_equals_sequential(model, dev, mode, n):
    """up to 7 frames per call (8 table words per grid cell); statistics are the sums of the per-frame statistics."""
    from bnv_fusion_b200 import config
    config.set_mlp_mode(mode)
    spec = synth.stream_spec("parity64")
    fr, Ks, Ts = _frames(spec, n, seed=11)
    va = _volume(spec, dev, pool_capacity=1 << 16)
    vb = _volume(spec, dev, pool_capacity=1 << 16, frame_batch=n)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    total = np.zeros(4, np.int64)
    devf = [_depth_to_dev(d, dev) for d, _, _ in fr]
    for i in range(n):
        model.fuse_depth_frame(va, devf[i], fr[i][1], fr[i][2], spec.max_depth, stats=stats)
        total += np.asarray(stats.tolist())
    model.fuse_depth_frames(vb, devf, Ks, Ts, spec.max_depth, stats=stats)
    assert stats.tolist() == total.tolist() and total[3] > 0
    va.check_status(); vb.check_status()
    a, b = _map_sorted(va), _map_sorted(vb)
    assert len(a[0]) > 500
    _assert_maps_equal(a, b, exact=mode == "fp32")


@pytest.mark.parametrize("mode", ["fp32", "tc16"])
def test_batches_and_single_frames_interleave(model, dev, mode):
    """batch(4) -> single frame -> batch(3, stacked tensor, shared K) -> single frame on ONE map: the per-frame table,
    the scratch rows and the counters are re-armed by every finalize flavour for every other one."""
    from bnv_fusion_b200 import config
    config.set_mlp_mode(mode)
    spec = synth.stream_spec("parity64")
    fr, Ks, Ts = _frames(spec, 9, seed=4)
    devf = [_depth_to_dev(d, dev) for d, _, _ in fr]
    va = _volume(spec, dev, pool_capacity=1 << 16)
    vb = _volume(spec, dev, pool_capacity=1 << 16, frame_batch=4)
    for i in range(9):
        model.fuse_depth_frame(va, devf[i], fr[i][1], fr[i][2], spec.max_depth)
    model.fuse_depth_frames(vb, devf[0:4], Ks[0:4], Ts[0:4], spec.max_depth)
    model.fuse_depth_frame(vb, devf[4], fr[4][1], fr[4][2], spec.max_depth)
    model.fuse_depth_frames(vb, torch.stack([d.view(torch.int16) for d in devf[5:8]]), spec.K, Ts[5:8], spec.max_depth)
    model.fuse_depth_frame(vb, devf[8], fr[8][1], fr[8][2], spec.max_depth)
    va.check_status(); vb.check_status()
    _assert_maps_equal(_map_sorted(va), _map_sorted(vb), exact=mode == "fp32")


def test_batch_vs_oracle(model, tcnn_params, dev):
    """the batch path against the numpy oracle directly (exact-parity arithmetic): ids / weights exact, features 2e-5"""
    from bnv_fusion_b200 import config
    config.set_mlp_mode("fp32")
    spec = synth.stream_spec("arkit")
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    n = 5
    fr, Ks, Ts = _frames(spec, n, seed=5)
    vol = _volume(spec, dev, pool_capacity=1 << 18, frame_batch=n)
    vm = O.VoxelMap(grid)
    model.fuse_depth_frames(vol, [_depth_to_dev(d, dev) for d, _, _ in fr], Ks, Ts, spec.max_depth)
    for d, K, T in fr:
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        feats, counts, flat, coords, _, _ = O.encode_pointcloud(O.backproject(depth, mask, K, T), grid, tcnn_params["encoder"], 8)
        O.integrate(vm, flat, feats, counts)
    vol.check_status()
    flat, feats, w, h = _map_sorted(vol)
    rflat = np.sort(np.fromiter(vm.index.keys(), dtype=np.int64))
    assert len(flat) > 5000 and np.array_equal(flat, rflat)
    rfeats, rw, _, _ = vm.query(flat)
    np.testing.assert_allclose(w, rw, atol=1e-6, rtol=0)
    np.testing.assert_allclose(feats, rfeats, atol=2e-5, rtol=0)


def test_batch_lounge_full_frames_tc(model, dev):
    """headline shape: batches of 7, 7 and 2 full 640x480 frames into the 512^3 grid on the tensor cores vs 16
    single-frame calls: ids / weights exact, features within fp32 summation-order noise"""
    from bnv_fusion_b200 import config
    config.set_mlp_mode("tc16")
    spec = synth.stream_spec("lounge")
    fr, Ks, Ts = _frames(spec, 16, seed=0)
    devf = [_depth_to_dev(d, dev) for d, _, _ in fr]
    va = _volume(spec, dev, pool_capacity=1 << 20)
    vb = _volume(spec, dev, pool_capacity=1 << 20, frame_batch=7)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    total = np.zeros(4, np.int64)
    for i in range(16):
        model.fuse_depth_frame(va, devf[i], fr[i][1], fr[i][2], spec.max_depth, stats=stats)
        total += np.asarray(stats.tolist())
    got = np.zeros(4, np.int64)
    for sl in (slice(0, 7), slice(7, 14), slice(14, 16)):
        model.fuse_depth_frames(vb, devf[sl], Ks[sl], Ts[sl], spec.max_depth, stats=stats)
        got += np.asarray(stats.tolist())
    assert got.tolist() == total.tolist()
    va.check_status(); vb.check_status()
    a, b = _map_sorted(va), _map_sorted(vb)
    assert len(a[0]) > 50000
    _assert_maps_equal(a, b, exact=False)


def test_batch_host_call_and_errors(model, dev):
    """bnv_fuse_frames_host (pinned host frames, prefetch hint for the next batch, statistics to pinned host memory)
    equals the device-resident call; misuse fails loudly"""
    from bnv_fusion_b200 import config
    config.set_mlp_mode("fp32")
    spec = synth.stream_spec("parity64")
    fr, Ks, Ts = _frames(spec, 12, seed=3)
    hosts = [torch.from_numpy(d.view(np.int16).copy()).pin_memory() for d, _, _ in fr]
    devf = [_depth_to_dev(d, dev) for d, _, _ in fr]
    va = _volume(spec, dev, pool_capacity=1 << 16, frame_batch=4)
    vb = _volume(spec, dev, pool_capacity=1 << 16, frame_batch=4)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    stats_host = torch.zeros(4, dtype=torch.int64).pin_memory()
    for b in range(3):
        sl = slice(4 * b, 4 * b + 4)
        model.fuse_depth_frames(va, devf[sl], Ks[sl], Ts[sl], spec.max_depth, stats=stats)
        # batch 0 hints batch 1 (prefetched on the copy stream); batch 1 hints the WRONG frames; batch 2 has no hint
        nxt = hosts[4:8] if b == 0 else hosts[0:4] if b == 1 else None
        model.fuse_depth_frames_host(vb, hosts[sl], Ks[sl], Ts[sl], spec.max_depth, stats_host=stats_host, next_depths_mm_host=nxt)
        torch.cuda.synchronize()
        assert stats.cpu().tolist() == stats_host.tolist() and int(stats_host[3]) > 0
    _assert_maps_equal(_map_sorted(va), _map_sorted(vb), exact=True)
    with pytest.raises(RuntimeError):                       # more frames than the table was laid out for
        model.fuse_depth_frames(va, devf[0:5], Ks[0:5], Ts[0:5], spec.max_depth)
    plain = _volume(spec, dev, pool_capacity=1 << 16)
    with pytest.raises(RuntimeError):                       # a volume without the batch layout
        model.fuse_depth_frames(plain, devf[0:2], Ks[0:2], Ts[0:2], spec.max_depth)
    with pytest.raises(RuntimeError):
        plain.set_frame_batch(8)                            # at most 7 frames per call
    small = _volume(spec, dev, pool_capacity=1 << 16, max_points=64 * 64)
    small.set_frame_batch(2)
    with pytest.raises(RuntimeError):                       # max_points covers one frame only
        model.fuse_depth_frames(small, devf[0:2], Ks[0:2], Ts[0:2], spec.max_depth)
    model.fuse_depth_frames(small, devf[0:1], Ks[0:1], Ts[0:1], spec.max_depth)       # one frame fits
    small.check_status()
