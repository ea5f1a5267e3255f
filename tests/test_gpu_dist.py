"""GPU (>= 2 B200s, NCCL): the tile shard through libbnv_b200 -- every rank fuses the same depth
stream, keeps the rows of the voxels it owns, and exchanges boundary voxels with ONE all-gather per
exchange epoch (here every 3 frames of 4, so both the periodic epoch and the flush-on-read are exercised).  Union of the owned voxels must equal the single-GPU map bit for bit; every rank must decode
the queries of its own tile exactly like the single-GPU map does."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
N_FRAMES = 4


def _setup(dev):
    from bnv_fusion_b200.model import LitFusionPointNet
    p = np.load(os.path.join(ROOT, "tests", "golden", "tcnn_params.npz"))
    cfg = {"trainer": {"dense_volume": False},
           "model": {"feature_vector_size": 8, "voxel_size": 0.01, "min_pts_in_grid": 8,
                     "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
    m = LitFusionPointNet(cfg)
    m.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]),
                       "nerf.model.params": torch.from_numpy(p["decoder"])})
    m.eval(); m.to(dev); m.freeze()
    return m


def _worker(rank, world, port, out_dir, mode, exchange="nccl", batch=0):
    import torch.distributed as dist
    from bnv_fusion_b200 import synth, config
    config.set_mlp_mode(mode)
    from bnv_fusion_b200.dist import TileShardedFusion
    from bnv_fusion_b200.volume import SparseVolume
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    model = _setup(dev)
    spec = synth.stream_spec("lounge")
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, pool_capacity=1 << 21, frame_batch=batch)
    shard = TileShardedFusion(vol, model, rank, world, brick_log2=4, exchange=exchange, exchange_every=3)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    rows = 0
    fr = [synth.make_frame(spec, fi, seed=2) for fi in range(N_FRAMES)]
    dd = [torch.from_numpy(d.view(np.int16)).to(dev).view(torch.uint16) for d, _, _ in fr]
    if batch:                                   # frames 0..batch-1 in one bnv_fuse_frames call, the rest one by one
        shard.fuse_depth_frames(dd[:batch], np.stack([K for _, K, _ in fr[:batch]]), np.stack([T for _, _, T in fr[:batch]]),
                                spec.max_depth, stats=stats)
        rows += int(stats[1])
    for fi in range(batch, N_FRAMES):
        shard.fuse_depth_frame(dd[fi], fr[fi][1], fr[fi][2], spec.max_depth, stats=stats)
        rows += int(stats[1])
    vol.check_status()
    coords, feats, weights, _ = vol.to_tensor()
    own = shard.owned_rows()
    # decode the 27 samples of this rank's own voxels (their corners may be halo voxels of the other rank)
    vol.weights += 8.0
    blocks = vol.decode_voxel_blocks(model.nerf)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), coords=coords.cpu().numpy(), feats=feats.cpu().numpy(),
             weights=weights.cpu().numpy(), own=own.cpu().numpy(), blocks=blocks.cpu().numpy(), rows=rows)
    dist.barrier()
    dist.destroy_process_group()


EXCHANGES = ["nccl", "p2p"]


@pytest.mark.parametrize("batch", [0, 3])
@pytest.mark.parametrize("exchange", EXCHANGES)
@pytest.mark.parametrize("mode", ["fp32", "tc16"])
def test_two_gpu_tile_shard(tmp_path, mode, exchange, batch):
    """fp32 mode (order-independent fixed-point sums): owned + halo values bit-identical to one GPU.
    tc16 mode (fp32 `red.add` partial sums, arrival order differs between runs): same voxels, values to
    summation-order noise.  batch = 3: the first three frames go through ONE sharded bnv_fuse_frames call."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from bnv_fusion_b200 import synth, config
    from bnv_fusion_b200.volume import SparseVolume
    world = 2
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 1000, str(tmp_path), mode, exchange, batch), nprocs=world, join=True)
    config.set_mlp_mode(mode)
    exact = mode == "fp32"

    def same(a, b, tol):
        if exact:
            return np.array_equal(a, b)
        assert np.abs(a - b).max() <= tol, np.abs(a - b).max()
        return True

    dev = "cuda:0"
    model = _setup(dev)
    spec = synth.stream_spec("lounge")
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, pool_capacity=1 << 21)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    rows = 0
    for fi in range(N_FRAMES):
        d, K, T = synth.make_frame(spec, fi, seed=2)
        model.fuse_depth_frame(vol, torch.from_numpy(d.view(np.int16)).to(dev).view(torch.uint16), K, T,
                               spec.max_depth, stats=stats)
        rows += int(stats[1])
    coords, feats, weights, _ = vol.to_tensor()
    vol.weights += 8.0
    blocks = vol.decode_voxel_blocks(model.nerf).cpu().numpy()
    flat = (coords[:, 0] * 512 * 512 + coords[:, 1] * 512 + coords[:, 2]).cpu().numpy()
    order = np.argsort(flat)
    ref = {"flat": flat[order], "feats": feats.cpu().numpy()[order], "w": weights.cpu().numpy()[order, 0],
           "blocks": blocks[order]}
    owned_flat, total_rows = [], 0
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        c = z["coords"]
        f = c[:, 0] * 512 * 512 + c[:, 1] * 512 + c[:, 2]
        pos = np.searchsorted(ref["flat"], f)
        assert np.array_equal(ref["flat"][pos], f)                     # no voxel the single-GPU map lacks
        assert same(z["feats"], ref["feats"][pos], 5e-5)               # owned + halo values
        assert np.array_equal(z["weights"][:, 0], ref["w"][pos])
        own = z["own"]
        assert (~own).sum() > 0                                         # halo copies exist
        from bnv_fusion_b200 import dist as D
        assert D.on_brick_shell(c[~own], 4).all()                       # ... only brick shells
        # SDF of the own voxels' 27 samples == single GPU (needs the halo to be complete)
        assert same(z["blocks"][own], ref["blocks"][pos][own], 1e-4)
        owned_flat.append(f[own])
        total_rows += int(z["rows"])
    assert np.array_equal(np.sort(np.concatenate(owned_flat)), ref["flat"])
    assert total_rows == rows                                           # the MLP rows were divided, not duplicated
    config.set_mlp_mode("tc16")
