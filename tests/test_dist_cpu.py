"""CPU, gloo, world_size 2 and 3: the tile-shard logic (ownership, one all-gather of boundary records per exchange
EPOCH -- every EVERY frames and once more before the map is read, carrying the current values of every shell voxel
integrated since the last epoch -- halo selection) on top of the numpy oracle.  The union of the ranks' owned voxels must equal
the single-process map bit for bit, and every rank must hold the halo voxels its own queries need."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import bnv_oracle as O          # noqa: E402
from bnv_fusion_b200 import dist as D      # noqa: E402
from bnv_fusion_b200 import synth          # noqa: E402

BRICK = 2      # 4-voxel bricks so that the 32^3 test grid has several bricks per rank
CAP = 4096
N_FRAMES = 10
EVERY = 4        # frames per exchange epoch: epochs after frames 4 and 8, flush after frame 10


def _sharded_encode(pts6, grid, enc, rank, world):
    """encode_pointcloud restricted to the (point, corner) rows whose voxel this rank owns"""
    rows = O.encode_rows(pts6, grid)
    if not rows["keep"].any():
        return None
    x = rows["mlp_in"].reshape(-1, 6)
    flat = rows["flat"].reshape(-1)
    mine = D.owner_of(rows["corner_ijk"].reshape(-1, 3), world, BRICK) == rank
    x, flat = x[mine], flat[mine]
    if flat.size == 0:
        return None
    f = O.mlp_forward(enc, x, 6, 8)
    uniq, inv, cnt = np.unique(flat, return_inverse=True, return_counts=True)
    sums = np.zeros((uniq.size, 8))
    np.add.at(sums, inv, f.astype(np.float64))
    mean = (sums / cnt[:, None]).astype(np.float32)
    ok = cnt >= 8
    return mean[ok], cnt[ok], uniq[ok]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = np.load(os.path.join(ROOT, "tests", "golden", "tcnn_params.npz"))
    spec = synth.stream_spec("parity64")
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    vm = O.VoxelMap(grid)
    nyz = grid.n_xyz[1] * grid.n_xyz[2]
    dirty = set()
    for fi in range(N_FRAMES):
        d, K, T = synth.make_frame(spec, fi, seed=0)
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        res = _sharded_encode(O.backproject(depth, mask, K, T), grid, p["encoder"], rank, world)
        flat = np.zeros(0, np.int64)
        if res is not None:
            feats, cnt, flat = res
            O.integrate(vm, flat, feats, cnt)
        # shell voxels integrated since the last epoch, each remembered once (bnv_map.cu: dirty flag + list)
        dirty.update(flat[D.on_brick_shell(D.unflatten(flat, grid.n_xyz), BRICK)].tolist())
        if (fi + 1) % EVERY == 0 or fi == N_FRAMES - 1:
            # one epoch: records with the CURRENT values -> ONE all-gather -> upsert what this rank needs
            b = np.fromiter(sorted(dirty), dtype=np.int64, count=len(dirty))
            dirty.clear()
            f, w, _, _ = vm.query(b)
            mine = torch.from_numpy(D.pack_halo(b, w, f, CAP))
            gathered = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            per_rank = D.unpack_gathered(torch.cat(gathered).numpy(), world, CAP)
            for r, (hf, hw, hfeat) in enumerate(per_rank):
                if r == rank:
                    continue
                need = D.select_needed(hf, grid.n_xyz, rank, world, BRICK)
                vm.insert(hf[need], hfeat[need], hw[need], np.zeros(int(need.sum()), np.float32))
    keys = np.fromiter(vm.index.keys(), dtype=np.int64)
    f, w, _, _ = vm.query(keys)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), keys=keys, feats=f, weights=w)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_shard_gloo(tmp_path, tcnn_params, world):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    # single-process reference
    spec = synth.stream_spec("parity64")
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    ref = O.VoxelMap(grid)
    for fi in range(N_FRAMES):
        d, K, T = synth.make_frame(spec, fi, seed=0)
        depth, mask = O.load_depth_u16(d, spec.max_depth)
        feats, counts, flat, _, _, _ = O.encode_pointcloud(O.backproject(depth, mask, K, T), grid, tcnn_params["encoder"], 8)
        O.integrate(ref, flat, feats, counts)
    ref_keys = np.sort(np.fromiter(ref.index.keys(), dtype=np.int64))
    nyz = grid.n_xyz[1] * grid.n_xyz[2]
    owned_all = []
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        keys, feats, weights = z["keys"], z["feats"], z["weights"]
        ijk = D.unflatten(keys, grid.n_xyz)
        own = D.owner_of(ijk, world, BRICK) == rank
        owned_all.append(keys[own])
        # every voxel this rank holds (owned or halo) carries exactly the single-process values
        f_ref, w_ref, _, found = ref.query(keys)
        assert found.all()
        assert np.array_equal(feats, f_ref) and np.array_equal(weights, w_ref)
        # halo completeness: every map voxel within one voxel of an owned voxel is present
        held = set(keys.tolist())
        rijk = D.unflatten(ref_keys, grid.n_xyz)
        need = D.select_needed(ref_keys, grid.n_xyz, rank, world, BRICK)
        assert set(ref_keys[need].tolist()) <= held
        for k in keys[own][:200]:
            kk = D.unflatten([k], grid.n_xyz)[0]
            for d in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 1), (-1, -1, -1)):
                q = kk + np.array(d)
                fq = int(q[0] * nyz + q[1] * grid.n_xyz[2] + q[2])
                if fq in ref.index:
                    assert fq in held, (rank, kk, d)
        # no foreign voxels beyond the brick shells
        assert D.on_brick_shell(ijk[~own], BRICK).all()
    union = np.sort(np.concatenate(owned_all))
    assert np.array_equal(union, ref_keys)          # disjoint cover of the single-process map


def test_halo_buffer_protocol():
    rng = np.random.default_rng(0)
    flat = rng.choice(32 ** 3, 100, replace=False)
    w = rng.random(100).astype(np.float32)
    f = rng.standard_normal((100, 8)).astype(np.float32)
    buf = D.pack_halo(flat, w, f, 256)
    assert buf.dtype == np.int32 and buf.size == D.HEADER_WORDS + 256 * D.RECORD_WORDS and buf[0] == 100
    (f2, w2, ft2), (e0, e1, e2) = D.unpack_gathered(np.concatenate([buf, D.pack_halo([], [], np.zeros((0, 8)), 256)]), 2, 256)
    assert np.array_equal(f2, flat) and np.array_equal(w2, w) and np.array_equal(ft2, f) and e0.size == 0
    with pytest.raises(RuntimeError):
        D.pack_halo(flat, w, f, 10)
    ijk = np.array([[0, 0, 0], [4, 0, 0], [4, 4, 0], [5, 5, 5], [7, 1, 1]])
    assert D.owner_of(ijk, 2, 2).tolist() == [0, 1, 0, 1, 1]
    assert D.on_brick_shell(ijk, 2).tolist() == [True, True, True, False, True]
    flat = ijk[:, 0] * 1024 + ijk[:, 1] * 32 + ijk[:, 2]
    assert np.array_equal(D.unflatten(flat, (32, 32, 32)), ijk)
    # (5,5,5) sits inside brick (1,1,1) (rank 1): only rank 1 touches it; (4,0,0) touches brick (0,0,0) of rank 0
    assert D.select_needed(flat, (32, 32, 32), 0, 2, 2).tolist() == [True, True, True, False, True]
    assert D.select_needed(flat, (32, 32, 32), 1, 2, 2).tolist() == [False, True, True, True, True]


def test_sender_side_routing_equals_receiver_side_selection():
    """the peer-memory exchange routes at the sender (`needed_by`, mirrored by halo_push_kernel); the all-gather
    path filters at the receiver (`select_needed`, mirrored by insert_halo_kernel): same rule, both directions"""
    rng = np.random.default_rng(1)
    n_xyz = (40, 33, 37)
    flat = rng.integers(0, n_xyz[0] * n_xyz[1] * n_xyz[2], 3000)
    flat = np.concatenate([flat, [0, n_xyz[0] * n_xyz[1] * n_xyz[2] - 1]])          # grid corners: clipped neighbourhoods
    for world, b in ((2, 2), (3, 1), (8, 2), (5, 3)):
        mask = D.needed_by(flat, n_xyz, world, b)
        for rank in range(world):
            assert np.array_equal((mask >> rank) & 1, D.select_needed(flat, n_xyz, rank, world, b).astype(np.int64))


def test_epoch_counter_counts_frames_of_batches():
    """TileShardedFusion._frame_done(n): a 7-frame bnv_fuse_frames call counts as 7 frames towards the boundary-exchange
    epoch (host logic only: the object is built without a device)."""
    from bnv_fusion_b200.dist import TileShardedFusion
    sh = TileShardedFusion.__new__(TileShardedFusion)
    sh.world, sh.exchange_every, sh._since, sh.epochs = 4, 16, 0, 0

    def fake_exchange():
        sh._since = 0
        sh.epochs += 1
    sh.exchange_now = fake_exchange
    for _ in range(2):
        sh._frame_done(7)
    assert (sh._since, sh.epochs) == (14, 0)
    sh._frame_done(7)                       # 21 >= 16: one epoch, counter back to zero
    assert (sh._since, sh.epochs) == (0, 1)
    for _ in range(16):
        sh._frame_done()
    assert (sh._since, sh.epochs) == (0, 2)
    sh.world = 1                            # a single GPU never exchanges
    for _ in range(5):
        sh._frame_done(7)
    assert sh.epochs == 2

