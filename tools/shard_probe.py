"""single-GPU probe of the encode kernel's tile-shard path (world = 2, no NCCL): rank 0 and rank 1 maps on one device"""
import os, sys, faulthandler
import numpy as np, torch
faulthandler.dump_traceback_later(40, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bnv_fusion_b200 import synth
from bnv_fusion_b200.model import LitFusionPointNet
from bnv_fusion_b200.volume import SparseVolume
dev = "cuda:0"
p = np.load(os.path.join(ROOT, "tests", "golden", "tcnn_params.npz"))
cfg = {"trainer": {"dense_volume": False}, "model": {"feature_vector_size": 8, "voxel_size": 0.01, "min_pts_in_grid": 8,
       "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
model = LitFusionPointNet(cfg)
model.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]), "nerf.model.params": torch.from_numpy(p["decoder"])})
model.eval(); model.cuda(); model.freeze()
spec = synth.stream_spec("lounge")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
stats = torch.zeros(4, dtype=torch.int64, device=dev)
for rank in range(world):
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, pool_capacity=1 << 21)
    vol.set_shard(rank, world, 4)
    for fi in range(3):
        d, K, T = synth.make_frame(spec, fi, seed=2)
        dd = torch.from_numpy(d.view(np.int16)).to(dev).view(torch.uint16)
        print("rank", rank, "frame", fi, "launch", flush=True)
        model.fuse_depth_frame(vol, dd, K, T, spec.max_depth, stats=stats)
        torch.cuda.synchronize()
        print("rank", rank, "frame", fi, "stats", stats.tolist(), "voxels", len(vol), flush=True)
    vol.check_status()
    del vol
print("ok")
