"""diagnostic: smoke()'s map comparison with the error printed instead of asserted"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bnv_fusion_b200 import synth
from bnv_fusion_b200.model import LitFusionPointNet
from bnv_fusion_b200.volume import SparseVolume
from oracle import bnv_oracle as O
dev = "cuda:0"
p = np.load(os.path.join(ROOT, "tests", "golden", "tcnn_params.npz"))
cfg = {"trainer": {"dense_volume": False}, "model": {"feature_vector_size": 8, "voxel_size": 0.01, "min_pts_in_grid": 8,
       "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
model = LitFusionPointNet(cfg)
model.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]), "nerf.model.params": torch.from_numpy(p["decoder"])})
model.eval(); model.cuda(); model.freeze()
spec = synth.stream_spec("parity64")
vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, pool_capacity=1 << 16)
grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
vm = O.VoxelMap(grid)
nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for fi in range(nfr):
    d, K, T = synth.make_frame(spec, fi, seed=0)
    model.fuse_depth_frame(vol, torch.from_numpy(d.view(np.int16)).to(dev).view(torch.uint16), K, T, spec.max_depth)
    depth, mask = O.load_depth_u16(d, spec.max_depth)
    feats, counts, flat, coords, _, _ = O.encode_pointcloud(O.backproject(depth, mask, K, T), grid, p["encoder"], 8)
    O.integrate(vm, flat, feats, counts)
    vol.check_status()
    coords, feats, weights, _ = vol.to_tensor()
    n = grid.n_xyz
    flat = (coords[:, 0] * n[1] * n[2] + coords[:, 1] * n[2] + coords[:, 2]).cpu().numpy()
    same = np.array_equal(np.sort(flat), np.sort(np.fromiter(vm.index.keys(), dtype=np.int64)))
    f_ref, w_ref, _, _ = vm.query(flat)
    df = np.abs(feats.cpu().numpy() - f_ref)
    print(f"frame {fi}: ids same={same} n={len(flat)} max|dfeat|={df.max():.3e} mean={df.mean():.3e} n>1e-3={(df.max(1)>1e-3).sum()} max|dw|={np.abs(weights.cpu().numpy()[:,0]-w_ref).max():.2e}")
