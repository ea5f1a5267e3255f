"""Short workload for ncu captures: 6 fused lounge frames (one call each), two 7-frame batches, one decode_pts over
every active voxel's 27 samples, one factored block decode, one TSDF integration (see tools/gpu_profile.sh)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnv_fusion_b200 import synth
from bnv_fusion_b200.model import LitFusionPointNet
from bnv_fusion_b200.volume import SparseVolume, get_world_range
from bnv_fusion_b200.tsdf import TSDFVolume
dev = "cuda:0"
p = np.load("tests/golden/tcnn_params.npz")
cfg = {"trainer": {"dense_volume": False}, "model": {"feature_vector_size": 8, "voxel_size": 0.01, "min_pts_in_grid": 8,
       "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
m = LitFusionPointNet(cfg)
m.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]), "nerf.model.params": torch.from_numpy(p["decoder"])})
m.eval(); m.cuda(); m.freeze()
spec = synth.stream_spec("lounge")
vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev)
frames = [synth.make_frame(spec, i, seed=0) for i in range(6)]
dd = [torch.from_numpy(d.view(np.int16)).to(dev).view(torch.uint16) for d, _, _ in frames]
for i in range(6):
    m.fuse_depth_frame(vol, dd[i], frames[i][1], frames[i][2], spec.max_depth)
volb = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, frame_batch=7)
fb = [synth.make_frame(spec, i, seed=0) for i in range(14)]
db = [torch.from_numpy(d.view(np.int16)).to(dev).view(torch.uint16) for d, _, _ in fb]
for b0 in (0, 7):
    m.fuse_depth_frames(volb, db[b0:b0 + 7], np.stack([K for _, K, _ in fb[b0:b0 + 7]]), np.stack([T for _, _, T in fb[b0:b0 + 7]]),
                        spec.max_depth)
volb.check_status()
vol.to_tensor()
vol.weights += 8.0
A = vol.active_coordinates.shape[0]
off = torch.tensor([[a, b, c] for a in (-.5, 0, .5) for b in (-.5, 0, .5) for c in (-.5, 0, .5)], device=dev)
qc = (vol.active_coordinates.float()[:, None, :] + off[None]).reshape(1, A, 27, 3).contiguous()
vol.decode_pts(qc, m.nerf, None, is_coords=True)
vol.decode_voxel_blocks(m.nerf)
mn, mx, _ = get_world_range(spec.dimensions, 0.025)
tsdf = TSDFVolume(np.stack([mn, mx], 1), 0.025, device=dev, verbose=False)
tsdf.integrate(None, dd[0], frames[0][1], frames[0][2], 1.0)
torch.cuda.synchronize()
print("active voxels", A, "queries", A * 27)
