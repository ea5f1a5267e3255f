"""Where the time of SparseVolume.extract_triangles goes: events around the two library calls and the torch glue."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bnv_fusion_b200 import _lib
from bnv_fusion_b200.model import LitFusionPointNet
from bnv_fusion_b200.volume import SparseVolume

dev = "cuda:0"
spec, frames = bench.make_frames(16)
p = np.load(os.path.join(ROOT, "tests", "golden", "tcnn_params.npz"))
cfg = {"trainer": {"dense_volume": False}, "model": {"feature_vector_size": 8, "voxel_size": spec.voxel_size, "min_pts_in_grid": 8,
       "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
m = LitFusionPointNet(cfg)
m.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]), "nerf.model.params": torch.from_numpy(p["decoder"])})
m.eval(); m.cuda(); m.freeze()
vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev)
for d, K, T in frames:
    m.fuse_depth_frame(vol, torch.from_numpy(d.view(np.int16)).to(dev).view(torch.uint16), K, T, spec.max_depth)
vol.to_tensor(); vol.weights += 8.0
blocks = vol.decode_voxel_blocks(m.nerf)
lib = vol._lib
n = vol.active_coordinates.shape[0]
sdf = blocks.detach().reshape(n, 27).float().contiguous()
coords = vol.active_coordinates.contiguous()
offsets = torch.empty(n + 1, dtype=torch.int32, device=dev)
mn = np.ascontiguousarray(vol.min_coords.detach().cpu().numpy().astype(np.float32))
nxyz = (C.c_int32 * 3)(*vol._n_xyz_host)


def ev():
    return torch.cuda.Event(enable_timing=True)


for rep in range(3):
    e = [ev() for _ in range(4)]
    t0 = time.perf_counter()
    e[0].record()
    _lib.check(lib.bnv_mesh_count(_lib.ptr(sdf), n, _lib.ptr(offsets), vol._stream()), "count")
    e[1].record()
    n_tri = int(offsets[-1].item())
    t1 = time.perf_counter()
    verts = torch.empty((3 * n_tri, 3), dtype=torch.float32, device=dev)
    keys = torch.empty(3 * n_tri, dtype=torch.int64, device=dev)
    e[2].record()
    _lib.check(lib.bnv_mesh_emit(_lib.ptr(sdf), _lib.ptr(coords), n, _lib.ptr(offsets), float(vol.voxel_size), _lib.ptr(mn), nxyz, n_tri,
                                 _lib.ptr(verts), _lib.ptr(keys), None, vol._stream()), "emit")
    e[3].record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"rep {rep}: voxels {n} triangles {n_tri}  count+scan {e[0].elapsed_time(e[1]):.3f} ms  emit {e[2].elapsed_time(e[3]):.3f} ms  "
          f"host: to n_tri {1e3 * (t1 - t0):.3f} ms, rest {1e3 * (t2 - t1):.3f} ms")
