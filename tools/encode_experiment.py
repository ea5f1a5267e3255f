"""Ablation of the tensor-core encode kernel (profiling aid).  BNV_DEBUG_ENCODE is a bit mask: 1 no MMA chain,
2 no feature reductions, 4 no claims (+ no reductions), 8 whole tiles per chain instead of the (tile, corner)
split, 16 run-length shuffle aggregation of the reductions; each is timed depth-in and points-in (no
back-projection).  usage: python tools/encode_experiment.py 0 1 2 4 7 ...   (results: profiles/r1c_umma_microbench2.txt)"""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = r'''
import os, sys, numpy as np, torch, ctypes as C
sys.path.insert(0, ".")
from bnv_fusion_b200 import synth, _lib
from bnv_fusion_b200.model import LitFusionPointNet, backproject
from bnv_fusion_b200.volume import SparseVolume
p = np.load("tests/golden/tcnn_params.npz")
cfg = {"trainer": {"dense_volume": False}, "model": {"feature_vector_size": 8, "voxel_size": 0.01, "min_pts_in_grid": 8, "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
m = LitFusionPointNet(cfg); m.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]), "nerf.model.params": torch.from_numpy(p["decoder"])}); m.eval(); m.cuda(); m.freeze()
spec = synth.stream_spec("lounge")
vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8)
lib = _lib.load(); lib.bnv_map_set_timing(vol._handle, 1)
frames = [synth.make_frame(spec, i, seed=0) for i in range(4)]
dd = [torch.from_numpy(d.view(np.int16)).cuda().view(torch.uint16) for d, _, _ in frames]
pts = [backproject(vol, dd[i], frames[i][1], frames[i][2], 3.0) for i in range(4)]
def run(kind):
    ts = []
    for it in range(12):
        i = it % 4
        if kind == "depth": m.fuse_depth_frame(vol, dd[i], frames[i][1], frames[i][2], 3.0)
        else: m.fuse_points(vol, pts[i])
        a, b = C.c_float(), C.c_float(); lib.bnv_map_get_timing(vol._handle, C.byref(a), C.byref(b)); ts.append((a.value, b.value))
    ts = np.array(ts[4:]); return ts.mean(0)
print("RESULT", os.environ.get("BNV_DEBUG_ENCODE", "0"), "depth enc/fin ms", run("depth").round(4).tolist(), "points enc/fin ms", run("points").round(4).tolist())
'''
for dbg in (sys.argv[1:] or ["0"]):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BNV_DEBUG_ENCODE=dbg), capture_output=True, text=True)
    print([l for l in r.stdout.splitlines() if l.startswith("RESULT")] or r.stderr[-400:])
