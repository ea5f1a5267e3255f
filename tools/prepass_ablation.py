"""Ablation of the frame prepass kernel (profiling aid): bnv_debug_prepass flags 1 no claim atomics, 2 no global counter
atomics, 4 no back-projection, 8 no record stores, 16 no depth staging; prints the three stage times (prepass / MLP / finalize) of a lounge
frame, L2 flushed.  The map is garbage after a run with flags != 0 -- each configuration uses a fresh volume."""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnv_fusion_b200 import synth, _lib
from bnv_fusion_b200.model import LitFusionPointNet
from bnv_fusion_b200.volume import SparseVolume
lib = _lib.load()
lib.bnv_debug_prepass.argtypes = [C.c_int]
p = np.load("tests/golden/tcnn_params.npz")
cfg = {"trainer": {"dense_volume": False}, "model": {"feature_vector_size": 8, "voxel_size": 0.01, "min_pts_in_grid": 8,
       "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
m = LitFusionPointNet(cfg)
m.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]), "nerf.model.params": torch.from_numpy(p["decoder"])})
m.eval(); m.cuda(); m.freeze()
spec = synth.stream_spec("lounge")
frames = [synth.make_frame(spec, i, seed=0) for i in range(4)]
dd = [torch.from_numpy(d.view(np.int16)).cuda().view(torch.uint16) for d, _, _ in frames]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for flags in [int(a) for a in (sys.argv[1:] or ["0", "1", "3", "4", "16", "31"])]:
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8)
    lib.bnv_map_set_timing(vol._handle, 1)
    lib.bnv_debug_prepass(flags)
    for cold in (False, True):
        ts = []
        for it in range(10):
            if cold:
                flush.zero_()
            i = it % 4
            m.fuse_depth_frame(vol, dd[i], frames[i][1], frames[i][2], spec.max_depth)
            ms3 = (C.c_float * 3)()
            lib.bnv_map_get_timing_stages(vol._handle, ms3)
            ts.append(list(ms3))
        t = np.array(ts[3:]).mean(0)
        print(f"flags {flags:2d} {'cold' if cold else 'warm'}: prepass {t[0]*1e3:6.1f} us  mlp {t[1]*1e3:6.1f} us  finalize {t[2]*1e3:6.1f} us")
    lib.bnv_debug_prepass(0)
    del vol
