"""Phase timers of the chain inside the real decode kernel (decode_pts over every active voxel's 27 samples).
Build the library with the timers first:   make -C bnv_fusion_b200/csrc -B EXTRA=-DBNV_CHAIN_PROFILE=1
and rebuild without EXTRA afterwards.  Prints average cycles per corner (item) and phase, thread 0 of each chain."""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnv_fusion_b200 import synth, _lib
from bnv_fusion_b200.model import LitFusionPointNet
from bnv_fusion_b200.volume import SparseVolume
lib = _lib.load()
prof = getattr(lib, "bnv_debug_chain_profile", None)
if prof is None:
    sys.exit("library built without -DBNV_CHAIN_PROFILE=1")
p = np.load("tests/golden/tcnn_params.npz")
cfg = {"trainer": {"dense_volume": False}, "model": {"feature_vector_size": 8, "voxel_size": 0.01, "min_pts_in_grid": 8,
       "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
m = LitFusionPointNet(cfg)
m.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]), "nerf.model.params": torch.from_numpy(p["decoder"])})
m.eval(); m.cuda(); m.freeze()
spec = synth.stream_spec("lounge")
vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8)
for i in range(8):
    d, K, T = synth.make_frame(spec, i, seed=0)
    m.fuse_depth_frame(vol, torch.from_numpy(d.view(np.int16)).cuda().view(torch.uint16), K, T, spec.max_depth)
vol.to_tensor(); vol.weights += 8.0
A = vol.active_coordinates.shape[0]
off = torch.tensor([[a, b, c] for a in (-.5, 0, .5) for b in (-.5, 0, .5) for c in (-.5, 0, .5)], device="cuda")
qc = (vol.active_coordinates.float()[:, None, :] + off[None]).reshape(1, A, 27, 3).contiguous()
out = (C.c_ulonglong * 16)()
names = ["wait L0", "epilogue+issue L1", "shadow1", "wait L1", "epilogue+issue L2", "shadow2",
         "wait L2", "epilogue 3", "finish (issue L3 + next L0)"]


def report(title, extra):
    items = max(out[11], 1)
    print("==", title)
    tot = 0.0
    for i, n in enumerate(names):
        print(f"{n:32s} {out[i] / items:8.1f} cyc / item")
        tot += out[i] / items
    print(f"{'sum of phases':32s} {tot:8.1f} cyc / item")
    print(f"{'per-tile precompute':32s} {out[9] / items:8.1f} cyc / item (amortised over the 8 corners)")
    print(f"{'chain lifetime':32s} {out[10] / items:8.1f} cyc / item; {items} items timed; {extra}")


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
vol.decode_pts(qc, m.nerf, None, is_coords=True)
prof(out, 1)                                   # discard warm-up (and the encode kernel's counts)
for cold in (False, True):
    if cold:
        flush.zero_()
    vol.decode_pts(qc, m.nerf, None, is_coords=True)
    prof(out, 1)
    report("decode_tc_kernel, " + ("L2 flushed" if cold else "warm L2") + " (shadow1 = gather next + blend previous, shadow2 = stage next)",
           f"{A * 27} queries")
# the encode chain kernel (shadow1 = drain previous output + reductions, shadow2 = stage next corner)
for cold in (False, True):
    d, K, T = synth.make_frame(spec, 9, seed=0)
    dd = torch.from_numpy(d.view(np.int16)).cuda().view(torch.uint16)
    torch.cuda.synchronize(); prof(out, 1)
    if cold:
        flush.zero_()
    m.fuse_depth_frame(vol, dd, K, T, spec.max_depth)
    prof(out, 1)
    report("encode_chain_kernel, " + ("L2 flushed" if cold else "warm L2"), "one 640x480 frame")
