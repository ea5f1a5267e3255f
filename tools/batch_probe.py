"""Frame-batch probe: cold (L2 flushed) time of one bnv_fuse_frames call per batch size on the bench workload, with the
per-kernel stage times, next to the single-frame call.  Usage: python tools/batch_probe.py [batch sizes ...]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bnv_fusion_b200 import _lib  # noqa: E402
from bnv_fusion_b200.model import LitFusionPointNet  # noqa: E402
from bnv_fusion_b200.volume import SparseVolume  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 7]
    dbg = int(os.environ.get("BNV_PROBE_DBG", "0"))   # prepass ablation flags (timing only: the maps are garbage)
    if dbg:
        _lib.load().bnv_debug_prepass.argtypes = [C.c_int]
        _lib.load().bnv_debug_prepass(dbg)
    dev = "cuda:0"
    lib = _lib.load()
    spec, frames = bench.make_frames(32)
    p = np.load(os.path.join(ROOT, "tests", "golden", "tcnn_params.npz"))
    cfg = {"trainer": {"dense_volume": False},
           "model": {"feature_vector_size": 8, "voxel_size": spec.voxel_size, "min_pts_in_grid": 8,
                     "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
    model = LitFusionPointNet(cfg)
    model.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]),
                           "nerf.model.params": torch.from_numpy(p["decoder"])})
    model.eval(); model.cuda(); model.freeze()
    devf = [torch.from_numpy(d.view(np.int16).copy()).to(dev).view(torch.uint16) for d, _, _ in frames]
    Ks = np.stack([K for _, K, _ in frames]); Ts = np.stack([T for _, _, T in frames])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    out = []
    for B in sizes:
        vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, frame_batch=B)

        def step(i):
            if B == 0:
                j = i % 32
                model.fuse_depth_frame(vol, devf[j], frames[j][1], frames[j][2], spec.max_depth, stats=stats)
            else:
                ids = [(i * B + j) % 32 for j in range(B)]
                model.fuse_depth_frames(vol, [devf[j] for j in ids], Ks[ids], Ts[ids], spec.max_depth, stats=stats)

        for i in range(6):
            step(i)
        torch.cuda.synchronize()
        steps = 20
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()
            ev[i][0].record()
            step(6 + i)
            ev[i][1].record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        lib.bnv_map_set_timing(vol._handle, 1)
        st = []
        for i in range(steps):
            flush.zero_()
            step(6 + steps + i)
            ms3 = (C.c_float * 3)()
            _lib.check(lib.bnv_map_get_timing_stages(vol._handle, ms3), "timing")
            st.append([ms3[0], ms3[1], ms3[2]])
        lib.bnv_map_set_timing(vol._handle, 0)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for i in range(steps):
            step(6 + 2 * steps + i)
        t1.record()
        torch.cuda.synchronize()
        warm = t0.elapsed_time(t1) / steps
        if not dbg:
            vol.check_status()
        nb = max(B, 1)
        st = np.mean(st, axis=0)
        rec = {"dbg": dbg, "batch": B, "ms_per_call": round(ms, 4), "frames_per_s_cold": round(nb * 1e3 / ms), "frames_per_s_warm": round(nb * 1e3 / warm),
               "prepass_ms": round(float(st[0]), 4), "encode_ms": round(float(st[1]), 4), "finalize_ms": round(float(st[2]), 4),
               "voxels": 0 if dbg else int(vol.to_tensor()[0].shape[0])}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del vol
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
