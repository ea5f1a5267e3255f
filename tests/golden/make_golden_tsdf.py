"""Mint tests/golden/golden_tsdf.npz by running the reference's own third_parties/fusion.py (CPU mode:
numba + numpy; PyCUDA is absent) on the parity64 frames.  Build-container only."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
m = types.ModuleType("skimage.measure")
sys.modules.setdefault("skimage", types.ModuleType("skimage"))
sys.modules["skimage.measure"] = m
sys.modules["skimage"].measure = m
sys.path.insert(0, "/root/reference")
import third_parties.fusion as fusion          # noqa: E402
from bnv_fusion_b200 import synth              # noqa: E402
from oracle import bnv_oracle as O             # noqa: E402


def main():
    spec = synth.stream_spec("parity64")
    out = {}
    rng = np.random.default_rng(11)
    for tag, vs in (("v25", 0.025), ("v10", 0.01)):
        mn, mx, n = O.get_world_range(spec.dimensions, vs)
        bnds = np.zeros((3, 2)); bnds[:, 0] = mn; bnds[:, 1] = mx
        vol = fusion.TSDFVolume(bnds, voxel_size=vs, use_gpu=False)
        for fi in range(6):
            d, K, T = synth.make_frame(spec, fi, seed=0)
            depth, mask = O.load_depth_u16(d, spec.max_depth)
            depth32 = depth.astype(np.float32)            # rgbd goes through .cuda().float() (run_e2e.py:247-249)
            rgb = rng.integers(0, 256, size=(spec.height, spec.width, 3)).astype(np.float32)
            out[f"{tag}/rgb{fi}"] = rgb.astype(np.uint8)
            vol.integrate(rgb, depth32, K, T, obs_weight=1.0)
        tsdf, color = vol.get_volume()
        out[f"{tag}/tsdf"] = tsdf.copy()
        out[f"{tag}/color"] = color.copy()
        out[f"{tag}/weight"] = vol._weight_vol_cpu.copy()
        out[f"{tag}/bnds"] = bnds
        print(tag, tsdf.shape, float(tsdf.min()), float(tsdf.max()), int((vol._weight_vol_cpu > 0).sum()))
    np.savez_compressed(os.path.join(HERE, "golden_tsdf.npz"), **out)
    print(os.path.getsize(os.path.join(HERE, "golden_tsdf.npz")))


if __name__ == "__main__":
    main()
