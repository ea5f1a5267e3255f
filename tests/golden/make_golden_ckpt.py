"""Mint tests/golden/golden_ckpt_sparse_volume.pth (+ golden_ckpt.npz) by running the UNMODIFIED reference
SparseVolume.save (/root/reference/src/models/sparse_volume.py:835-861) over the fakes of oracle/ref_stubs.py: a map
checkpoint written by the reference's own code, which the B200 SparseVolume must load.
Build-container only:   python tests/golden/make_golden_ckpt.py"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import bnv_oracle as O          # noqa: E402
from oracle import ref_stubs as R           # noqa: E402
from bnv_fusion_b200 import synth           # noqa: E402
from make_golden import ref_backproject     # noqa: E402


def main():
    work = tempfile.mkdtemp(prefix="bnv_golden_ckpt_")
    model, SparseVolume = R.build_reference(work, voxel_size=0.01, min_pts=8, mlp_mode="fp32")
    spec = synth.stream_spec("parity64")
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device="cpu")
    for fi in range(12):
        d, K, T = synth.make_frame(spec, fi, seed=0)
        pts6, _ = ref_backproject(d, K, T, spec.max_depth)
        with torch.no_grad(), R.cuda_div_semantics():
            feats, counts, flat, coords, navg = model.encode_pointcloud(
                torch.from_numpy(pts6)[None].clone(), vol.n_xyz, vol.min_coords, vol.max_coords, vol.voxel_size,
                return_dense=False)
            vol.track_n_pts(navg)                                   # run_e2e.py:93
            model._integrate(vol, coords, feats, counts)
    vol.to_tensor()
    vol.weights += 8.0                                              # as after NeuralMap.optimize's count_optim rounds
    path = os.path.join(HERE, "golden_ckpt")
    vol.save(path)                                                  # the reference's own writer
    ck = torch.load(path + "_sparse_volume.pth", weights_only=False)
    print({k: (tuple(v.shape), str(v.dtype)) if torch.is_tensor(v) else v for k, v in ck.items()})
    # what the reference decodes from a volume re-loaded from that file (its own load: sparse_volume.py:863-892)
    vol2 = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device="cpu")
    _tl = torch.load                       # the reference targets torch 1.10 (weights_only did not exist)
    torch.load = lambda f, *a, **k: _tl(f, *a, **{**k, "weights_only": False})
    try:
        vol2.load(path + "_sparse_volume.pth")
    finally:
        torch.load = _tl
    rng = np.random.default_rng(21)
    sel = rng.choice(vol2.active_coordinates.shape[0], size=150, replace=False)
    q = O.meshlize_samples(vol2.active_coordinates.numpy()[sel])
    with torch.no_grad(), R.cuda_div_semantics():
        sdf = vol2.decode_pts(torch.from_numpy(q)[None], model.nerf, None, is_coords=True)[0, :, :, 0].numpy()
    np.savez_compressed(os.path.join(HERE, "golden_ckpt.npz"), q_mesh=q, sdf_mesh=sdf)
    print("voxels", vol2.active_coordinates.shape[0], "blended", float((sdf != np.float32(0.01)).mean()),
          os.path.getsize(path + "_sparse_volume.pth"))


if __name__ == "__main__":
    main()
