"""Mint tests/golden/golden_decode_grad.npz: gradients of SparseVolume.decode_pts w.r.t. volume.features as
the reference's own code + torch autograd compute them (NeuralMap.optimize, src/run_e2e.py:111-156), over
the fakes of oracle/ref_stubs.py (the fake tcnn network is plain torch ops, so autograd flows through the
reference's decode_pts unchanged).  Build-container only."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_stubs as R      # noqa: E402
from bnv_fusion_b200 import synth      # noqa: E402


def main():
    model, SparseVolume = R.build_reference(tempfile.mkdtemp())
    g = np.load(os.path.join(HERE, "golden_parity64.npz"))
    spec = synth.stream_spec("parity64")
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device="cpu")
    vol.insert(torch.from_numpy(g["recip/map_coords"]), torch.from_numpy(g["recip/map_feats"]),
               torch.from_numpy(g["recip/map_weights"]), torch.from_numpy(g["recip/map_hits"]))
    vol.to_tensor()
    vol.features = torch.nn.Parameter(vol.features)
    rng = np.random.default_rng(21)
    out = {}
    delta = torch.from_numpy(g["recip/tsdf_delta"])[None, None]
    for name, q in (("mesh", g["recip/q_mesh"][:60]), ("rand", g["recip/q_rand"][:800, None, :])):
        qt = torch.from_numpy(q)[None]
        r = torch.from_numpy(rng.standard_normal(qt.shape[1:3]).astype(np.float32))
        with R.cuda_div_semantics():
            sdf = vol.decode_pts(qt, model.nerf, delta.clone(), is_coords=True)
        loss = (sdf[0, :, :, 0] * r).sum()
        vol.features.grad = None
        loss.backward()
        out[f"{name}_q"] = q
        out[f"{name}_r"] = r.numpy()
        out[f"{name}_grad"] = vol.features.grad.numpy().copy()
        out[f"{name}_sdf"] = sdf.detach()[0, :, :, 0].numpy()
    out["coords"] = vol.active_coordinates.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_decode_grad.npz"), **out)


if __name__ == "__main__":
    main()
