"""Coarse TSDF prior: oracle pinned to the reference's own CPU-mode output (CPU test), CUDA kernel vs the
same goldens (GPU test)."""
import os

import numpy as np
import pytest

from oracle import bnv_oracle as O
from oracle.tsdf_oracle import TSDFOracle, prepare_tsdf_volume
from bnv_fusion_b200 import synth


def _frames():
    spec = synth.stream_spec("parity64")
    out = []
    for fi in range(6):
        d, K, T = synth.make_frame(spec, fi, seed=0)
        depth, _ = O.load_depth_u16(d, spec.max_depth)
        out.append((d, depth.astype(np.float32), K, T))
    return out


@pytest.mark.parametrize("tag,vs", [("v25", 0.025), ("v10", 0.01)])
def test_tsdf_oracle_matches_reference(golden_dir, tag, vs):
    g = np.load(os.path.join(golden_dir, "golden_tsdf.npz"))
    vol = TSDFOracle(g[f"{tag}/bnds"], vs)
    for fi, (_, depth, K, T) in enumerate(_frames()):
        vol.integrate(g[f"{tag}/rgb{fi}"].astype(np.float32), depth, K, T, 1.0)
    t, c = vol.get_volume()
    assert np.array_equal(t, g[f"{tag}/tsdf"]) and np.array_equal(vol.weight, g[f"{tag}/weight"])
    assert np.array_equal(c, g[f"{tag}/color"])
    assert (vol.weight > 0).sum() > 1000 and t.min() < -0.5 and t.max() == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("tag,vs", [("v25", 0.025), ("v10", 0.01)])
@pytest.mark.parametrize("depth_kind", ["float", "u16"])
def test_tsdf_gpu_matches_reference(golden_dir, tag, vs, depth_kind):
    import torch
    from bnv_fusion_b200.tsdf import TSDFVolume
    g = np.load(os.path.join(golden_dir, "golden_tsdf.npz"))
    vol = TSDFVolume(g[f"{tag}/bnds"], vs)
    assert tuple(vol._vol_dim) == g[f"{tag}/tsdf"].shape
    for fi, (d16, depth, K, T) in enumerate(_frames()):
        vol.integrate(g[f"{tag}/rgb{fi}"].astype(np.float32), depth if depth_kind == "float" else d16, K, T, 1.0)
    t, c = vol.get_volume()
    # (1) bit-exact against the oracle (pinned to the reference's output by the CPU test above) executed on THIS
    # host: both sides then use the same np.linalg.inv(cam_pose), the one platform-dependent step of fusion.py:254
    orc = TSDFOracle(g[f"{tag}/bnds"], vs)
    for fi, (_, depth, K, T) in enumerate(_frames()):
        orc.integrate(g[f"{tag}/rgb{fi}"].astype(np.float32), depth, K, T, 1.0)
    ot, oc = orc.get_volume()
    assert np.array_equal(t, ot), float(np.abs(t - ot).max())
    assert np.array_equal(c, oc)
    assert np.array_equal(vol._view(2).cpu().numpy(), orc.weight)
    # (2) against the goldens minted in the build container: LAPACK's float32 inverse may differ in the last bit
    # between hosts and flip a half-way pixel rounding for a handful of voxels
    ref = g[f"{tag}/tsdf"]
    bad = np.abs(t - ref) > 1e-6
    assert bad.mean() < 2e-3, bad.mean()
    assert np.abs(t - ref)[~bad].max() <= 1e-6
    assert (c != g[f"{tag}/color"]).mean() < 2e-3
    # prepare_tsdf_volume on the device == the reference formula on the reference volume
    p = vol.prior(0.05, 0.1)
    assert p.shape == (1, 1) + ref.shape
    want = prepare_tsdf_volume(t, vs, 0.05, 0.1)
    assert np.array_equal(p[0, 0].cpu().numpy(), want)
