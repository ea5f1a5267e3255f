"""Mint the golden vectors of tests/golden/ by running the UNMODIFIED reference sources
(/root/reference) on CPU over the fakes of oracle/ref_stubs.py.  Runs only in the build
container (the GPU box has no /root/reference); the outputs are committed.

    python tests/golden/make_golden.py

Files written next to this script:
  tcnn_params.npz          the two flat parameter tensors of pretrained/pointnet_tcnn.ckpt
  golden_parity64.npz      24 frames of the 64x64 / 32^3 workload through
                           encode_pointcloud -> _integrate -> to_tensor -> decode_pts
  golden_lounge_crop.npz   2 frames of a 160x120 crop of the 640x480 / 512^3 workload
  golden_edge.npz          integral-coordinate / out-of-bounds / empty-frame cases + MLP KATs

Every case is stored twice where division semantics matter: `true` (reference as executed on
CPU: IEEE divide) and `recip` (tensor / python-scalar evaluated as x * (1.0f/s), the PyTorch-CUDA
fast path the reference takes on a GPU, forced with ref_stubs.cuda_div_semantics).
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import bnv_oracle as O          # noqa: E402  (normals restatement only, see below)
from oracle import ref_stubs as R           # noqa: E402
from bnv_fusion_b200 import synth           # noqa: E402


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def div_ctx(mode):
    return R.cuda_div_semantics() if mode == "recip" else _nullctx()


def ref_backproject(depth_u16, K, T_wc, max_depth):
    """xyz through the reference's own geometry code; normals through the kornia restatement
    (kornia is un-vendored and absent: that part is NOT pinned by the reference)."""
    import src.utils.geometry as geometry
    depth = depth_u16.astype(np.float64) / 1000.0            # cv2.imread(path,-1)/1000.
    mask = depth > 0
    mask = mask * (depth < max_depth)
    depth = depth * mask
    mask = mask.astype(bool)
    pts_c = geometry.depth2xyz(depth, K).reshape(-1, 3)
    pts_w = (T_wc @ geometry.get_homogeneous(pts_c).T)[:3, :].T
    normal = O.depth_to_normals(depth, K)
    normal_w = (T_wc[:3, :3] @ normal.reshape(-1, 3).T).T
    input_pts = np.concatenate([pts_w, normal_w], axis=-1)[mask.reshape(-1)]
    # run_e2e.py:247-249  frame[k].cuda().float()
    return torch.from_numpy(input_pts).float().numpy(), mask


def run_stream(model, SparseVolume, spec, frames, mode, weight_scale=None):
    out = {}
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device="cpu")
    out["n_xyz"] = vol.n_xyz.numpy()
    out["bmin"] = vol.min_coords.numpy()
    out["bmax"] = vol.max_coords.numpy()
    for fi, (depth, K, T) in enumerate(frames):
        pts6, _ = ref_backproject(depth, K, T, spec.max_depth)
        out[f"f{fi}_pts6"] = pts6
        with torch.no_grad(), div_ctx(mode):
            res = model.encode_pointcloud(torch.from_numpy(pts6)[None].clone(), vol.n_xyz,
                                          vol.min_coords, vol.max_coords, vol.voxel_size,
                                          return_dense=False)
            feats, counts, flat, coords, navg = res
            if feats is None:
                out[f"f{fi}_empty"] = np.array(1)
                continue
            model._integrate(vol, coords, feats, counts)
        out[f"f{fi}_feats"] = feats.numpy()
        out[f"f{fi}_counts"] = counts.numpy()
        out[f"f{fi}_flat"] = flat.numpy()
        out[f"f{fi}_coords"] = coords.numpy()
        out[f"f{fi}_navg"] = np.asarray(float(navg), np.float32)
    vol.to_tensor()
    out["map_coords"] = vol.active_coordinates.numpy()
    out["map_feats"] = vol.features.numpy()
    out["map_weights"] = vol.weights.numpy().copy()
    out["map_hits"] = vol.num_hits.numpy()
    if weight_scale is not None:
        # too few frames for any fusion weight to reach min_pts_in_grid: scale the compacted
        # weights (state manipulation outside the reference code, mirrored by the tests) so the
        # decode goldens exercise the valid-mask branch as well
        vol.weights *= weight_scale
        out["weight_scale"] = np.asarray(weight_scale, np.float32)
    # decode: meshlize samples of a subset of active voxels + random in-grid coordinates
    rng = np.random.default_rng(7)
    A = vol.active_coordinates.shape[0]
    heavy = np.nonzero(vol.weights.numpy()[:, 0] >= 10.0)[0]
    sel = np.concatenate([rng.choice(heavy, size=min(heavy.size, 120), replace=False),
                          rng.choice(A, size=min(A, 40), replace=False)])
    q_mesh = O.meshlize_samples(vol.active_coordinates.numpy()[sel])      # [S,27,3]
    lo = vol.active_coordinates.numpy().min(0) - 1.5
    hi = vol.active_coordinates.numpy().max(0) + 1.5
    q_rand = (lo + (hi - lo) * rng.random((1500, 3))).astype(np.float32)
    q_rand = np.clip(q_rand, 0.0, np.asarray(out["n_xyz"], np.float32) - 1.0)
    tsdf = (rng.standard_normal((13, 14, 15)) * 0.004).astype(np.float32)
    out["q_mesh"] = q_mesh
    out["q_rand"] = q_rand
    out["tsdf_delta"] = tsdf
    delta_t = torch.from_numpy(tsdf)[None, None]
    with torch.no_grad(), div_ctx(mode):
        out["sdf_mesh"] = vol.decode_pts(torch.from_numpy(q_mesh)[None], model.nerf, None,
                                         is_coords=True)[0, :, :, 0].numpy()
        out["sdf_mesh_prior"] = vol.decode_pts(torch.from_numpy(q_mesh)[None], model.nerf,
                                               delta_t.clone(), is_coords=True)[0, :, :, 0].numpy()
        out["sdf_rand_prior"] = vol.decode_pts(torch.from_numpy(q_rand)[None, :, None, :], model.nerf,
                                               delta_t.clone(), is_coords=True)[0, :, 0, 0].numpy()
        # world-coordinate entry (is_coords=False) on the same random points
        q_world = q_rand * np.float32(spec.voxel_size) + out["bmin"][None]
        out["q_world"] = q_world.astype(np.float32)
        out["sdf_world"] = vol.decode_pts(torch.from_numpy(out["q_world"])[None, :, None, :],
                                          model.nerf, None, is_coords=False)[0, :, 0, 0].numpy()
    return out


def main():
    work = tempfile.mkdtemp(prefix="bnv_golden_")
    model, SparseVolume = R.build_reference(work, voxel_size=0.01, min_pts=8, mlp_mode="fp32")
    sd = R.load_ckpt_state_dict("pointnet_tcnn.ckpt")
    np.savez(os.path.join(HERE, "tcnn_params.npz"),
             encoder=sd["pointnet_backbone.model.params"].numpy(),
             decoder=sd["nerf.model.params"].numpy())

    # ---------------- parity64: 24 frames (4 poses x 6, fresh noise each) ------------------
    spec = synth.stream_spec("parity64")
    frames = [synth.make_frame(spec, i, seed=0) for i in range(24)]
    pack = {"depth": np.stack([f[0] for f in frames]), "K": np.stack([f[1] for f in frames]),
            "T_wc": np.stack([f[2] for f in frames])}
    for mode in ("true", "recip"):
        res = run_stream(model, SparseVolume, spec, frames, mode)
        keep = {k: v for k, v in res.items()
                if not k.startswith("f") or k.split("_")[0] in ("f0", "f3", "f23")}
        pack.update({f"{mode}/{k}": v for k, v in keep.items()})
    np.savez_compressed(os.path.join(HERE, "golden_parity64.npz"), **pack)
    print("parity64:", {k: v.shape for k, v in pack.items() if k.startswith("recip/map")})

    # ---------------- lounge crop: 160x120 window of the 640x480 stream, 512^3 grid --------
    spec = synth.stream_spec("lounge")
    frames = []
    for i in (0, 5):
        d, K, T = synth.make_frame(spec, i, seed=0)
        y0, x0 = 180, 240
        d = np.ascontiguousarray(d[y0:y0 + 120, x0:x0 + 160])
        K = K.copy()
        K[0, 2] -= x0
        K[1, 2] -= y0
        frames.append((d, K, T))
    pack = {"depth": np.stack([f[0] for f in frames]), "K": np.stack([f[1] for f in frames]),
            "T_wc": np.stack([f[2] for f in frames])}
    res = run_stream(model, SparseVolume, spec, frames, "recip", weight_scale=6.0)
    res = {k: v for k, v in res.items() if not k.endswith("_pts6") or k.startswith("f0")}
    pack.update({f"recip/{k}": v for k, v in res.items()})
    np.savez_compressed(os.path.join(HERE, "golden_lounge_crop.npz"), **pack)
    print("lounge_crop:", {k: v.shape for k, v in pack.items() if "map" in k or "flat" in k})

    # ---------------- edge cases + MLP known answers ----------------------------------------
    spec = synth.stream_spec("parity64")
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device="cpu")
    bmin = vol.min_coords.numpy()
    bmax = vol.max_coords.numpy()
    # points whose voxel coordinate is exactly integral under the recip form (rule A3)
    cand = (bmin[0] + np.arange(2, 30, dtype=np.float32) * np.float32(0.01)).astype(np.float32)
    c = ((cand - bmin[0]).astype(np.float32) * np.float32(100.0)).astype(np.float32)
    integral = cand[c == np.rint(c)]
    assert integral.size >= 3, integral.size
    rng = np.random.default_rng(3)
    pts = []
    a, b, cc = integral[:3]
    pts += [[a, a, a]] * 2                 # triple-integral: 8 rows into one voxel per point
    pts += [[b, b, -0.003]] * 5            # double-integral
    pts += [[cc, 0.0123, 0.0456]] * 9      # single-integral
    pts += (rng.random((400, 3)) * 0.05 + 0.01).tolist()       # a dense blob
    pts += [[bmax[0], 0, 0], [bmin[0], 0, 0], [0.5, 0.5, 0.5], [bmax[0] - 0.01, 0, 0],
            [bmin[0] + 0.01, 0, 0], [float(np.nextafter(bmax[0] - np.float32(0.01), np.float32(-1))), 0, 0]]
    pts = np.asarray(pts, np.float32)
    nrm = rng.standard_normal((pts.shape[0], 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pts6 = np.concatenate([pts, nrm.astype(np.float32)], axis=1).astype(np.float32)
    pack = {"pts6": pts6, "bmin": bmin, "bmax": bmax, "n_xyz": vol.n_xyz.numpy()}
    for mode in ("true", "recip"):
        with torch.no_grad(), div_ctx(mode):
            feats, counts, flat, coords, navg = model.encode_pointcloud(
                torch.from_numpy(pts6)[None].clone(), vol.n_xyz, vol.min_coords, vol.max_coords,
                vol.voxel_size, return_dense=False)
        pack.update({f"{mode}/feats": feats.numpy(), f"{mode}/counts": counts.numpy(),
                     f"{mode}/flat": flat.numpy(), f"{mode}/coords": coords.numpy(),
                     f"{mode}/navg": np.asarray(float(navg), np.float32)})
    far = pts6.copy()
    far[:, :3] += 5.0
    with torch.no_grad():
        r = model.encode_pointcloud(torch.from_numpy(far)[None], vol.n_xyz, vol.min_coords,
                                    vol.max_coords, vol.voxel_size, return_dense=False)
    assert all(v is None for v in r)
    pack["far_pts6"] = far
    # MLP known answers through the reference's own module wrappers
    xe = rng.uniform(-1, 1, size=(64, 6)).astype(np.float32)
    xd = np.concatenate([rng.uniform(-1, 1, size=(64, 3)), rng.normal(0, 0.8, size=(64, 8))], 1).astype(np.float32)
    with torch.no_grad():
        ye = model.pointnet_backbone(torch.from_numpy(xe).t()[None], False)[0].t().numpy()
        geo_in = torch.cat([model.nerf.xyz_encoding(torch.from_numpy(xd[:, :3])),
                            torch.from_numpy(xd[:, 3:])], dim=-1)
        yd = model.nerf.geo_forward(geo_in).numpy()
    pack.update({"kat_enc_x": xe, "kat_enc_y": ye, "kat_dec_x": xd, "kat_dec_y": yd})
    np.savez_compressed(os.path.join(HERE, "golden_edge.npz"), **pack)
    print("edge:", pack["recip/flat"].shape, pack["recip/counts"].reshape(-1)[:8])
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
