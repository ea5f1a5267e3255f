"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of BNV-Fusion's per-frame dense hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / timed CPU baseline.  The product path
(`bnv_fusion_b200`) never imports it and fails loudly when its CUDA library is missing.

Pinning status: the reference ships NO tests or golden vectors for this path, so
this oracle is pinned against outputs of the reference's own UNMODIFIED Python sources executed
in the build container over fake third-party modules (`oracle/ref_stubs.py`); the vectors are
committed under `tests/golden/` with their generator (`tests/golden/make_golden.py`) and checked
by `tests/test_oracle_golden.py`.  The tiny-cuda-nn MLP semantics (un-vendored, un-pinned
dependency) are a restatement of its published FullyFusedMLP layout; the
fp16 accumulation order of the real tcnn kernels is unobservable here, so MLP parity is
"fp32 math, <=1e-4 abs SDF" as BASELINE.json states.

All citations are relative to /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- #
# grid geometry
# --------------------------------------------------------------------------- #
def get_world_range(dimensions, voxel_size):
    """src/utils/voxel_utils.py:83-88 (float64 numpy arithmetic, identical expression)."""
    dimensions = np.asarray(dimensions, dtype=np.float64)
    min_ = -dimensions / 2 - voxel_size
    max_ = dimensions / 2 + voxel_size
    n_xyz = np.ceil((max_ - min_) / voxel_size).astype(int).tolist()
    max_ = min_ + voxel_size * np.asarray(n_xyz)
    return min_, max_, n_xyz


@dataclass
class Grid:
    """SparseVolume geometry (src/models/sparse_volume.py:485-497)."""
    bmin: np.ndarray      # float32[3]  = torch.from_numpy(min_coords).float()
    bmax: np.ndarray      # float32[3]
    n_xyz: tuple          # int[3]
    voxel_size: float     # python float (double), as the reference passes it around

    @staticmethod
    def from_dimensions(dimensions, voxel_size):
        mn, mx, n = get_world_range(dimensions, voxel_size)
        return Grid(mn.astype(F32), mx.astype(F32), tuple(int(v) for v in n), float(voxel_size))


def _scalar_div(x, s, div_mode):
    """fp32 tensor / python scalar.  'recip' = PyTorch-CUDA fast path a * (1.0f/(float)s)
    (ATen BinaryDivTrueKernel.cu; SURVEY.md rule A2); 'true' = IEEE divide (PyTorch-CPU)."""
    x = np.asarray(x, dtype=F32)
    if div_mode == "recip":
        return x * (F32(1.0) / F32(s))
    if div_mode == "true":
        return x / F32(s)
    raise ValueError(div_mode)


CORNER_CEIL = np.array([  # rows k=0..7, columns axis x,y,z: 1 = ceil, 0 = floor
    [0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1],
    [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]], dtype=bool)
"""Corner order of tcnnNeRFModel.get_neighbors (src/models/fusion/modules.py:178-247) and of
fusion/utils.get_neighbors (src/models/fusion/utils.py:98-167)."""


def get_neighbors(c):
    """c: float32 [N,3] voxel-unit coords -> float32 [8,N,3] floor/ceil combinations."""
    fl = np.floor(c)
    ce = np.ceil(c)
    return np.stack([np.where(CORNER_CEIL[k][None, :], ce, fl) for k in range(8)], axis=0)


def flatten_i32(ijk, n_xyz):
    """src/utils/voxel_utils.py:62-65 evaluated on an int32 tensor (int32 * 0-dim int64 stays
    int32 in PyTorch): wraps like int32, then .long()."""
    ny, nz = int(n_xyz[1]), int(n_xyz[2])
    x = ijk[..., 0].astype(np.int64)
    y = ijk[..., 1].astype(np.int64)
    z = ijk[..., 2].astype(np.int64)
    w = lambda v: v.astype(np.int32).astype(np.int64)  # noqa: E731  int32 wrap
    a = w(w(x * ny) * nz)
    b = w(y * nz)
    return w(w(a + b) + z)


def unflatten(flat, n_xyz):
    """src/utils/voxel_utils.py:68-80."""
    ny, nz = int(n_xyz[1]), int(n_xyz[2])
    flat = np.asarray(flat, dtype=np.int64)
    x = flat // (ny * nz)
    rest = flat % (ny * nz)
    y = rest // nz
    z = flat - x * ny * nz - y * nz
    return np.stack([x, y, z], axis=-1)


# --------------------------------------------------------------------------- #
# tiny-cuda-nn FullyFusedMLP restatement
# --------------------------------------------------------------------------- #
def mlp_blocks(params, n_in, n_out, width=64, n_hidden=3):
    """Split a tcnn params vector into row-major [out,in] blocks."""
    in_pad = (n_in + 15) // 16 * 16
    out_pad = (n_out + 15) // 16 * 16
    p = np.asarray(params, dtype=F32)
    o = 0
    blocks = [p[o:o + width * in_pad].reshape(width, in_pad)]
    o += width * in_pad
    for _ in range(n_hidden - 1):
        blocks.append(p[o:o + width * width].reshape(width, width))
        o += width * width
    blocks.append(p[o:o + out_pad * width].reshape(out_pad, width))
    o += out_pad * width
    assert o == p.size, (o, p.size)
    return blocks


def _round_f16(a):
    return a.astype(np.float16).astype(np.float64)


def mlp_forward(params, x, n_in, n_out, mode="fp32"):
    """y = W3 relu(W2 relu(W1 relu(W0 [x, 1...]))) (src/utils/pointnet_utils.py:274-286,
    src/models/fusion/modules.py:171-176,249-253 over tcnn).  Accumulates in float64 from fp32
    operands ('fp32') or from fp16-rounded weights/activations ('fp16', emulating tensor-core
    operand precision with wide accumulators).  Returns float32 [N, n_out]."""
    x = np.asarray(x, dtype=F32)
    blocks = mlp_blocks(params, n_in, n_out)
    in_pad = blocks[0].shape[1]
    X = np.ones((x.shape[0], in_pad), dtype=np.float64)
    X[:, :n_in] = x
    Ws = [b.astype(np.float64) for b in blocks]
    if mode == "fp16":
        X = _round_f16(X)
        Ws = [_round_f16(w) for w in Ws]
    h = X
    for i, W in enumerate(Ws):
        h = h @ W.T
        if i < len(Ws) - 1:
            h = np.maximum(h, 0.0)
            h = _round_f16(h) if mode == "fp16" else h.astype(F32).astype(np.float64)
    return h[:, :n_out].astype(F32)


# --------------------------------------------------------------------------- #
# encode (local fusion of one frame's points)
# --------------------------------------------------------------------------- #
def bound_mask(xyz, grid):
    """Rule A1, src/models/fusion/local_point_fusion.py:94-100 (strict, fp32)."""
    vs32 = F32(grid.voxel_size)
    hi = (grid.bmax - vs32).astype(F32)
    lo = (grid.bmin + vs32).astype(F32)
    return np.all((xyz < hi[None]) & (xyz > lo[None]), axis=1)


def encode_rows(pts6, grid, div_mode="recip"):
    """Everything of encode_pointcloud up to (not including) the MLP.

    Returns dict with: keep [N] bool; flat int64 [8,n] (corner-major like the reference's
    reshape(1,-1,3) of [1,8,N,3]); mlp_in float32 [8,n,6]."""
    pts6 = np.asarray(pts6, dtype=F32)
    xyz = pts6[:, :3] * F32(1.0)
    keep = bound_mask(xyz, grid)
    xyz = xyz[keep]
    nrm = pts6[keep, 3:6]
    vs = grid.voxel_size
    # get_relative_xyz, local_point_fusion.py:153-165
    c = _scalar_div((xyz - grid.bmin[None]).astype(F32), vs, div_mode)
    nb = get_neighbors(c).astype(np.int32)                       # .int(), modules.py:247
    rel_n = (c[None] - nb.astype(F32)).astype(F32)
    rel = (rel_n * F32(vs)).astype(F32)
    # forward(normalize=True), local_point_fusion.py:58-61
    xin = _scalar_div(rel, vs, div_mode)
    mlp_in = np.concatenate([xin, np.broadcast_to(nrm[None], (8,) + nrm.shape)], axis=-1)
    flat = flatten_i32(nb, grid.n_xyz)
    return {"keep": keep, "flat": flat, "mlp_in": mlp_in.astype(F32), "corner_ijk": nb}


def encode_pointcloud(pts6, grid, enc_params, min_pts=8, div_mode="recip", mlp_mode="fp32"):
    """LitFusionPointNet.encode_pointcloud(..., return_dense=False),
    src/models/fusion/local_point_fusion.py:81-151.

    Returns (feats [M,8] f32, counts [M,1] i64, flat_ids [M] i64 ascending, coords [M,3] i64,
    n_avg f32, extra) or 5 x None + extra when no point survives the bound mask (rule A1)."""
    rows = encode_rows(pts6, grid, div_mode)
    if not rows["keep"].any():
        return None, None, None, None, None, {"touched_flat": np.zeros(0, np.int64)}
    x = rows["mlp_in"].reshape(-1, 6)
    if x[:, :3].max(initial=0) > 1 or x[:, :3].min(initial=0) < -1:
        raise AssertionError("relative xyz outside [-1,1] (local_point_fusion.py:60-61)")
    f = mlp_forward(enc_params, x, 6, 8, mlp_mode)               # rows are corner-major k*n+i
    flat = rows["flat"].reshape(-1)
    uniq, inv, cnt = np.unique(flat, return_inverse=True, return_counts=True)
    ijk = unflatten(uniq, grid.n_xyz)
    assert ijk.min() >= 0 and np.all(ijk.max(0) < np.asarray(grid.n_xyz)), "ids outside grid"
    sums = np.zeros((uniq.size, 8), dtype=np.float64)
    np.add.at(sums, inv, f.astype(np.float64))
    mean = (sums / np.maximum(cnt, 1)[:, None]).astype(F32)      # scatter_mean, :125
    n_avg = F32(cnt.astype(np.float64).mean())
    valid = cnt >= min_pts                                       # :143-147
    extra = {"touched_flat": uniq, "touched_count": cnt, "n_rows": flat.size}
    return (mean[valid], cnt[valid].astype(np.int64)[:, None], uniq[valid].astype(np.int64),
            ijk[valid].astype(np.int64), n_avg, extra)


# --------------------------------------------------------------------------- #
# sparse voxel map  (SparseVolume over o3c.HashMap; compared as a set of key -> values)
# --------------------------------------------------------------------------- #
class VoxelMap:
    """key (flat id) -> (feat[8], weight, num_hits): SparseVolume.insert/query semantics
    (src/models/sparse_volume.py:561-585,661-695)."""

    def __init__(self, grid, n_feats=8):
        self.grid = grid
        self.n_feats = n_feats
        self.index = {}
        self.feats = np.zeros((0, n_feats), F32)
        self.weights = np.zeros((0,), F32)
        self.hits = np.zeros((0,), F32)
        self.keys = np.zeros((0,), np.int64)

    def __len__(self):
        return len(self.index)

    def query(self, flat):
        flat = np.asarray(flat, np.int64).reshape(-1)
        rows = np.array([self.index.get(int(k), -1) for k in flat], dtype=np.int64)
        found = rows >= 0
        f = np.zeros((flat.size, self.n_feats), F32)
        w = np.zeros(flat.size, F32)
        h = np.zeros(flat.size, F32)
        f[found] = self.feats[rows[found]]
        w[found] = self.weights[rows[found]]
        h[found] = self.hits[rows[found]]
        return f, w, h, found

    def insert(self, flat, feats, weights, hits):
        flat = np.asarray(flat, np.int64).reshape(-1)
        new = [int(k) for k in flat if int(k) not in self.index]
        if new:
            base = len(self.index)
            for i, k in enumerate(dict.fromkeys(new)):
                self.index[k] = base + i
            n = len(self.index)
            grow = n - self.feats.shape[0]
            self.feats = np.concatenate([self.feats, np.zeros((grow, self.n_feats), F32)])
            self.weights = np.concatenate([self.weights, np.zeros(grow, F32)])
            self.hits = np.concatenate([self.hits, np.zeros(grow, F32)])
            self.keys = np.concatenate([self.keys, np.zeros(grow, np.int64)])
        rows = np.array([self.index[int(k)] for k in flat], dtype=np.int64)
        self.keys[rows] = flat
        self.feats[rows] = np.asarray(feats, F32).reshape(-1, self.n_feats)
        self.weights[rows] = np.asarray(weights, F32).reshape(-1)
        self.hits[rows] = np.asarray(hits, F32).reshape(-1)

    def as_dict(self):
        return {int(k): (self.feats[r].copy(), float(self.weights[r]), float(self.hits[r]))
                for k, r in self.index.items()}


def integrate(vmap, flat_ids, feats, counts):
    """LitFusionPointNet._integrate/_update, src/models/fusion/local_point_fusion.py:647-673
    (fp32, separately rounded mul/add/div like the reference's separate torch kernels;
    num_hits written back unchanged -- quirk A8)."""
    if flat_ids is None or len(flat_ids) == 0:
        return
    w_new = np.minimum(np.asarray(counts, np.int64).reshape(-1).astype(F32) * F32(1.0 / 32), F32(1))
    f_old, w_old, h_old, _ = vmap.query(flat_ids)
    w = (w_old + w_new).astype(F32)
    num = ((f_old * w_old[:, None]).astype(F32) + (np.asarray(feats, F32) * w_new[:, None]).astype(F32)).astype(F32)
    f = (num / w[:, None]).astype(F32)
    vmap.insert(flat_ids, f, w, h_old)


# --------------------------------------------------------------------------- #
# decode (SDF at query points)
# --------------------------------------------------------------------------- #
def positional_encoding(l):
    """[x, sin x, cos x], one frequency 2^0 (src/models/fusion/modules.py:81-123,158-162)."""
    l = np.asarray(l, F32)
    return np.concatenate([l, np.sin(l).astype(F32), np.cos(l).astype(F32)], axis=-1)


def tsdf_nearest(delta, nbr, n_xyz):
    """F.grid_sample(mode='nearest', padding_mode='zeros', align_corners=True) as driven by
    src/models/sparse_volume.py:819-832 (axis order swapped [2,1,0]; fp32 op order of ATen's
    grid_sampler_unnormalize; nearbyint = round-half-even)."""
    delta = np.asarray(delta, F32)
    T = delta.shape
    out = np.zeros(nbr.shape[:-1], F32)
    idx = []
    ok = np.ones(nbr.shape[:-1], bool)
    for ax in range(3):
        g = (nbr[..., ax].astype(F32) / F32(n_xyz[ax] - 1)).astype(F32)
        g = (g * F32(2)).astype(F32)
        g = (g - F32(1)).astype(F32)
        u = ((g + F32(1)).astype(F32) / F32(2)).astype(F32)
        u = (u * F32(T[ax] - 1)).astype(F32)
        i = np.rint(u).astype(np.int64)
        ok &= (i >= 0) & (i < T[ax])
        idx.append(np.clip(i, 0, T[ax] - 1))
    v = delta[idx[0], idx[1], idx[2]]
    out[ok] = v[ok]
    return out


def decode_pts(vmap, coords, dec_params, min_pts=8, sdf_delta=None, is_coords=True,
               div_mode="recip", mlp_mode="fp32", return_parts=False):
    """SparseVolume.decode_pts, src/models/sparse_volume.py:768-833 (rules D1-D7).

    coords: [Q,3] float32 (voxel units if is_coords else world).  sdf_delta: optional float32
    [Tx,Ty,Tz] (the reference's [1,1,Tx,Ty,Tz] squeezed).  Returns sdf float32 [Q]."""
    grid = vmap.grid
    c = np.asarray(coords, F32).reshape(-1, 3)
    if not is_coords:
        c = _scalar_div((c - grid.bmin[None]).astype(F32), grid.voxel_size, div_mode)
    nbr = get_neighbors(c)                                        # [8,Q,3] float
    l = (c[None] - nbr).astype(F32)
    if l.size and (l.min() < -1 or l.max() > 1):
        raise AssertionError("local coords outside [-1,1] (sparse_volume.py:796-797)")
    w = np.prod((F32(1) - np.abs(l)).astype(F32), axis=-1).astype(F32)   # [8,Q]
    ijk = nbr.astype(np.int64)
    inside = np.all((ijk >= 0) & (ijk < np.asarray(grid.n_xyz)[None, None]), axis=-1)
    flat = flatten_i32(np.where(inside[..., None], ijk, 0), grid.n_xyz)
    f, wt, _, found = vmap.query(flat.reshape(-1))
    found = found.reshape(8, -1) & inside
    f = np.where(found.reshape(-1)[:, None], f, F32(0)).reshape(8, -1, vmap.n_feats)
    wt = np.where(found, wt.reshape(8, -1), F32(0))
    mask = wt.min(axis=0) >= F32(min_pts)                        # D3
    x = np.concatenate([positional_encoding(l), f], axis=-1).reshape(-1, 9 + vmap.n_feats)
    y = mlp_forward(dec_params, x, 9 + vmap.n_feats, 1, mlp_mode).reshape(8, -1)
    s = (y * F32(grid.voxel_size)).astype(F32)                   # D4
    wn = (w / w.sum(axis=0, keepdims=True).astype(F32)).astype(F32)      # D2
    sdf = (s.astype(np.float64) * wn).sum(axis=0).astype(F32)    # D5
    sdf = np.where(mask, sdf, F32(0) + F32(grid.voxel_size)).astype(F32)
    if sdf_delta is not None:                                    # D6
        d = tsdf_nearest(sdf_delta, nbr, grid.n_xyz)
        sdf = (sdf + (d.astype(np.float64) * wn).sum(axis=0).astype(F32)).astype(F32)
    if return_parts:
        return sdf, {"mask": mask, "wn": wn, "y": y, "flat": flat, "found": found}
    return sdf


def meshlize_samples(active_ijk):
    """The sampling half of SparseVolume.meshlize (src/models/sparse_volume.py:717-731):
    27 voxel-unit coordinates id + {-0.5, 0, 0.5}^3 per active voxel, 'ij' meshgrid order."""
    r = np.arange(0, 1.5, 0.5) - 0.5
    off = np.stack(np.meshgrid(r, r, r, indexing="ij"), axis=-1).reshape(27, 3)
    return (np.asarray(active_ijk, np.float64)[:, None, :] + off[None]).astype(F32)


# --------------------------------------------------------------------------- #
# back-projection (dataset side, float64 on the CPU in the reference)
# --------------------------------------------------------------------------- #
def load_depth_u16(depth_u16, max_depth=None):
    """src/utils/common.py:86-120 after cv2.imread: /1000., mask = 0<d(<max_depth), d*=mask."""
    depth = np.asarray(depth_u16).astype(np.float64) / 1000.0
    mask = depth > 0
    if max_depth is not None:
        mask = mask & (depth < max_depth)
        depth = depth * mask
    return depth, mask


def depth2xyz(depth, K):
    """src/utils/geometry.py:150-171: (u-cx)/fx evaluated in float32, promoted to float64 by
    the np.ones stack, times float64 depth."""
    K = np.asarray(K, F32)
    h, w = depth.shape
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    ur = ((np.arange(w, dtype=F32) - cx) / fx).astype(F32)
    vr = ((np.arange(h, dtype=F32) - cy) / fy).astype(F32)
    xyz = np.stack([np.broadcast_to(ur[None, :], (h, w)).astype(np.float64),
                    np.broadcast_to(vr[:, None], (h, w)).astype(np.float64),
                    np.ones((h, w))], axis=-1)
    return xyz * depth[:, :, None]


def depth_to_normals(depth, K):
    """kornia.geometry.depth.depth_to_normals (kornia 0.6.2, environment.yml:439; un-vendored)
    as called at src/datasets/fusion_inference_dataset.py:52-55 with a float64 depth and a
    float32 K: unproject ((u-cx)/fx in float64), Sobel/8 with replicate padding per xyz channel,
    cross(d/du, d/dv), L2-normalise with eps 1e-12.  Summation order here (row-major over the
    3x3 taps, zero taps skipped) is the order the CUDA kernel uses."""
    K = np.asarray(K, F32).astype(np.float64)
    h, w = depth.shape
    u = np.arange(w, dtype=np.float64)
    v = np.arange(h, dtype=np.float64)
    x = ((u - K[0, 2]) / K[0, 0])[None, :] * depth
    y = ((v - K[1, 2]) / K[1, 1])[:, None] * depth
    xyz = np.stack([x, y, depth], axis=0)                        # [3,H,W]
    p = np.pad(xyz, ((0, 0), (1, 1), (1, 1)), mode="edge")

    def tap(dy, dx):
        return p[:, 1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    e = 1.0 / 8.0
    gx = ((((((-e) * tap(-1, -1)) + e * tap(-1, 1)) + (-2 * e) * tap(0, -1)) + (2 * e) * tap(0, 1))
          + (-e) * tap(1, -1)) + e * tap(1, 1)
    gy = ((((((-e) * tap(-1, -1)) + (-2 * e) * tap(-1, 0)) + (-e) * tap(-1, 1)) + e * tap(1, -1))
          + (2 * e) * tap(1, 0)) + e * tap(1, 1)
    n = np.stack([gx[1] * gy[2] - gx[2] * gy[1],
                  gx[2] * gy[0] - gx[0] * gy[2],
                  gx[0] * gy[1] - gx[1] * gy[0]], axis=0)
    nn = np.sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2])
    n = n / np.maximum(nn, 1e-12)[None]
    return np.transpose(n, (1, 2, 0))                            # [H,W,3]


def backproject(depth, mask, K, T_wc):
    """src/datasets/fusion_inference_dataset.py:52-74 + run_e2e.py:247-249: world points and
    world normals of the masked pixels in row-major pixel order, float64 maths, then .float()."""
    T = np.asarray(T_wc, F32).astype(np.float64)
    pts_c = depth2xyz(depth, K).reshape(-1, 3)
    pw = ((pts_c[:, 0:1] * T[None, :3, 0] + pts_c[:, 1:2] * T[None, :3, 1])
          + pts_c[:, 2:3] * T[None, :3, 2]) + T[None, :3, 3]
    n = depth_to_normals(depth, K).reshape(-1, 3)
    nw = (n[:, 0:1] * T[None, :3, 0] + n[:, 1:2] * T[None, :3, 1]) + n[:, 2:3] * T[None, :3, 2]
    out = np.concatenate([pw, nw], axis=-1)[np.asarray(mask).reshape(-1)]
    return out.astype(F32)


# --------------------------------------------------------------------------- #
# decode backward (global optimisation)
# --------------------------------------------------------------------------- #
def decode_pts_backward(vmap, coords, dec_params, grad_out, min_pts=8, is_coords=True, div_mode="recip"):
    """d(sum(grad_out * sdf)) / d(features) for SparseVolume.decode_pts (sparse_volume.py:768-833) as
    autograd computes it in NeuralMap.optimize (src/run_e2e.py:111-156; the features are the only leaf).
    Returns {flat_id: grad[8]} (float64 accumulation).  The TSDF prior and the `voxel_size` fallback of
    masked queries do not depend on the features."""
    grid = vmap.grid
    c = np.asarray(coords, F32).reshape(-1, 3)
    if not is_coords:
        c = _scalar_div((c - grid.bmin[None]).astype(F32), grid.voxel_size, div_mode)
    g_out = np.asarray(grad_out, np.float64).reshape(-1)
    nbr = get_neighbors(c)
    l = (c[None] - nbr).astype(F32)
    w = np.prod((F32(1) - np.abs(l)).astype(F32), axis=-1).astype(np.float64)
    ijk = nbr.astype(np.int64)
    inside = np.all((ijk >= 0) & (ijk < np.asarray(grid.n_xyz)[None, None]), axis=-1)
    flat = flatten_i32(np.where(inside[..., None], ijk, 0), grid.n_xyz)
    f, wt, _, found = vmap.query(flat.reshape(-1))
    found = found.reshape(8, -1) & inside
    f = np.where(found.reshape(-1)[:, None], f, F32(0))
    wt = np.where(found, wt.reshape(8, -1), F32(0))
    mask = wt.min(axis=0) >= F32(min_pts)
    wn = w / w.sum(axis=0, keepdims=True)
    x = np.concatenate([positional_encoding(l).reshape(-1, 9), f], axis=-1).astype(np.float64)
    blocks = [b.astype(np.float64) for b in mlp_blocks(dec_params, 17, 1)]
    X = np.ones((x.shape[0], 32))
    X[:, :17] = x
    h1 = X @ blocks[0].T
    h2 = np.maximum(h1, 0) @ blocks[1].T
    h3 = np.maximum(h2, 0) @ blocks[2].T
    dy = (g_out[None] * wn * grid.voxel_size * mask[None]).reshape(-1)          # d loss / d y_k
    d3 = (dy[:, None] * blocks[3][0][None]) * (h3 > 0)
    d2 = (d3 @ blocks[2]) * (h2 > 0)
    d1 = (d2 @ blocks[1]) * (h1 > 0)
    dfeat = (d1 @ blocks[0])[:, 9:17]
    out = {}
    fl = flat.reshape(-1)
    fd = found.reshape(-1)
    for i in np.nonzero(fd & (dy != 0))[0]:
        k = int(fl[i])
        out[k] = out.get(k, 0) + dfeat[i]
    return out
