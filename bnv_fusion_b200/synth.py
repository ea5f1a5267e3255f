"""Synthetic depth streams shaped like the reference's datasets (there is no network: scene3d /
ScanNet / ARKit data cannot be downloaded).  Host-side numpy only.

A frame is what the reference's FusionInferenceDataset hands to run_e2e.py
(/root/reference/src/datasets/fusion_inference_dataset.py:75-90) *before* back-projection:
uint16 millimetre depth (what cv2.imread returns, src/utils/common.py:93), float32 K [3,3],
float32 T_wc [4,4].

Workloads:
  * "parity64"  : 64x64 depth, K = (64,64,32,32), smooth surface + 1 mm noise, 0.30 m cube
                  -> 32^3 grid at 1 cm (BASELINE.json configs[0]).
  * "lounge"    : 640x480, K = (525,525,319.5,239.5), camera orbiting inside a room of planes
                  and spheres, depths 0.5-3 m, 5.1 m cube -> 512^3 grid at 1 cm (configs[1]).
  * "arkit"     : 256x192, K scaled accordingly, 2 cm voxels, 10 % random invalidation.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class StreamSpec:
    name: str
    height: int
    width: int
    K: np.ndarray            # float32 [3,3]
    dimensions: np.ndarray   # float64 [3] scene extent in metres
    voxel_size: float
    max_depth: float


def stream_spec(name: str) -> StreamSpec:
    if name == "parity64":
        K = np.array([[64, 0, 32], [0, 64, 32], [0, 0, 1]], np.float32)
        return StreamSpec(name, 64, 64, K, np.asarray([0.30] * 3), 0.01, 3.0)
    if name == "lounge":
        K = np.array([[525, 0, 319.5], [0, 525, 239.5], [0, 0, 1]], np.float32)
        return StreamSpec(name, 480, 640, K, np.asarray([5.1] * 3), 0.01, 3.0)
    if name == "arkit":
        K = np.array([[212.0, 0, 127.5], [0, 212.0, 95.5], [0, 0, 1]], np.float32)
        return StreamSpec(name, 192, 256, K, np.asarray([5.1] * 3), 0.02, 3.0)
    raise ValueError(name)


def _look_at(pos, fwd):
    fwd = fwd / np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = right, down, fwd, pos
    return T.astype(np.float32)


def _room_depth(spec: StreamSpec, T_wc, rng):
    """Ray-cast a box room (walls at +-2.2 m, floor/ceiling at -+1.2 m) with five spheres."""
    h, w = spec.height, spec.width
    K = spec.K.astype(np.float64)
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d_c = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)], -1)
    R = T_wc[:3, :3].astype(np.float64)
    o = T_wc[:3, 3].astype(np.float64)
    d = d_c @ R.T                                    # world ray dirs, z_cam component == 1
    lo = np.array([-2.2, -2.2, -1.2])
    hi = np.array([2.2, 2.2, 1.2])
    with np.errstate(divide="ignore", invalid="ignore"):
        t_hi = (hi - o) / d
        t_lo = (lo - o) / d
    t_exit = np.where(d > 0, t_hi, t_lo)
    t_exit = np.where(np.abs(d) < 1e-12, np.inf, t_exit)
    t = t_exit.min(axis=-1)
    spheres = [((1.4, 0.6, -0.7), 0.5), ((-1.2, 1.3, -0.8), 0.4), ((0.3, -1.6, -0.6), 0.6),
               ((-1.5, -1.0, 0.2), 0.35), ((1.0, -0.4, 0.5), 0.3)]
    for c, r in spheres:
        oc = o - np.asarray(c)
        a = (d * d).sum(-1)
        b = 2 * (d * oc).sum(-1)
        cc = (oc * oc).sum() - r * r
        disc = b * b - 4 * a * cc
        ts = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
        ts = np.where(ts > 1e-3, ts, np.inf)
        t = np.minimum(t, ts)
    depth = t                                        # camera-z depth because d_c.z == 1
    depth = depth + rng.normal(0.0, 0.0015, size=depth.shape)
    mm = np.clip(np.rint(depth * 1000.0), 0, 65535)
    mm = np.where(np.isfinite(depth), mm, 0)
    return mm.astype(np.uint16)


def make_frame(spec: StreamSpec, index: int, seed: int = 0):
    """Returns (depth_u16 [H,W] uint16 millimetres, K float32 [3,3], T_wc float32 [4,4])."""
    rng = np.random.default_rng([seed, index])
    if spec.name == "parity64":
        h, w = spec.height, spec.width
        u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
        depth = 0.35 + 0.02 * np.sin(u / 9.0) + 0.02 * np.cos(v / 7.0)
        depth = depth + rng.normal(0.0, 0.001, size=depth.shape)
        T = np.eye(4, dtype=np.float32)
        shift = np.array([[0, 0, 0], [0.02, 0, 0], [0, -0.02, 0], [-0.02, 0.02, 0.01]])[index % 4]
        # surface spans about +-0.18 m laterally at 0.35 m; put it in the middle of the box
        T[:3, 3] = np.array([0.0, 0.0, -0.35]) + shift
        return np.rint(depth * 1000.0).astype(np.uint16), spec.K.copy(), T
    th = 2 * np.pi * (index % 200) / 200.0
    pos = np.array([0.6 * np.cos(th), 0.6 * np.sin(th), 0.15 * np.sin(2 * th)])
    fwd = np.array([np.cos(th + 0.35), np.sin(th + 0.35), -0.25 + 0.1 * np.sin(3 * th)])
    T = _look_at(pos, fwd)
    depth = _room_depth(spec, T, rng)
    if spec.name == "arkit":                         # confidence_level >= 2 emulation
        depth = np.where(rng.random(depth.shape) < 0.10, 0, depth).astype(np.uint16)
    return depth, spec.K.copy(), T
