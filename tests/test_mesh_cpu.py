"""Mesh extraction, CPU side: the generated marching-cubes table and the numpy oracle, pinned on
analytic SDFs (scikit-image, whose marching_cubes the reference calls, is absent from this image: parity with its exact
triangle lists is unpinned, see oracle/mesh_oracle.py)."""
import os

import numpy as np

from bnv_fusion_b200 import mc_tables as T
from oracle import mesh_oracle as MO


def test_table_file_is_current():
    assert open(T.INC_PATH).read() == T.render_inc(), "run `python -m bnv_fusion_b200.mc_tables`"


def test_table_properties():
    assert T.MAX_TRI == 5 and T.TRI_COUNT[0] == 0 and T.TRI_COUNT[255] == 0
    for case in range(256):
        tris = T.TRI_TABLE[case][: 3 * T.TRI_COUNT[case]].reshape(-1, 3)
        inside = [(case >> c) & 1 for c in range(8)]
        crossed = {e for e in range(12) if inside[T.EDGE_CORNERS[e][0]] != inside[T.EDGE_CORNERS[e][1]]}
        assert set(tris.reshape(-1).tolist()) == crossed          # every crossed edge carries a vertex, no other does
        # inside the cube every triangle edge is either shared by two triangles (opposite directions) or lies on a face
        half = {}
        for a, b, c in tris.tolist():
            for u, v in ((a, b), (b, c), (c, a)):
                half[(u, v)] = half.get((u, v), 0) + 1
        assert all(n == 1 for n in half.values())


def _sample_blocks(fn, coords):
    off = np.array([-0.5, 0.0, 0.5], np.float32)
    g = np.stack(np.meshgrid(off, off, off, indexing="ij"), -1)                     # [3,3,3,3]
    pts = coords[:, None, None, None, :].astype(np.float32) + g[None]
    return fn(pts).astype(np.float32)


def _manifold_checks(verts, keys):
    v, f = MO.weld(verts, keys)
    half = {}
    for a, b, c in f.tolist():
        assert a != b and b != c and a != c
        for e in ((a, b), (b, c), (c, a)):
            assert e not in half, "two triangles run along the same edge in the same direction"
            half[e] = 1
    assert all((b, a) in half for (a, b) in half), "open edge: the surface has a crack"
    return v, f


def test_oracle_sphere_is_closed_oriented_and_on_the_surface():
    n, c, r = 24, np.array([11.3, 12.1, 11.8], np.float32), 7.4
    ijk = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3)
    sdf = _sample_blocks(lambda p: np.linalg.norm(p - c, axis=-1) - r, ijk)
    vs, mn = 0.01, np.array([-0.13, -0.12, -0.11], np.float32)
    verts, keys, owner = MO.marching_blocks(sdf, ijk, vs, mn, (n, n, n))
    assert len(verts) > 3000 and len(verts) % 3 == 0
    v, f = _manifold_checks(verts, keys)
    assert len(v) - 3 * len(f) // 2 + len(f) == 2                                    # Euler characteristic of a sphere
    vox = (v - mn) / np.float32(vs)
    assert np.abs(np.linalg.norm(vox - c, axis=1) - r).max() < 0.02                 # linear interpolation error, voxels
    tri = vox[f]
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert (np.einsum("ij,ij->i", nrm, tri.mean(1) - c) > 0).all()                  # normals point to sdf > 0
    area = 0.5 * np.linalg.norm(nrm, axis=1).sum()
    assert abs(area / (4 * np.pi * r * r) - 1) < 0.02


def test_oracle_plane_torus_and_block_rule():
    n = 12
    ijk = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3)
    nrm = np.array([0.3, -0.5, 0.81], np.float32); nrm /= np.linalg.norm(nrm)
    sdf = _sample_blocks(lambda p: (p - 5.6) @ nrm, ijk)
    verts, keys, owner = MO.marching_blocks(sdf, ijk, 1.0, np.zeros(3, np.float32), (n, n, n))
    assert np.abs((verts - 5.6) @ nrm).max() < 1e-5                                  # exact for a linear field
    v, f = MO.weld(verts, keys)
    assert len(v) < len(verts) // 2
    # the reference's block rule (sparse_volume.py:740): no sign change inside the block -> nothing, even when a
    # sample is exactly 0
    blk = np.full((1, 3, 3, 3), -1.0, np.float32); blk[0, 1, 1, 1] = 0.0
    assert len(MO.marching_blocks(blk, np.array([[3, 3, 3]]), 1.0, np.zeros(3), (n, n, n))[0]) == 0
    # torus: genus 1 -> Euler characteristic 0, closed and oriented
    n = 20
    ijk = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3)
    def torus(p):
        q = p - 9.7
        return np.sqrt((np.sqrt(q[..., 0] ** 2 + q[..., 1] ** 2) - 5.5) ** 2 + q[..., 2] ** 2) - 2.2
    verts, keys, _ = MO.marching_blocks(_sample_blocks(torus, ijk), ijk, 1.0, np.zeros(3, np.float32), (n, n, n))
    v, f = _manifold_checks(verts, keys)
    assert len(v) - 3 * len(f) // 2 + len(f) == 0
