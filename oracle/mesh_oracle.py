"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the mesh-extraction half of SparseVolume.meshlize
(/root/reference/src/models/sparse_volume.py:738-766): marching cubes at level 0 over each active voxel's 3x3x3
block of SDF samples (spacing 0.5 voxel), vertices moved to `coord - 0.5`, scaled by voxel_size, shifted by min_coords.

PARITY UNPINNED against the reference's triangulation: the reference calls skimage.measure.marching_cubes (Lewiner),
a third-party routine absent from this image and from /root/reference, so its exact triangle lists cannot be executed
here.  What is restated is the published algorithm family (linear interpolation of the level-0 crossing on every cube
edge whose end points differ in sign; the block is meshed only if max > 0 and min < 0, like :740); the case table is
bnv_fusion_b200/mc_tables.py (generated, crack-free by construction).  tests/test_mesh_cpu.py pins this oracle on
analytic SDFs (closed 2-manifold, consistent outward orientation, vertices on the analytic surface); the CUDA kernel
(csrc/bnv_mesh.cu) is then held to this oracle bit for bit (vertex keys) / to 1 ulp-level tolerance (positions).
"""
import numpy as np

from bnv_fusion_b200 import mc_tables as T

F32 = np.float32


def marching_blocks(sdf_blocks, coords, voxel_size, min_coords, n_xyz):
    """sdf_blocks [A,3,3,3] float32 ('ij' order: index = offset / 0.5 + 1 per axis), coords [A,3] int.
    Returns verts [3T,3] float32 (world units, triangle t = rows 3t..3t+2), keys [3T] int64 (id of the half-voxel
    lattice edge a vertex lies on: equal keys <=> same vertex), tri_voxel [T] (index of the generating voxel).
    Order: voxels ascending, sub-cubes (a,b,c) with a slowest, table order."""
    sdf = np.asarray(sdf_blocks, F32).reshape(-1, 3, 3, 3)
    coords = np.asarray(coords, np.int64).reshape(-1, 3)
    vs, mn = F32(voxel_size), np.asarray(min_coords, F32)
    active = (sdf.reshape(len(sdf), -1).max(1) > 0) & (sdf.reshape(len(sdf), -1).min(1) < 0)
    verts, keys, owner = [], [], []
    ky, kz = 2 * int(n_xyz[1]) + 2, 2 * int(n_xyz[2]) + 2
    for v in np.nonzero(active)[0]:
        for s in range(8):
            base = np.array([(s >> 2) & 1, (s >> 1) & 1, s & 1])
            val = np.array([sdf[v][tuple(base + T.CORNERS[c])] for c in range(8)], F32)
            case = int(sum((1 << c) for c in range(8) if val[c] < 0))
            for t in range(int(T.TRI_COUNT[case])):
                for e in T.TRI_TABLE[case, 3 * t: 3 * t + 3]:
                    c0, c1 = T.EDGE_CORNERS[e]
                    ax = int(T.EDGE_AXIS[e])
                    p0 = base + T.CORNERS[c0]
                    tt = F32(val[c0] / F32(val[c0] - val[c1]))
                    pos = p0.astype(F32)
                    pos[ax] = F32(pos[ax] + tt)
                    vox = (F32(0.5) * pos + (coords[v].astype(F32) - F32(0.5))).astype(F32)
                    verts.append((vox * vs).astype(F32) + mn)
                    L = 2 * coords[v] + p0
                    keys.append(((int(L[0]) * ky + int(L[1])) * kz + int(L[2])) * 3 + ax)
                owner.append(v)
    if not verts:
        return np.zeros((0, 3), F32), np.zeros((0,), np.int64), np.zeros((0,), np.int64)
    return np.asarray(verts, F32), np.asarray(keys, np.int64), np.asarray(owner, np.int64)


def weld(verts, keys):
    """Merge the vertices that lie on the same lattice edge: (unique verts [V,3], faces [T,3])."""
    uk, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    return verts[first], inv.reshape(-1, 3)
