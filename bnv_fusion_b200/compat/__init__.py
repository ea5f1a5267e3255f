"""Drop-in wiring: make the reference's own import paths resolve to the B200 modules so that
/root/reference/src/run_e2e.py runs unchanged (`run_e2e.py:12-21` imports).

    import bnv_fusion_b200.compat as compat
    compat.install(reference_root)          # before run_e2e.py is imported / executed

replaces, in sys.modules,

    src.models.fusion.local_point_fusion   -> LitFusionPointNet        (bnv_fusion_b200.model)
    src.models.sparse_volume               -> SparseVolume             (bnv_fusion_b200.volume)
    third_parties.fusion                   -> TSDFVolume               (bnv_fusion_b200.tsdf)
    src.utils.voxel_utils                  -> get_world_range / flatten / unflatten (+ the reference's
                                              own remaining helpers when the reference tree is importable)
    src.utils.render_utils.calculate_loss  -> calculate_loss           (bnv_fusion_b200.render)

and registers light parent packages (src, src.models, src.models.fusion, third_parties) whose __path__
still points into the reference tree, so everything OFF the hot path (datasets, hydra/rich helpers,
o3d mesh post-processing, render_utils) keeps coming from the reference checkout.  `python -m
bnv_fusion_b200.compat.run_e2e <hydra overrides>` does the install and then executes the reference's
run_e2e.py (needs the reference's Python dependencies: hydra, omegaconf, lightning's seed_everything,
trimesh, scikit-image for marching cubes).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np


def _voxel_utils_module():
    import torch
    m = types.ModuleType("src.utils.voxel_utils")

    def flatten(voxels, volume_resolution):
        """src/utils/voxel_utils.py:62-65"""
        return voxels[..., 0] * volume_resolution[1] * volume_resolution[2] + voxels[..., 1] * volume_resolution[2] + voxels[..., 2]

    def unflatten(flat_id, volume_resolution):
        """src/utils/voxel_utils.py:68-80"""
        nyz = volume_resolution[1] * volume_resolution[2]
        if isinstance(flat_id, torch.Tensor):
            x = torch.div(flat_id, nyz, rounding_mode="floor")
            rest = flat_id % nyz
            y = torch.div(rest, volume_resolution[2], rounding_mode="floor")
            z = flat_id - x * nyz - y * volume_resolution[2]
            return torch.stack([x, y, z], axis=-1)
        x = flat_id // nyz
        y = (flat_id % nyz) // volume_resolution[2]
        return np.stack([x, y, flat_id - x * nyz - y * volume_resolution[2]], axis=-1)

    from ..volume import get_world_range
    m.flatten, m.unflatten, m.get_world_range = flatten, unflatten, get_world_range
    return m


def install(reference_root: str | None = None):
    """Register the B200 modules under the reference's import names (idempotent)."""
    from .. import model, tsdf, volume
    root = reference_root or os.environ.get("BNV_REFERENCE_ROOT")

    def pkg(name, rel):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(root, rel)] if root else []
            sys.modules[name] = m
        return m

    src = pkg("src", "src")
    models = pkg("src.models", "src/models")
    fusion_pkg = pkg("src.models.fusion", "src/models/fusion")
    utils = pkg("src.utils", "src/utils")
    tp = pkg("third_parties", "third_parties")
    src.models, src.utils, models.fusion = models, utils, fusion_pkg

    lpf = types.ModuleType("src.models.fusion.local_point_fusion")
    lpf.LitFusionPointNet = model.LitFusionPointNet
    sys.modules[lpf.__name__] = lpf
    fusion_pkg.local_point_fusion = lpf

    sv = types.ModuleType("src.models.sparse_volume")
    sv.SparseVolume = volume.SparseVolume
    sys.modules[sv.__name__] = sv
    models.sparse_volume = sv

    tf = types.ModuleType("third_parties.fusion")
    tf.TSDFVolume = tsdf.TSDFVolume
    tf.FUSION_GPU_MODE = 1
    sys.modules[tf.__name__] = tf
    tp.fusion = tf

    # calculate_loss: with a reference root the reference's render_utils module is loaded (its other helpers stay) and
    # only calculate_loss is replaced; without one a bare module with the B200 function is registered
    from .. import render
    ru = None
    if root and os.path.exists(os.path.join(root, "src", "utils", "render_utils.py")):
        import importlib
        try:
            ru = importlib.import_module("src.utils.render_utils")
        except ImportError:
            ru = None
    if ru is None:
        ru = types.ModuleType("src.utils.render_utils")
        sys.modules[ru.__name__] = ru
    ru.calculate_loss = render.calculate_loss
    utils.render_utils = ru

    vu = _voxel_utils_module()
    if root:
        # the reference's remaining helpers (get_frustrum_range, depth_to_tsdf, ...): executed from the reference file
        # under a private name; the three hot-path functions above keep their B200 definitions
        ref_file = os.path.join(root, "src", "utils", "voxel_utils.py")
        if os.path.exists(ref_file):
            import importlib.util
            spec = importlib.util.spec_from_file_location("_bnv_ref_voxel_utils", ref_file)
            ref_vu = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(ref_vu)
                for k, v in vars(ref_vu).items():
                    if not k.startswith("__") and not hasattr(vu, k):
                        setattr(vu, k, v)
            except ImportError:
                pass                       # a dependency of the reference file is missing: hot-path functions only
    sys.modules[vu.__name__] = vu
    utils.voxel_utils = vu
    if root and root not in sys.path:
        sys.path.insert(0, root)
    return {"src.models.fusion.local_point_fusion": lpf, "src.models.sparse_volume": sv, "third_parties.fusion": tf,
            "src.utils.voxel_utils": vu, "src.utils.render_utils": ru}
