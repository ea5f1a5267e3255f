"""bnv_fusion_b200 -- BNV-Fusion's per-frame dense hot path (back-projection, voxel scatter of
PointNet features, 8-corner gather + tiny-MLP SDF decode) as hand-written CUDA for B200 (sm_100a)
behind the reference's own Python API.  See DESIGN.md / INTEGRATION.md."""
from . import _lib, config, synth                      # noqa: F401
from .config import set_mlp_mode, mlp_mode_name       # noqa: F401


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import bnv_fusion_b200` stays cheap
    if name in ("SparseVolume", "get_world_range"):
        from . import volume
        return getattr(volume, name)
    if name in ("LitFusionPointNet", "tcnnNeRFModel", "tcnnPointNetEncoder", "backproject"):
        from . import model
        return getattr(model, name)
    raise AttributeError(name)
