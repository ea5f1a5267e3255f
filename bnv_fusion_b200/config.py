"""Process-wide knobs of the B200 hot path."""
import os

from . import _lib

DEFAULT_POOL_CAPACITY = 1 << 24     # voxel value-pool rows per map (0.67 GB of 180 GB HBM3e)
DEFAULT_MAX_POINTS = 640 * 480      # most points one frame may carry (sizes per-frame scratch)

_MODE = {"fp32": _lib.MLP_FP32, "tc16": _lib.MLP_TC16}
_mode = _MODE[os.environ.get("BNV_MLP_MODE", "tc16").lower()]


def set_mlp_mode(name: str):
    """'fp32' = CUDA-core exact-parity arithmetic; 'tc16' = tcgen05 tensor cores (fp16 operands,
    fp32 accumulate in TMEM)."""
    global _mode
    _mode = _MODE[name.lower()]


def mlp_mode() -> int:
    return _mode


def mlp_mode_name() -> str:
    return "tc16" if _mode == _lib.MLP_TC16 else "fp32"
