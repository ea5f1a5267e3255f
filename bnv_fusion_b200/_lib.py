"""ctypes binding of libbnv_b200.so (the C ABI declared in include/bnv_b200.h).

The product path has no CPU fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  `build()` compiles the library in-tree with nvcc for sm_100a (works on a
machine without a GPU; the built .so travels to the GPU box with the repository snapshot).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbnv_b200.so")
CSRC = os.path.join(_HERE, "csrc")

MLP_FP32 = 0
MLP_TC16 = 1

_lib = None


class Geom(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("voxel_size", C.c_double),
                ("n_xyz", C.c_int32 * 3)]


# name -> (restype, argtypes); mirrors include/bnv_b200.h one to one
_P = C.c_void_p
_I64 = C.c_int64
SIGNATURES = {
    "bnv_abi_version": (C.c_int, []),
    "bnv_last_error": (C.c_char_p, []),
    "bnv_launch_count": (_I64, []),
    "bnv_mlp_create": (C.c_int, [C.POINTER(_P), _P, _I64, C.c_int, C.c_int, C.c_int]),
    "bnv_mlp_destroy": (C.c_int, [_P]),
    "bnv_mlp_forward": (C.c_int, [_P, _P, _I64, _P, C.c_int, _P]),
    "bnv_map_create": (C.c_int, [C.POINTER(_P), C.POINTER(Geom), C.c_int, _I64, _I64, C.c_int]),
    "bnv_map_destroy": (C.c_int, [_P]),
    "bnv_map_reset": (C.c_int, [_P, _P]),
    "bnv_map_size": (C.c_int, [_P, C.POINTER(_I64), _P]),
    "bnv_map_status": (C.c_int, [_P, _P]),
    "bnv_map_set_shard": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "bnv_tsdf_create": (C.c_int, [C.POINTER(_P), _P, C.c_double, C.c_int]),
    "bnv_tsdf_destroy": (C.c_int, [_P]),
    "bnv_tsdf_dims": (C.c_int, [_P, _P]),
    "bnv_tsdf_integrate": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_double, _P]),
    "bnv_tsdf_volume": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "bnv_tsdf_copy": (C.c_int, [_P, C.c_int, _P, _P]),
    "bnv_tsdf_prior": (C.c_int, [_P, C.c_double, C.c_double, _P, _P]),
    "bnv_map_set_timing": (C.c_int, [_P, C.c_int]),
    "bnv_map_get_timing": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "bnv_map_get_timing_stages": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "bnv_map_halo_enable": (C.c_int, [_P, _I64]),
    "bnv_map_halo_pack": (C.c_int, [_P, _P, _I64, _P]),
    "bnv_map_insert_halo": (C.c_int, [_P, _P, C.c_int, _I64, _P]),
    "bnv_exchange_create": (C.c_int, [C.POINTER(_P), _P, _I64]),
    "bnv_exchange_handle": (C.c_int, [_P, _P]),
    "bnv_exchange_connect": (C.c_int, [_P, _P]),
    "bnv_exchange_push": (C.c_int, [_P, _P]),
    "bnv_exchange_join": (C.c_int, [_P, _P]),
    "bnv_exchange_destroy": (C.c_int, [_P]),
    "bnv_map_query": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P, _P]),
    "bnv_map_insert": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P]),
    "bnv_map_export": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P]),
    "bnv_map_count_optim": (C.c_int, [_P, _P, _I64, _P, _I64, _P]),
    "bnv_map_count_optim_queries": (C.c_int, [_P, _P, _I64, C.c_int, _P, _I64, _P]),
    "bnv_ray_samples": (C.c_int, [_P, _P, _I64, _P, _P, _P, C.c_int, _P, C.c_int, C.c_double, _P, _P]),
    "bnv_ray_sdf_loss": (C.c_int, [_P, _P, _I64, C.c_int, _P, _P, _P, _P, C.c_int, _P, _P, C.c_double, _P, _P, _P]),
    "bnv_backproject": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, C.c_double, _P, _P, _P]),
    "bnv_encode_points": (C.c_int, [_P, _P, _I64, _P, C.c_int, C.c_int, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "bnv_integrate": (C.c_int, [_P, _P, _P, _P, _I64, _P]),
    "bnv_fuse_frame_host": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, C.c_double, _P, C.c_int, C.c_int, _P, _P, _P]),
    "bnv_fuse_frame": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, C.c_double, _P, C.c_int, C.c_int, _P, _P, _P]),
    "bnv_map_set_frame_batch": (C.c_int, [_P, C.c_int]),
    "bnv_fuse_frames": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_double, _P, C.c_int, C.c_int, _P, _P, _P]),
    "bnv_fuse_frames_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_double, _P, C.c_int, C.c_int, _P, _P,
                                       C.c_int, _P]),
    "bnv_fuse_points": (C.c_int, [_P, _P, _I64, _P, C.c_int, C.c_int, _P, _P, _P]),
    "bnv_decode_sdf": (C.c_int, [_P, _P, _I64, C.c_int, _P, _P, _I64, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "bnv_decode_sdf_backward": (C.c_int, [_P, _P, _I64, C.c_int, _P, _P, _I64, _P, C.c_int, _P, _P, _P]),
    "bnv_mesh_count": (C.c_int, [_P, _I64, _P, _P]),
    "bnv_mesh_emit": (C.c_int, [_P, _P, _I64, _P, C.c_double, _P, _P, _I64, _P, _P, _P, _P]),
    "bnv_decode_voxel_blocks": (C.c_int, [_P, _I64, _I64, _P, _P, _I64, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
}


def build(verbose: bool = False) -> str:
    """Compile libbnv_b200.so in-tree (nvcc, -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    r = subprocess.run(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libbnv_b200.so failed (see output above)")
    return LIB_PATH


def load():
    """Load the library (never builds implicitly on a GPU box: the .so must have been shipped)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the BNV-Fusion hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.bnv_abi_version() != 1:
        raise RuntimeError("libbnv_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().bnv_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libbnv_b200 {what} failed ({rc}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
