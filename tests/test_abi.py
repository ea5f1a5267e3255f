"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/bnv_b200.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

from bnv_fusion_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "bnv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bnv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bnv_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.bnv_abi_version() == 1
    assert isinstance(lib.bnv_last_error(), bytes)
    assert lib.bnv_launch_count() >= 0


def test_bad_arguments_fail_loudly_without_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.bnv_map_create(ctypes.byref(h), None, 8, 1024, 1024, 0)
    assert rc == -1 and b"null" in lib.bnv_last_error()
    rc = lib.bnv_mlp_create(ctypes.byref(h), None, 0, 6, 8, 0)
    assert rc == -1
    # host-buffer call and the experimental exchange validate their arguments before touching CUDA
    assert lib.bnv_fuse_frame_host(None, None, 480, 640, None, None, 3.0, None, 8, 1, None, None, None) == -1
    assert lib.bnv_exchange_create(ctypes.byref(h), None, 1024) == -1
    assert lib.bnv_map_halo_enable(None, 16) == -1 and lib.bnv_map_halo_pack(None, None, 16, None) == -1
    assert lib.bnv_exchange_push(None, None) == -1 and lib.bnv_exchange_join(None, None) == -1
    assert lib.bnv_exchange_destroy(None) == 0
    # frame batches: null map / null frame list / out-of-range batch size
    assert lib.bnv_fuse_frames(None, None, 2, 480, 640, None, None, 3.0, None, 8, 1, None, None, None) == -1
    assert lib.bnv_fuse_frames_host(None, None, 2, 480, 640, None, None, 3.0, None, 8, 1, None, None, 0, None) == -1
    assert lib.bnv_map_set_frame_batch(None, 7) == -1 and b"frames per batch" in lib.bnv_last_error()


def test_batch_camera_broadcast():
    """LitFusionPointNet._batch_cameras: one shared K [3,3] or one K per frame; T_wc per frame (host logic of
    fuse_depth_frames)"""
    import numpy as np
    from bnv_fusion_b200.model import LitFusionPointNet
    K = np.arange(9, dtype=np.float64).reshape(3, 3)
    Ts = np.stack([np.eye(4) * (i + 1) for i in range(3)])
    k, t = LitFusionPointNet._batch_cameras(3, K, Ts)
    assert k.shape == (3, 9) and k.dtype == np.float32 and k.flags.c_contiguous and np.array_equal(k[2], np.arange(9))
    assert t.shape == (3, 16) and t.dtype == np.float32 and t[1, 0] == 2.0
    k2, _ = LitFusionPointNet._batch_cameras(3, np.stack([K, K + 1, K + 2]), Ts)
    assert np.array_equal(k2[1], np.arange(9) + 1)


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU only: the oracle port on the host cores) prints ONE JSON line with
    the driver's keys; runs here without a GPU."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-rows", "40"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
