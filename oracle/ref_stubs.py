"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stub-import harness that executes the UNMODIFIED reference sources
(/root/reference, read-only) on CPU in THIS container, with only the missing
third-party packages faked (open3d.core.HashMap, torch_scatter, tinycudann,
pytorch_lightning, ...).  It exists to (1) pin `oracle/bnv_oracle.py` against the
reference's own control flow / indexing / thresholds and (2) mint the golden
vectors committed under `tests/golden/` (see `tests/golden/make_golden.py`).

It cannot travel to the GPU box (no /root/reference there); nothing in
`tests -m gpu`, `bench.py` or `__graft_entry__.smoke()` imports it.

Fakes and the reference API surface they cover:
  * open3d.core : Device, int64, Dtype.Float32, Tensor (dlpack, index, to, cpu,
    numpy, ==), HashMap(insert/find/active_buf_indices/key_tensor/value_tensor)
    -- used by src/models/sparse_volume.py:525-695,835-892
  * torch_scatter.scatter_mean -- src/models/fusion/local_point_fusion.py:125
  * tinycudann.NetworkWithInputEncoding -- src/utils/pointnet_utils.py:274-279,
    src/models/fusion/modules.py:171-176.  The fake restates tiny-cuda-nn's
    FullyFusedMLP semantics (NVlabs/tiny-cuda-nn, version un-pinned by the
    reference): params = [W0(64 x in_pad) | W1(64x64) | W2(64x64) | W3(16x64)],
    each row-major [out,in], no bias, ReLU hidden, identity encoding padding the
    input with ONES to a multiple of 16.  (SURVEY.md Appendix B validates layout
    and padding against pretrained/pointnet_tcnn.ckpt.)
"""
from __future__ import annotations

import os
import pickle
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("BNV_REFERENCE_ROOT", "/root/reference")


# --------------------------------------------------------------------------- #
# fake tinycudann
# --------------------------------------------------------------------------- #
class _FakeTcnnNetwork(torch.nn.Module):
    """Restatement of tcnn.NetworkWithInputEncoding(Identity -> FullyFusedMLP).

    mode: 'fp32' (exact fp32 math, the oracle contract) or 'fp16' (weights and
    per-layer activations rounded to fp16, fp32 accumulate, fp16 output -- an
    emulation of the reference's tensor-core precision).
    """

    mode = "fp32"

    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config):
        super().__init__()
        assert encoding_config["otype"] == "Identity"
        assert network_config["otype"] == "FullyFusedMLP"
        self.n_in = int(n_input_dims)
        self.n_out = int(n_output_dims)
        self.width = int(network_config["n_neurons"])
        self.n_hidden = int(network_config["n_hidden_layers"])
        self.in_pad = ((self.n_in + 15) // 16) * 16
        self.out_pad = ((self.n_out + 15) // 16) * 16
        n_params = (self.width * self.in_pad + (self.n_hidden - 1) * self.width * self.width
                    + self.out_pad * self.width)
        self.params = torch.nn.Parameter(torch.zeros(n_params, dtype=torch.float32))

    def blocks(self):
        p = self.params.detach()
        o = 0
        out = []
        w = self.width
        out.append(p[o:o + w * self.in_pad].reshape(w, self.in_pad)); o += w * self.in_pad
        for _ in range(self.n_hidden - 1):
            out.append(p[o:o + w * w].reshape(w, w)); o += w * w
        out.append(p[o:o + self.out_pad * w].reshape(self.out_pad, w)); o += self.out_pad * w
        assert o == p.numel()
        return out

    def forward(self, x):
        assert x.shape[-1] == self.n_in
        n = x.shape[0]
        X = torch.ones(n, self.in_pad, dtype=torch.float32)
        X[:, :self.n_in] = x.float()
        Ws = self.blocks()
        if self.mode == "fp16":
            X = X.half().float()
            Ws = [w.half().float() for w in Ws]
        h = X
        for i, W in enumerate(Ws):
            h = h @ W.t()
            if i < len(Ws) - 1:
                h = torch.relu(h)
            if self.mode == "fp16":
                h = h.half().float()
        y = h[:, :self.n_out]
        return y.half() if self.mode == "fp16" else y


# --------------------------------------------------------------------------- #
# fake open3d.core
# --------------------------------------------------------------------------- #
class _O3Tensor:
    """Thin wrapper over a torch tensor with the o3c.Tensor surface SparseVolume uses."""

    def __init__(self, t):
        self.t = t

    @staticmethod
    def from_dlpack(cap):
        if isinstance(cap, torch.Tensor):
            return _O3Tensor(cap)
        return _O3Tensor(torch.utils.dlpack.from_dlpack(cap))

    def to_dlpack(self):
        return torch.utils.dlpack.to_dlpack(self.t.contiguous())

    def to(self, dtype):
        if dtype is _INT64:
            return _O3Tensor(self.t.to(torch.int64))
        if dtype is _Dtype.Float32:
            return _O3Tensor(self.t.to(torch.float32))
        return self

    def cpu(self):
        return self

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __len__(self):
        return self.t.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, _O3Tensor):
            idx = idx.t
        return _O3Tensor(self.t[idx])

    def __setitem__(self, idx, val):
        if isinstance(idx, _O3Tensor):
            idx = idx.t
        if isinstance(val, _O3Tensor):
            val = val.t
        self.t[idx] = val

    def __eq__(self, other):
        return _O3Tensor(self.t == other)


class _Dtype:
    Float32 = object()


_INT64 = object()


class _FakeHashMap:
    """dict + growable torch buffers; insertion order defines buf indices."""

    def __init__(self, capacity, key_dtype=None, key_element_shape=None, value_dtype=None,
                 value_element_shape=None, value_dtypes=None, value_element_shapes=None,
                 device=None):
        if value_dtypes is None:
            value_dtypes = (value_dtype,)
            value_element_shapes = (value_element_shape,)
        self._kshape = tuple(key_element_shape)
        self._vshapes = [tuple(s) for s in value_element_shapes]
        self._vdt = [torch.float32 if d is _Dtype.Float32 else torch.int64 for d in value_dtypes]
        self._cap = max(int(capacity), 16)
        self._keys = torch.zeros((self._cap,) + self._kshape, dtype=torch.int64)
        self._vals = [torch.zeros((self._cap,) + s, dtype=d) for s, d in zip(self._vshapes, self._vdt)]
        self._map = {}
        self._n = 0

    def _grow(self, need):
        if need <= self._cap:
            return
        cap = max(need, self._cap * 2)
        k = torch.zeros((cap,) + self._kshape, dtype=torch.int64)
        k[:self._cap] = self._keys
        self._keys = k
        vs = []
        for v in self._vals:
            nv = torch.zeros((cap,) + tuple(v.shape[1:]), dtype=v.dtype)
            nv[:self._cap] = v
            vs.append(nv)
        self._vals = vs
        self._cap = cap

    def insert(self, keys, values):
        keys = keys.t if isinstance(keys, _O3Tensor) else keys
        if not isinstance(values, (tuple, list)):
            values = (values,)
        values = [v.t if isinstance(v, _O3Tensor) else v for v in values]
        n = keys.shape[0]
        self._grow(self._n + n)
        buf = torch.zeros(n, dtype=torch.int32)
        mask = torch.zeros(n, dtype=torch.bool)
        kl = keys.tolist()
        for i, k in enumerate(kl):
            k = tuple(k)
            if k in self._map:
                buf[i] = self._map[k]
            else:
                j = self._n
                self._n += 1
                self._map[k] = j
                self._keys[j] = keys[i]
                for v, src in zip(self._vals, values):
                    v[j] = src[i].reshape(v[j].shape).to(v.dtype)
                buf[i] = j
                mask[i] = True
        return _O3Tensor(buf), _O3Tensor(mask)

    def find(self, keys):
        keys = keys.t if isinstance(keys, _O3Tensor) else keys
        n = keys.shape[0]
        buf = torch.zeros(n, dtype=torch.int32)
        mask = torch.zeros(n, dtype=torch.bool)
        for i, k in enumerate(keys.tolist()):
            j = self._map.get(tuple(k))
            if j is not None:
                buf[i] = j
                mask[i] = True
        return _O3Tensor(buf), _O3Tensor(mask)

    def active_buf_indices(self):
        return _O3Tensor(torch.arange(self._n, dtype=torch.int32))

    def key_tensor(self):
        return _O3Tensor(self._keys)

    def value_tensor(self, i=0):
        return _O3Tensor(self._vals[i])


def _scatter_mean(src, index, dim=-1, **kw):
    """torch_scatter.scatter_mean along the last dim (fp32 sums, exact counts)."""
    assert dim == -1
    index = index.expand_as(src)
    n = int(index.max()) + 1
    out = torch.zeros(src.shape[:-1] + (n,), dtype=src.dtype)
    out.scatter_add_(-1, index, src)
    cnt = torch.zeros(src.shape[:-1] + (n,), dtype=src.dtype)
    cnt.scatter_add_(-1, index, torch.ones_like(src))
    return out / cnt.clamp(min=1)


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return _AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


_INSTALLED = False


def install_stubs():
    """Register the fake modules and put the reference on sys.path (idempotent)."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}; this harness only runs "
                           "in the build container")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    core = mod("open3d.core", Device=lambda s: s, int64=_INT64, Dtype=_Dtype, Tensor=_O3Tensor,
               HashMap=_FakeHashMap)
    mod("open3d", core=core)

    class _Trimesh:
        def __init__(self, vertices=None, faces=None, **kw):
            self.vertices, self.faces = vertices, faces

        def export(self, path):
            pass

    mod("trimesh", Trimesh=_Trimesh)

    def _no_mc(*a, **k):
        raise RuntimeError("marching cubes is out of the oracle's scope")

    measure = mod("skimage.measure", marching_cubes=_no_mc, marching_cubes_lewiner=_no_mc)
    transform = mod("skimage.transform")
    mod("skimage", measure=measure, transform=transform)
    mod("torch_scatter", scatter_mean=_scatter_mean)

    class _LM(torch.nn.Module):
        @property
        def device(self):
            return torch.device("cpu")

        def freeze(self):
            for p in self.parameters():
                p.requires_grad = False
            self.eval()

        def log(self, *a, **k):
            pass

    util = mod("pytorch_lightning.utilities", rank_zero_only=lambda f: f)
    cbs = mod("pytorch_lightning.callbacks")
    mod("pytorch_lightning", LightningModule=_LM, seed_everything=lambda s: None, utilities=util,
        callbacks=cbs)
    import json
    mod("commentjson", load=json.load, loads=json.loads)
    mod("tinycudann", NetworkWithInputEncoding=_FakeTcnnNetwork)
    kd = mod("kornia.geometry.depth", depth_to_normals=None, depth_to_3d=None)
    kg = mod("kornia.geometry", depth=kd)
    mod("kornia", geometry=kg)
    mod("imageio")
    mod("omegaconf", DictConfig=dict, OmegaConf=types.SimpleNamespace(to_yaml=lambda *a, **k: ""))
    if "rich" not in sys.modules:
        try:
            import rich  # noqa: F401
            import rich.syntax  # noqa: F401
            import rich.tree  # noqa: F401
        except Exception:
            mod("rich.syntax")
            mod("rich.tree")
            mod("rich")
    sys.path.insert(0, REF_ROOT)
    _INSTALLED = True


def load_ckpt_state_dict(name="pointnet_tcnn.ckpt"):
    """Lightning checkpoints reference pytorch_lightning classes: unpickle tolerantly."""

    class _U(pickle.Unpickler):
        def find_class(self, module, cls):
            try:
                return super().find_class(module, cls)
            except Exception:
                return type(cls, (), {})

    pm = types.ModuleType("bnv_tolerant_pickle")
    pm.Unpickler = _U
    pm.load = pickle.load
    path = os.path.join(REF_ROOT, "pretrained", name)
    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=pm)["state_dict"]


def write_tcnn_json(tmpdir):
    """The reference opens its tcnn config with json.load although the file holds // comments
    (src/models/tcnn_config.json); write a comment-free copy with the same values."""
    import json
    cfg = {
        "encoding": {"otype": "Identity", "scale": 1.0, "offset": 0.0},
        "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                    "n_neurons": 64, "n_hidden_layers": 3},
    }
    p = os.path.join(tmpdir, "tcnn_config.json")
    with open(p, "w") as f:
        json.dump(cfg, f)
    return p


def make_cfg(tcnn_json, voxel_size=0.01, min_pts=8):
    return to_attr({
        "device_type": "cpu",
        "trainer": {"dense_volume": False},
        "model": {
            "feature_vector_size": 8, "tiny_cuda": True, "tcnn_config": tcnn_json,
            "point_net": {"in_channels": 6},
            "nerf": {"hidden_size": 256, "num_layers": 4, "num_encoding_fn_xyz": 1,
                     "num_encoding_fn_dir": 6, "include_input_xyz": True, "include_input_dir": True,
                     "interpolate_decode": True, "global_coords": False, "xyz_agnostic": False},
            "voxel_size": voxel_size, "bound_min": [-1, -1, -1], "bound_max": [1, 1, 1],
            "training_global": False, "loss": {"bce_loss": 1.0, "reg_loss": 0.001},
            "min_pts_in_grid": min_pts,
        },
    })


def build_reference(workdir, voxel_size=0.01, min_pts=8, mlp_mode="fp32"):
    """Instantiate the reference's LitFusionPointNet (tcnn variant over the fake tcnn) with the
    shipped pretrained weights.  Returns (model, SparseVolume class)."""
    install_stubs()
    _FakeTcnnNetwork.mode = mlp_mode
    cwd = os.getcwd()
    os.chdir(workdir)  # the ctor makes ./plots (local_point_fusion.py:47-49)
    try:
        from src.models.fusion.local_point_fusion import LitFusionPointNet
        from src.models.sparse_volume import SparseVolume
        cfg = make_cfg(write_tcnn_json(workdir), voxel_size, min_pts)
        model = LitFusionPointNet(cfg)
        sd = load_ckpt_state_dict("pointnet_tcnn.ckpt")
        missing = model.load_state_dict(sd)
        assert not missing.missing_keys and not missing.unexpected_keys, missing
        model.eval()
        model.freeze()
    finally:
        os.chdir(cwd)
    return model, SparseVolume


class cuda_div_semantics:
    """Context manager: make `tensor / python_scalar` behave like PyTorch-CUDA's true-div fast
    path (a * (1.0f / (float)b), ATen BinaryDivTrueKernel.cu) while the reference code runs on
    CPU.  Used to mint the 'cuda-form' goldens (SURVEY.md rule A2)."""

    def __enter__(self):
        self._orig = torch.Tensor.__truediv__

        orig = self._orig

        def patched(a, b):
            if isinstance(b, (int, float)) and a.dtype == torch.float32:
                inv = np.float32(1.0) / np.float32(b)
                return a * float(inv)
            return orig(a, b)

        torch.Tensor.__truediv__ = patched
        return self

    def __exit__(self, *exc):
        torch.Tensor.__truediv__ = self._orig
        return False
