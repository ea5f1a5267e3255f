#!/usr/bin/env python
"""bench.py -- BNV-Fusion per-frame dense hot path on B200.

A "step" is one pass of the hot path over one batch of `--frame-batch` (default 7) 640x480 synthetic
depth frames of the `lounge` workload (BASELINE.json configs[1] shape: K = (525,525,319.5,239.5), 1 cm
voxels, 5.1 m cube -> 512^3 sparse grid): back-projection -> 8-neighbour expansion + encoder MLP +
per-voxel scatter-mean -> running-average integration into the voxel map (NeuralMap.integrate's
local-fusion half, /root/reference/src/run_e2e.py:78-98), through ONE bnv_fuse_frames call (the map ends
up exactly as after one call per frame; tests/test_gpu_batch.py).  `--frame-batch 1` times the one-call-
per-frame form; at N = 1 its numbers are also reported under "single_frame_calls".

  value      frames/s, depth frames resident in HBM, CUDA-event time per step, L2 flushed between
             steps (cold-cache, conservative);  `value_warm` = same steps back to back, no flush
  e2e        frames/s through the public API with HOST buffers: pinned uint16 depth -> H2D copy ->
             fuse -> D2H read of the frame statistics, every step
  roofline   dominant kernel (the fused encode kernel) against the measured bf16 tensor peak, and
             the decode kernel's roofline under "decode" (SDF-decode Mqueries/s, the second half of
             BASELINE.json's metric)
  cpu_baseline  the numpy oracle port of the same path on the host cores (bounded sample)

`--impl reference` times the reference algorithm's CPU implementation (the oracle port; the
reference itself is Python that needs packages absent from this image) on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENC_FLOP_PER_ROW = 2 * (6 * 64 + 64 * 64 + 64 * 64 + 64 * 8)         # 18 176
DEC_FLOP_PER_QUERY = 8 * 2 * (17 * 64 + 64 * 64 + 64 * 64 + 64 * 1)  # 149 504
N_FRAMES = 16
WORKLOAD = "lounge"
WORKLOAD_DESC = ("lounge 640x480 depth, 1 cm voxels, 512^3 sparse grid, per-frame local fusion "
                 "(backproject+encode+integrate), pretrained pointnet_tcnn weights")


def traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), else None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region.  nvidia-smi needs a few hundred milliseconds to
    deliver its first sample (longer with 8 ranks starting one each), so the sampler is started early and `mark()`ed
    where the timed region begins; the report covers the samples after the mark."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.t_mark = [], None, index, 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t >= self.t_mark and len(r) >= 6]
        scope = "timed region .. end of the e2e loop"
        if not rows:
            rows, scope = [r for _, r in self.rows if len(r) >= 6], "whole run (no sample fell inside the timed region)"
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "scope": scope}


def make_frames(n):
    from bnv_fusion_b200 import synth
    spec = synth.stream_spec(WORKLOAD)
    return spec, [synth.make_frame(spec, i, seed=0) for i in range(n)]


# --------------------------------------------------------------------------------------------- #
def oracle_frame(O, grid, vm, enc, d, K, T, max_depth, rows=None):
    if rows is not None:                      # bounded sample: the first `rows` image rows
        d = d[:rows]
    depth, mask = O.load_depth_u16(d, max_depth)
    pts6 = O.backproject(depth, mask, K, T)
    feats, counts, flat, coords, navg, _ = O.encode_pointcloud(pts6, grid, enc, 8)
    O.integrate(vm, flat, feats, counts)
    return 0 if flat is None else len(flat)


def cpu_port(spec, frames, n_steps, warm=True, rows=None):
    """The oracle port of the hot path on the host cores, WHOLE 640x480 frames (numpy; its matmuls use the BLAS
    threads of the box).  Returns (frames/s, seconds per frame)."""
    from oracle import bnv_oracle as O
    p = np.load(os.path.join(ROOT, "tests", "golden", "tcnn_params.npz"))
    grid = O.Grid.from_dimensions(spec.dimensions, spec.voxel_size)
    vm = O.VoxelMap(grid)
    d, K, T = frames[0]
    if warm:
        oracle_frame(O, grid, vm, p["encoder"], d, K, T, spec.max_depth, rows=8)       # warm BLAS
    t0 = time.perf_counter()
    for i in range(n_steps):
        d, K, T = frames[i % len(frames)]
        oracle_frame(O, grid, vm, p["encoder"], d, K, T, spec.max_depth, rows=rows)
    dt = (time.perf_counter() - t0) / n_steps
    return 1.0 / dt, dt


def ref_tsdf(spec, frames, n_timed=5, voxel=0.025):
    """The CPU path BASELINE.json names, as is: the reference's third_parties/fusion.py (numba parallel=True,
    fusion.py:169-206,251-294) staged unmodified into oracle/_ref by build(); TSDFVolume(use_gpu=False).integrate at
    the reference's fixed 2.5 cm (run_e2e.py:62) over the workload's volume; 1 warm-up call (JIT), median of n_timed.
    Falls back to the numpy port (oracle/tsdf_oracle.py) when the staged file is absent."""
    from oracle import bnv_oracle as O
    from oracle import stage_ref
    mn, mx, _ = O.get_world_range(spec.dimensions, voxel)
    bnds = np.stack([mn, mx], 1).astype(np.float64)
    deps = [O.load_depth_u16(frames[i % len(frames)][0], spec.max_depth)[0].astype(np.float32) for i in range(n_timed + 1)]
    rgb = np.zeros(deps[0].shape + (3,), np.float32)
    if stage_ref.available():
        import contextlib, io
        import numba
        fusion = stage_ref.load_fusion()
        with contextlib.redirect_stdout(io.StringIO()):
            tv = fusion.TSDFVolume(bnds.copy(), voxel_size=voxel, use_gpu=False)
        kind, threads = "reference", int(numba.get_num_threads())
        dims = [int(v) for v in tv._vol_dim]
    else:
        from oracle.tsdf_oracle import TSDFOracle
        tv = TSDFOracle(bnds, voxel)
        kind, threads, n_timed = "port", 1, 1
        dims = [int(v) for v in tv.dim]
    ts = []
    for i in range(n_timed + 1):
        _, K, T = frames[i % len(frames)]
        t0 = time.perf_counter()
        tv.integrate(rgb, deps[i], K, T, 1.0)
        ts.append(time.perf_counter() - t0)
    sec = float(np.median(ts[1:])) if len(ts) > 1 else ts[0]
    return {"value": 1.0 / sec, "unit": "frames/s", "kind": kind, "threads": threads, "sec_per_frame": sec,
            "warmup_sec": ts[0], "timed_calls": len(ts) - 1,
            "sample": "third_parties/fusion.py TSDFVolume(use_gpu=False).integrate, %dx%dx%d voxels @ %g cm, 640x480 frames, "
                      "1 warm-up + median of %d" % (dims[0], dims[1], dims[2], voxel * 100, len(ts) - 1)}


def run_reference(args):
    """--impl reference: the path's CPU implementation on the box's host cores, WHOLE frames, `steps` of them: the
    numpy oracle port of the neural local fusion (the reference's own modules need tinycudann / Open3D / torch_scatter,
    absent from this image), plus the reference's own coarse-TSDF CPU code for the reference's 'local' timer scope."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec, frames = make_frames(4)
    steps = max(1, args.steps)
    rows = args.ref_rows or None          # tests only: crop the frames (the driver's runs use whole frames)
    cpu_port(spec, frames, min(args.warmup, 1), warm=True, rows=rows)
    fps, dt = cpu_port(spec, frames, steps, warm=False, rows=rows)
    cores = os.cpu_count()
    tsdf = ref_tsdf(spec, frames, n_timed=1 if rows else 5, voxel=0.1 if rows else 0.025)
    what = "whole 640x480 frames" if not rows else f"frames CROPPED to {rows} rows (--ref-rows, contract test only)"
    sample = f"{steps} {what}, numpy oracle port (oracle/bnv_oracle.py, float64 MLP), BLAS threads <= {cores}"
    print(json.dumps({
        "impl": "reference", "metric": "fusion_frames_per_sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC, "frames": 4, "mlp": "float64 numpy (oracle port)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample, "tsdf": tsdf},
        "local_scope": {"value": 1.0 / (dt + tsdf["sec_per_frame"]), "unit": "frames/s",
                        "what": "neural fusion (port) + coarse TSDF (%s) per frame, reference 'local' timer scope "
                                "(run_e2e.py:78-109)" % tsdf["kind"]},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def shard_parity_check(dist, torch, model, spec, frames, dev, rank, world, brick_log2, exchange, n_frames=8, batch=1):
    """N > 1 only: every rank fuses the first `n_frames` frames into a fresh tile-sharded map; rank 0 fuses the same
    frames into an UNSHARDED map and compares it with the union of the ranks' owned voxels: keys and fusion weights
    bit-exact, features <= 5e-5 (fp32 summation order), the 27 meshlize samples of every owned voxel (decoded on
    its owner from halo copies of foreign corners) <= 1e-4.  Returns the dict printed as "shard_parity"."""
    from bnv_fusion_b200.volume import SparseVolume
    from bnv_fusion_b200.dist import TileShardedFusion
    if batch > 1:
        n_frames = min(len(frames), 2 * batch)          # two batches through the sharded bnv_fuse_frames path
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, pool_capacity=1 << 21,
                       frame_batch=batch if batch > 1 else 0)
    sh = TileShardedFusion(vol, model, rank, world, brick_log2=brick_log2, exchange=exchange, exchange_every=5)
    devf = [torch.from_numpy(frames[i][0].view(np.int16).copy()).to(dev).view(torch.uint16) for i in range(n_frames)]
    if batch > 1:
        for b0 in range(0, n_frames, batch):
            ids = list(range(b0, min(b0 + batch, n_frames)))
            sh.fuse_depth_frames([devf[i] for i in ids], np.stack([frames[i][1] for i in ids]),
                                 np.stack([frames[i][2] for i in ids]), spec.max_depth)
    else:
        for i in range(n_frames):
            sh.fuse_depth_frame(devf[i], frames[i][1], frames[i][2], spec.max_depth)
    vol.check_status()
    vol.to_tensor()
    vol.weights += 8.0
    sdf = vol.decode_voxel_blocks(model.nerf).reshape(-1, 27)
    own = sh.owned_rows()
    n = vol._n_xyz_host
    c = vol.active_coordinates[own]
    rec = torch.cat([(c[:, 0] * (n[1] * n[2]) + c[:, 1] * n[2] + c[:, 2]).double()[:, None],
                     vol.weights[own].double(), vol.features[own].double(), sdf[own].double()], dim=1).contiguous()
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = rec.shape[0]
    dist.all_reduce(sizes)
    cap = int(sizes.max())
    pad = torch.zeros((cap, rec.shape[1]), dtype=torch.float64, device=dev)
    pad[: rec.shape[0]] = rec
    allr = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, allr, dst=0)
    sh.detach()
    out = None
    if rank == 0:
        got = torch.cat([allr[r][: int(sizes[r])] for r in range(world)], dim=0)
        got = got[torch.argsort(got[:, 0])]
        ref = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, pool_capacity=1 << 21)
        for i in range(n_frames):
            model.fuse_depth_frame(ref, devf[i], frames[i][1], frames[i][2], spec.max_depth)
        ref.check_status()
        ref.to_tensor()
        ref.weights += 8.0
        rsdf = ref.decode_voxel_blocks(model.nerf).reshape(-1, 27)
        rc = ref.active_coordinates
        rflat = rc[:, 0] * (n[1] * n[2]) + rc[:, 1] * n[2] + rc[:, 2]
        order = torch.argsort(rflat)
        keys_equal = got.shape[0] == rflat.shape[0] and bool((got[:, 0].long() == rflat[order]).all())
        out = {"frames": n_frames, "frame_batch": batch, "voxels": int(rflat.shape[0]), "owned_per_rank": [int(v) for v in sizes.tolist()],
               "keys_equal": keys_equal, "weights_equal": False, "max_dfeat": None, "max_dsdf": None}
        if keys_equal:
            out["weights_equal"] = bool((got[:, 1].float() == ref.weights[order, 0]).all())
            out["max_dfeat"] = float((got[:, 2:10].float() - ref.features[order]).abs().max())
            out["max_dsdf"] = float((got[:, 10:].float() - rsdf[order]).abs().max())
        out["ok"] = bool(keys_equal and out["weights_equal"] and out["max_dfeat"] <= 5e-5 and out["max_dsdf"] <= 1e-4)
        del ref
    del vol
    return out


def make_rays(spec, frame, n_rays, rng):
    """synthetic rays of one view in the shape NeuralMap.optimize feeds calculate_loss (run_e2e.py:119-146): n_rays
    interior pixels, their world points and 3 x 3 neighbour points"""
    import torch
    d, K, T = frame
    h, w = d.shape
    z = d.astype(np.float64) / 1000.0
    valid = (z > 0) & (z < spec.max_depth)
    v, u = np.mgrid[0:h, 0:w]
    pc = np.stack([(u - K[0, 2]) / K[0, 0] * z, (v - K[1, 2]) / K[1, 1] * z, z], -1)
    pw = pc @ np.asarray(T, np.float64)[:3, :3].T + np.asarray(T, np.float64)[:3, 3]
    inner = np.zeros_like(valid)
    inner[1:-1, 1:-1] = valid[1:-1, 1:-1]
    idx = rng.permutation(np.flatnonzero(inner))[:n_rays]
    vv, uu = idx // w, idx % w
    off = np.array([[a, b] for a in (-1, 0, 1) for b in (-1, 0, 1)])
    nv, nu = vv[:, None] + off[None, :, 0], uu[:, None] + off[None, :, 1]
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return {"uv": f32(np.stack([uu, vv], -1))[None], "gt_pts": f32(pw[vv, uu])[None], "mask": torch.ones(1, len(idx)),
            "neighbor_pts": f32(pw[nv, nu])[None], "neighbor_masks": f32(valid[nv, nu])[None],
            "T_wc": f32(T)[None], "intr_mat": f32(K)[None]}


def optim_iteration(torch, model, vol, spec, frames, dev, iters=10):
    """One iteration of NeuralMap.optimize (run_e2e.py:111-156) at the reference's configured shape: 5 000 rays of one
    view in 5 splits of 1 000 (configs/dataset/fusion_inference_dataset.yaml:8, configs/model/fusion_pointnet_model.yaml:7),
    35 samples per ray, loss + backward per split, one Adam step on volume.features.  Extra key, not the headline."""
    from bnv_fusion_b200.render import calculate_loss
    rng = np.random.default_rng(0)
    views = [{k: t.to(dev) for k, t in make_rays(spec, frames[i], 5000, rng).items()} for i in range(4)]
    keep = vol.features
    vol.features = torch.nn.Parameter(vol.features.clone())
    opt = torch.optim.Adam([vol.features], lr=0.001)
    td = min(10 * spec.voxel_size * 0.5, 0.1)

    def iteration(i):
        rays = views[i % len(views)]
        opt.zero_grad()
        n = rays["uv"].shape[1]
        for s0 in range(0, n, 1000):
            sl = slice(s0, min(s0 + 1000, n))
            part = {k: (t if k in ("T_wc", "intr_mat") else t[:, sl]) for k, t in rays.items()}
            calculate_loss(vol, part, model.nerf, truncated_units=10, truncated_dist=td, ray_max_dist=3)["depth_bce_loss"].backward()
        opt.step()
        return n

    for i in range(3):
        n = iteration(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for i in range(iters):
        iteration(3 + i)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / iters
    ms = e0.elapsed_time(e1) / iters
    vol.features = keep
    return {"ms_per_iteration": ms, "host_wall_ms_per_iteration": wall_ms, "rays": n, "samples_per_ray": 35, "splits": 5,
            "queries_per_s": n * 35 / (ms * 1e-3),
            "what": "NeuralMap.optimize iteration: 5 x (ray samples, count_optim, decode, SDF loss, decode backward) + Adam step; "
                    "the backward runs on the fp32 CUDA cores"}


def single_frame_calls(torch, model, spec, frames, devf, host, dev, flush, steps):
    """The one-call-per-frame form (bnv_fuse_frame / bnv_fuse_frame_host, what an unchanged run_e2e.py loop issues) on a
    map of its own with the single-frame table layout: cold device-resident frames/s and host-buffer frames/s."""
    from bnv_fusion_b200.volume import SparseVolume
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    stats_host = torch.zeros(4, dtype=torch.int64).pin_memory()
    n = len(frames)
    for i in range(5):
        model.fuse_depth_frame(vol, devf[i % n], frames[i % n][1], frames[i % n][2], spec.max_depth, stats=stats)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        j = (5 + i) % n
        flush.zero_()
        ev[i][0].record()
        model.fuse_depth_frame(vol, devf[j], frames[j][1], frames[j][2], spec.max_depth, stats=stats)
        ev[i][1].record()
    torch.cuda.synchronize()
    cold_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))

    def e2e(i):
        j = i % n
        model.fuse_depth_frame_host(vol, host[j], frames[j][1], frames[j][2], spec.max_depth, stats_host=stats_host,
                                    next_depth_mm_host=host[(j + 1) % n])
        torch.cuda.current_stream().synchronize()
    for i in range(3):
        e2e(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        e2e(3 + i)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / steps
    vol.check_status()
    del vol
    return {"value": 1e3 / cold_ms, "unit": "frames/s", "ms_per_frame": cold_ms, "e2e": 1e3 / e2e_ms,
            "what": "one bnv_fuse_frame call per frame (an unchanged run_e2e.py loop): cold L2, device-resident frames; "
                    "e2e = bnv_fuse_frame_host per frame with a host sync after every frame"}


# --------------------------------------------------------------------------------------------- #
def run_b200(args):
    import torch
    import torch.distributed as dist
    from bnv_fusion_b200 import _lib, config
    from bnv_fusion_b200.model import LitFusionPointNet
    from bnv_fusion_b200.volume import SparseVolume
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    if args.mlp:
        config.set_mlp_mode(args.mlp)
    lib = _lib.load()
    pk = peaks()
    sampler = ClockSampler(local)
    sampler.start()

    spec, frames = make_frames(N_FRAMES)
    p = np.load(os.path.join(ROOT, "tests", "golden", "tcnn_params.npz"))
    cfg = {"trainer": {"dense_volume": False},
           "model": {"feature_vector_size": 8, "voxel_size": spec.voxel_size, "min_pts_in_grid": 8,
                     "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
    model = LitFusionPointNet(cfg)
    # pretrained/pointnet_tcnn.ckpt tensors (shipped as a test fixture); random-init would time the same
    model.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]),
                           "nerf.model.params": torch.from_numpy(p["decoder"])})
    model.eval(); model.cuda(); model.freeze()
    B = max(1, int(args.frame_batch))               # frames per step (one bnv_fuse_frames call); 1: one call per frame
    if B > 7:
        raise SystemExit("--frame-batch: bnv_fuse_frames takes at most 7 frames per call")
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, frame_batch=B if B > 1 else 0)
    shard = None
    if world > 1:
        from bnv_fusion_b200.dist import TileShardedFusion
        shard = TileShardedFusion(vol, model, rank, world, brick_log2=args.brick_log2, exchange=args.exchange,
                                  exchange_every=args.exchange_every)
    H, W = spec.height, spec.width
    host = [torch.from_numpy(d.view(np.int16).copy()).pin_memory() for d, _, _ in frames]
    devf = [h.to(dev).view(torch.uint16) for h in host]
    Ks_all, Ts_all = np.stack([K for _, K, _ in frames]), np.stack([T for _, _, T in frames])
    stage = torch.empty((B, H, W), dtype=torch.int16, device=dev)

    def ids_of(i):                                   # the frames of step i
        return [(i * B + j) % N_FRAMES for j in range(B)]
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    stats_host = torch.zeros(4, dtype=torch.int64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    target = shard if shard is not None else model   # tile shard: fuse own rows; boundary voxels go out once per epoch
    pre = () if shard is not None else (vol,)

    def fuse(depths, ids):
        if B == 1:
            target.fuse_depth_frame(*pre, depths[0], Ks_all[ids[0]], Ts_all[ids[0]], spec.max_depth, stats=stats)
        else:
            target.fuse_depth_frames(*pre, depths, Ks_all[ids], Ts_all[ids], spec.max_depth, stats=stats)

    def step(i):
        ids = ids_of(i)
        fuse([devf[j] for j in ids], ids)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if shard is not None:
        # set-up, not warm-up: two exchange epochs, so that both double-buffered all-gather buffer pairs have been through
        # NCCL once before anything is timed (an epoch's first use of a buffer pair inside the timed region cost ~20 % of
        # an 8-GPU run: profiles/r2f_scaling_breakdown.md, run 2)
        for i in range(2):
            step(i)
            shard.exchange_now()
        shard.synchronize()
    for i in range(args.warmup):
        step(i)
    barrier()
    # ---- THE timed region: K steps, frames resident in HBM, L2 flushed between steps, one CUDA-event pair per step ----
    sampler.mark()
    launches0 = lib.bnv_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        step(args.warmup + i)
        ev[i][1].record()
    barrier()
    launches = lib.bnv_launch_count() - launches0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    # ---- the same K cold steps again with CUDA events around every kernel (roofline): the three extra event records
    # per step serialise the kernels' launch overlap and cost ~15 % of a 0.1 ms step, so they stay out of `value` ----
    lib.bnv_map_set_timing(vol._handle, 1)
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    enc_ms, fin_ms, pre_ms, rows_total, touched_total, kept_total = [], [], [], 0, 0, 0
    barrier()
    for i in range(args.steps):
        flush.zero_()
        ev2[i][0].record()
        step(args.warmup + i)
        ev2[i][1].record()
        ms3 = (C.c_float * 3)()
        _lib.check(lib.bnv_map_get_timing_stages(vol._handle, ms3), "timing")
        pre_ms.append(ms3[0]); enc_ms.append(ms3[1]); fin_ms.append(ms3[2])
        st = stats.tolist()
        rows_total += int(st[1]); touched_total += int(st[2]); kept_total += int(st[3])
    barrier()
    lib.bnv_map_set_timing(vol._handle, 0)
    staged_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev2]))
    # ---- same steps back to back (warm L2), one event pair ------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + args.steps + i)
    e1.record()
    barrier()
    warm_ms = e0.elapsed_time(e1) / args.steps
    # ---- end to end through the public API / C ABI with HOST buffers ----------------------------
    # every step: pinned uint16 depth -> H2D -> fuse -> D2H of the frame statistics -> host sync
    # (bnv_fuse_frame_host, one library call per frame; the tile shard adds its boundary exchange on the side stream)
    def fuse_host(i, stats_to):
        ids, nxt = ids_of(i), ids_of(i + 1)
        if B == 1:
            target.fuse_depth_frame_host(*pre, host[ids[0]], Ks_all[ids[0]], Ts_all[ids[0]], spec.max_depth, stats_host=stats_to,
                                         next_depth_mm_host=host[nxt[0]])    # prefetch hint
        else:
            target.fuse_depth_frames_host(*pre, [host[j] for j in ids], Ks_all[ids], Ts_all[ids], spec.max_depth,
                                          stats_host=stats_to, next_depths_mm_host=[host[j] for j in nxt])

    def step_e2e(i):
        fuse_host(i, stats_host)
        torch.cuda.current_stream().synchronize()          # the user reads the step's result

    for i in range(3):
        step_e2e(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_e2e(2 * args.steps + i)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / args.steps
    assert int(stats_host[0]) > 0
    # ---- the same host-buffer call, results read ONE FRAME BEHIND: frame i is enqueued before the host waits for frame
    # i - 1's statistics, so the launch latency after each host sync is hidden (extra key; `e2e` above is the strict form)
    stats2 = [torch.zeros(4, dtype=torch.int64).pin_memory() for _ in range(2)]
    done2 = [torch.cuda.Event(), torch.cuda.Event()]

    def run_pipelined(first, count):
        for j in range(count):
            fuse_host(first + j, stats2[j & 1])
            done2[j & 1].record()
            if j > 0:
                done2[(j - 1) & 1].synchronize()           # the user reads frame i - 1's result while frame i runs
                assert int(stats2[(j - 1) & 1][0]) > 0
        done2[(count - 1) & 1].synchronize()
        assert int(stats2[(count - 1) & 1][0]) > 0

    run_pipelined(0, 4)
    barrier()
    e0.record()
    run_pipelined(3 * args.steps, args.steps)
    e1.record()
    barrier()
    e2e_pipe_ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    vol.check_status()
    # ---- reference "local" timer scope: neural fusion + coarse TSDF prior (run_e2e.py:78-109) -----------
    from bnv_fusion_b200.tsdf import TSDFVolume
    from bnv_fusion_b200.volume import get_world_range
    mn, mx, _ = get_world_range(spec.dimensions, 0.025)               # run_e2e.py:62-71: fixed 2.5 cm
    tsdf = TSDFVolume(np.stack([mn, mx], 1), 0.025, device=dev, verbose=False)
    def step_local(i):
        ids = ids_of(i)
        for b, j in enumerate(ids):
            stage[b].copy_(host[j], non_blocking=True)
        fuse([stage[b].view(torch.uint16) for b in range(B)], ids)
        for b, j in enumerate(ids):
            tsdf.integrate(None, stage[b].view(torch.uint16), Ks_all[j], Ts_all[j], 1.0)
        stats_host.copy_(stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for i in range(3):
        step_local(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_local(3 * args.steps + i)
    e1.record()
    barrier()
    local_ms = e0.elapsed_time(e1) / args.steps
    tsdf_dims = [int(v) for v in tsdf._vol_dim]

    # ---- decode: 27 samples per active voxel, repeated to >= 10 M queries ----------------------
    vol.to_tensor()
    A = vol.active_coordinates.shape[0]
    vol.weights += 8.0              # every voxel "valid": the MLP runs for all corners regardless (rule D7)
    reps = max(1, int(np.ceil(10_000_000 / max(A * 27, 1))))
    vol.decode_voxel_blocks(model.nerf)
    torch.cuda.synchronize()
    dq = []
    for _ in range(3):
        flush.zero_()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(reps):
            vol.decode_voxel_blocks(model.nerf)
        d1.record()
        torch.cuda.synchronize()
        dq.append(d0.elapsed_time(d1))
    dec_ms = float(np.median(dq))
    n_q = A * 27 * reps
    dec_tflops = DEC_FLOP_PER_QUERY * n_q / (dec_ms * 1e-3) / 1e12
    # FLOPs the tensor pipe actually executes on the block path: 27 distinct MLP rows per exported voxel (+1 miss voxel)
    exe_tflops = (DEC_FLOP_PER_QUERY / 8) * 27 * (A + 1) * reps / (dec_ms * 1e-3) / 1e12 if config.mlp_mode_name() == "tc16" else dec_tflops
    # the same queries through the generic per-query kernel (arbitrary coordinates: SparseVolume.decode_pts)
    off = torch.tensor([[a, b, c] for a in (-.5, 0, .5) for b in (-.5, 0, .5) for c in (-.5, 0, .5)], device=dev)
    qc = (vol.active_coordinates.float()[:, None, :] + off[None]).reshape(1, A, 27, 3).contiguous()
    vol.decode_pts(qc, model.nerf, None, is_coords=True)
    torch.cuda.synchronize()
    gq = []
    for _ in range(3):
        flush.zero_()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(reps):           # >= 10 M queries, calls back to back like the block path above
            vol.decode_pts(qc, model.nerf, None, is_coords=True)
        d1.record()
        torch.cuda.synchronize()
        gq.append(d0.elapsed_time(d1))
    gen_ms = float(np.median(gq))
    gen_tflops = DEC_FLOP_PER_QUERY * A * 27 * reps / (gen_ms * 1e-3) / 1e12
    del qc

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    parity = None
    if world > 1 and not args.no_parity:
        parity = shard_parity_check(dist, torch, model, spec, frames, dev, rank, world, args.brick_log2, args.exchange, batch=B)
    # ---- BASELINE configs[3]: mesh extraction = SDF decode of the 27 samples of every active voxel + marching cubes ----
    mesh = None
    if world == 1:
        def timed(fn, n=4):
            ts, out = [], None
            for _ in range(n):
                out = None                     # the previous result goes back to the allocator's cache first
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                out = fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return float(np.median(ts)), out
        t_dec, blocks = timed(lambda: vol.decode_voxel_blocks(model.nerf))
        t_mc, (mv, mf) = timed(lambda: vol.extract_triangles(blocks, weld=False))
        t_weld, (wv, wf) = timed(lambda: vol.extract_triangles(blocks, weld=True))
        t0 = time.perf_counter()
        res = vol.meshlize(model.nerf)
        t_all = (time.perf_counter() - t0) * 1e3
        mesh = {"active_voxels": A, "queries": A * 27, "triangles": int(mf.shape[0]), "welded_vertices": int(wv.shape[0]),
                "decode_ms": t_dec, "marching_cubes_ms": t_mc, "marching_cubes_welded_ms": t_weld,
                "decode_plus_mc_Mqueries_per_s": A * 27 / ((t_dec + t_mc) * 1e-3) / 1e6,
                "meshlize_call_ms_incl_host_copy": t_all, "returned_mesh": res is not None,
                "what": "SparseVolume.meshlize on the fused map: bnv_decode_voxel_blocks + bnv_mesh_count / bnv_mesh_emit "
                        "(sparse_volume.py:697-766: 500-voxel decode batches + one skimage marching-cubes call per voxel on the CPU)"}
    optim = None
    if world == 1 and not args.no_optim:
        optim = optim_iteration(torch, model, vol, spec, frames, dev)
    n_q_job, dec_ms_job = n_q, dec_ms
    if world > 1:      # whole-job decode: every rank decodes its own voxels; halo copies do not count
        own = int(shard.owned_rows().sum())
        t = torch.tensor([own * 27 * reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        n_q_job, dec_ms_job = float(t[0]), maxr(dec_ms)
    ms = maxr(float(np.sum(step_ms))) / args.steps
    warm_ms, e2e_ms, local_ms, staged_ms = maxr(warm_ms), maxr(e2e_ms), maxr(local_ms), maxr(staged_ms)
    e2e_pipe_ms = maxr(e2e_pipe_ms)
    enc_avg = float(np.mean(enc_ms))
    rows_per_launch = rows_total / args.steps
    enc_tflops = ENC_FLOP_PER_ROW * rows_per_launch / (enc_avg * 1e-3) / 1e12
    pre_avg, fin_avg = float(np.mean(pre_ms)), float(np.mean(fin_ms))
    scatter_bytes = B * (2 * H * W + 64) + touched_total / args.steps * 44 + kept_total / args.steps * 80
    single = None
    if world == 1 and B > 1:
        single = single_frame_calls(torch, model, spec, frames, devf, host, dev, flush, args.steps)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        fps, dt = cpu_port(spec, frames, 2)
        cpu = {"value": fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "2 whole 640x480 frames through oracle/bnv_oracle.py (numpy, float64 MLP, BLAS threads)",
               "tsdf": ref_tsdf(spec, frames)}
    if rank == 0:
        out = {
            "metric": "fusion_frames_per_sec", "value": 1e3 * B / ms, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16" if config.mlp_mode_name() == "tc16" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC, "frames": N_FRAMES, "frames_per_step": B,
                       "step": ("one bnv_fuse_frames call over %d frames (map identical to one call per frame)" % B) if B > 1
                               else "one bnv_fuse_frame call",
                       "mlp": config.mlp_mode_name(), "l2": "flushed between timed steps (256 MB write)",
                       "parallelism": "1 GPU" if world == 1 else
                       f"tile shard over {world} GPUs, 3-D checkerboard of {1 << args.brick_log2}-voxel bricks, " +
                       (f"one all-gather of boundary voxels per {args.exchange_every} frames" if args.exchange == "nccl"
                        else f"peer-memory boundary routing every {args.exchange_every} frames")},
            "value_warm": 1e3 * B / warm_ms,
            "ms_per_step_with_kernel_events": staged_ms,
            "e2e": {"value": 1e3 * B / e2e_ms, "unit": "frames/s", "h2d_bytes_per_step": B * (H * W * 2 + 100),
                    "d2h_bytes_per_step": 32,
                    "what": "bnv_fuse_frame(s)_host per step: pinned uint16 depth frames -> H2D (the next step's copies are hinted "
                            "and overlap this step's kernels) -> fuse -> D2H statistics, then a host sync (the user reads "
                            "the result of every step)"},
            "e2e_results_one_frame_behind": {"value": 1e3 * B / e2e_pipe_ms, "unit": "frames/s",
                                             "what": "same call and copies per step as e2e, but step i is enqueued before the host "
                                                     "waits for step i - 1's statistics (every step's statistics are still read)"},
            "local_scope": {"value": 1e3 * B / local_ms, "unit": "frames/s", "tsdf_dims": tsdf_dims,
                            "what": "reference 'local' timer scope (run_e2e.py:250-252): neural fusion + coarse TSDF "
                                    "integration at 2.5 cm, host depth in, frame stats out"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "encode_ws_kernel (8 corner rows per point record -> encoder MLP on tcgen05 -> scatter-add)"
                                   if config.mlp_mode_name() == "tc16" else "encode_rows_simt_kernel (fp32 CUDA cores)",
                         "bound": "tensor",
                         "achieved": enc_tflops, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                         "frac": enc_tflops / pk["tf_burst"],
                         # ncu dram bytes per launch: captured for one frame per launch and for 7 frames per launch
                         "traffic": traffic(("encode_ws_kernel" if B == 1 else "encode_ws_kernel_batch7" if B == 7 else "-")
                                            if config.mlp_mode_name() == "tc16" else "encode_rows_simt_kernel"),
                         "peak_source": pk["src"],
                         "rows_per_launch": rows_per_launch, "kernel_ms": enc_avg, "prepass_ms": pre_avg, "finalize_ms": fin_avg,
                         "timing": "CUDA events around every kernel over a second pass of the same K cold steps "
                                   "(ms_per_step_with_kernel_events); each event-bracketed kernel reads ~3-5 us long"},
            # the scatter stage is "HBM-bound by contract": the prepass (depth in, claims + counts, point
            # records out) and finalize (scratch rows in, map upsert) kernels carry all of the frame's algorithmic bytes
            "roofline_hbm": {"kernel": ("frame_prepass_kernel + finalize_fused_kernel" if B == 1 else
                                        "frame_prepass_batch_kernel + finalize_batch_kernel") + " (scatter / upsert stage)", "bound": "hbm",
                             "achieved": scatter_bytes / ((pre_avg + fin_avg) * 1e-3) / 1e9,
                             "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": scatter_bytes / ((pre_avg + fin_avg) * 1e-3) / 1e9 / pk["hbm_gbs"],
                             "algorithmic_bytes_per_launch": scatter_bytes,
                             "traffic": (((traffic("frame_prepass_kernel") or 0) + (traffic("finalize_fused_kernel") or 0)) if B == 1 else
                                         ((traffic("frame_prepass_batch_kernel") or 0) + (traffic("finalize_batch_kernel") or 0)) if B == 7
                                         else 0) or None,
                             "note": "bytes = frames_per_step * (2*H*W (uint16 depth) + 64) + M_t*44 + M*80 (SURVEY 8d with the uint16 "
                                     "depth image this path reads); M_t, M from the step's statistics; time = prepass + finalize"},
            "decode": {"value": n_q_job / (dec_ms_job * 1e-3) / 1e6, "unit": "Mqueries/s", "queries": n_q_job,
                       "active_voxels_rank0": A, "ms": dec_ms_job,
                       "path": "bnv_decode_voxel_blocks (meshlize samples): G[voxel][offset] table on the tensor cores + blend",
                       "roofline": {"bound": "tensor", "achieved": exe_tflops, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                                    "frac": exe_tflops / pk["tf_sustained"], "frac_of_burst_peak": exe_tflops / pk["tf_burst"],
                                    "algorithmic_tflops": dec_tflops,
                                    "traffic": ((traffic("gtable_ws_kernel") or 0) + (traffic("blend_blocks_kernel") or 0)) or None
                                               if config.mlp_mode_name() == "tc16" else traffic("decode_simt_kernel"),
                                    "peak_source": pk["src"],
                                    "note": "achieved = FLOPs executed on the tensor pipe over the time of BOTH kernels (G table + "
                                            "blend): the block path evaluates each distinct (voxel, offset) MLP row once, 27 "
                                            "per voxel instead of the 216 of the reference; algorithmic_tflops uses SURVEY 8d's "
                                            "149 504 FLOP/query"},
                       "generic": {"value": A * 27 * reps / (gen_ms * 1e-3) / 1e6, "unit": "Mqueries/s", "ms": gen_ms,
                                   "queries": A * 27 * reps,
                                   "what": "same queries through bnv_decode_sdf (arbitrary coordinates, 8 MLP rows per query)",
                                   "roofline": {"bound": "tensor", "achieved": gen_tflops, "peak": pk["tf_sustained"],
                                                "unit": "TFLOP/s", "frac": gen_tflops / pk["tf_sustained"],
                                                "frac_of_burst_peak": gen_tflops / pk["tf_burst"],
                                                "traffic": traffic("decode_ws_kernel" if config.mlp_mode_name() == "tc16" else "decode_simt_kernel")}}},
            "cpu_baseline": cpu,
        }
        if single is not None:
            out["single_frame_calls"] = single
        if optim is not None:
            out["optim_iteration"] = optim
        if mesh is not None:
            out["mesh_extraction"] = mesh
        if parity is not None:
            out["shard_parity"] = parity
        print(json.dumps(out))
        if parity is not None and not parity["ok"]:
            print("shard_parity FAILED: the tile-sharded map differs from the single-GPU map", file=sys.stderr)
            if world > 1:
                dist.destroy_process_group()
            sys.exit(1)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- #
def run_paced(args):
    """BASELINE.json configs[4]: ARKit-shape 256x192 depth at a paced frame rate (default 60 fps), 2 cm voxels,
    10 % invalidated pixels; per-frame latency = host frame available -> frame statistics read back on the host
    (bnv_fuse_frame_host + stream sync; tile shard with its boundary exchange when launched on N > 1 GPUs).
    Opt-in (`--paced-fps 60`); prints its own JSON line (sustained fps, latency p50 / p99 / max, late frames)."""
    import torch
    import torch.distributed as dist
    from bnv_fusion_b200 import config, synth
    from bnv_fusion_b200.model import LitFusionPointNet
    from bnv_fusion_b200.volume import SparseVolume

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    if args.mlp:
        config.set_mlp_mode(args.mlp)
    spec = synth.stream_spec("arkit")
    n_src = 32
    frames = [synth.make_frame(spec, i, seed=0) for i in range(n_src)]
    p = np.load(os.path.join(ROOT, "tests", "golden", "tcnn_params.npz"))
    cfg = {"trainer": {"dense_volume": False},
           "model": {"feature_vector_size": 8, "voxel_size": spec.voxel_size, "min_pts_in_grid": 8,
                     "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
    model = LitFusionPointNet(cfg)
    model.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]),
                           "nerf.model.params": torch.from_numpy(p["decoder"])})
    model.eval(); model.cuda(); model.freeze()
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev)
    target, pre = model, (vol,)
    if world > 1:
        from bnv_fusion_b200.dist import TileShardedFusion
        target, pre = TileShardedFusion(vol, model, rank, world, brick_log2=args.brick_log2, exchange=args.exchange,
                                        exchange_every=args.exchange_every), ()
    host = [torch.from_numpy(d.view(np.int16).copy()).pin_memory() for d, _, _ in frames]
    stats_host = torch.zeros(4, dtype=torch.int64).pin_memory()

    def one(i):
        _, K, T = frames[i % n_src]
        target.fuse_depth_frame_host(*pre, host[i % n_src], K, T, spec.max_depth, stats_host=stats_host,
                                     next_depth_mm_host=host[(i + 1) % n_src])
        torch.cuda.current_stream().synchronize()

    for i in range(max(args.warmup, 10)):
        one(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = max(args.steps, 300)
    period = 1.0 / args.paced_fps
    lat = np.zeros(n)
    t0 = time.perf_counter() + 0.01
    for i in range(n):
        t_avail = t0 + i * period                   # the sensor delivers frame i at this instant
        while time.perf_counter() < t_avail:
            pass
        one(i)
        lat[i] = time.perf_counter() - t_avail
    total = time.perf_counter() - t0
    vol.check_status()
    p50, p99, mx = (float(np.percentile(lat, q)) * 1e3 for q in (50, 99, 100))
    late = int((lat > period).sum())
    if world > 1:
        t = torch.tensor([p50, p99, mx, float(late), total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        p50, p99, mx, late, total = float(t[0]), float(t[1]), float(t[2]), int(t[3]), float(t[4])
    if rank == 0:
        print(json.dumps({
            "metric": "paced_stream_frame_latency_ms", "value": p99, "unit": "ms (p99, max over ranks)", "n_gpus": world,
            "steps": n, "warmup": max(args.warmup, 10), "higher_is_better": False, "data": "synthetic",
            "dtype": "f16" if config.mlp_mode_name() == "tc16" else "f32",
            "config": {"workload": "arkit 256x192 depth, 2 cm voxels, 10 % invalidated pixels, paced live stream",
                       "paced_fps": args.paced_fps,
                       "parallelism": "1 GPU" if world == 1 else f"tile shard over {world} GPUs ({args.exchange} exchange)"},
            "latency_ms": {"p50": p50, "p99": p99, "max": mx}, "late_frames": late,
            "sustained_fps": n / total, "frame_budget_ms": period * 1e3,
            "what": "host frame available -> H2D -> fuse (+ boundary exchange enqueued) -> frame statistics on the host"}))
    if world > 1:
        dist.destroy_process_group()


def run_sustained(args):
    """BASELINE.json configs[2]: a long 640x480 stream (default 1000 frames) through the host-buffer call, unpaced: every
    frame is H2D-copied from pinned memory, fused, and its statistics read back before the next one.  64 distinct
    synthetic frames on a smooth orbit are cycled (the map keeps growing for the first cycle, then revisits).  Opt-in
    (`--sustained 1000`); prints its own JSON line: sustained frames/s over the whole run, per-100-frame rates (drift),
    clocks during the run, final map size.  Tile-sharded when launched on N > 1 GPUs."""
    import torch
    import torch.distributed as dist
    from bnv_fusion_b200 import config, synth
    from bnv_fusion_b200.model import LitFusionPointNet
    from bnv_fusion_b200.volume import SparseVolume

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    if args.mlp:
        config.set_mlp_mode(args.mlp)
    sampler = ClockSampler(local)
    sampler.start()
    spec = synth.stream_spec(WORKLOAD)
    n_src = 64
    frames = [synth.make_frame(spec, i, seed=0) for i in range(n_src)]
    p = np.load(os.path.join(ROOT, "tests", "golden", "tcnn_params.npz"))
    cfg = {"trainer": {"dense_volume": False},
           "model": {"feature_vector_size": 8, "voxel_size": spec.voxel_size, "min_pts_in_grid": 8,
                     "point_net": {"in_channels": 6}, "nerf": {"num_encoding_fn_xyz": 1}}}
    model = LitFusionPointNet(cfg)
    model.load_state_dict({"pointnet_backbone.model.params": torch.from_numpy(p["encoder"]),
                           "nerf.model.params": torch.from_numpy(p["decoder"])})
    model.eval(); model.cuda(); model.freeze()
    B = max(1, int(args.frame_batch))
    vol = SparseVolume(8, spec.voxel_size, spec.dimensions, 8, device=dev, frame_batch=B if B > 1 else 0)
    Ks_all, Ts_all = np.stack([K for _, K, _ in frames]), np.stack([T for _, _, T in frames])
    target, pre = model, (vol,)
    if world > 1:
        from bnv_fusion_b200.dist import TileShardedFusion
        target, pre = TileShardedFusion(vol, model, rank, world, brick_log2=args.brick_log2, exchange=args.exchange,
                                        exchange_every=args.exchange_every), ()
    host = [torch.from_numpy(d.view(np.int16).copy()).pin_memory() for d, _, _ in frames]
    stats_host = torch.zeros(4, dtype=torch.int64).pin_memory()

    def one(i):                                      # step i: B frames in, their statistics out, host sync
        ids, nxt = [(i * B + j) % n_src for j in range(B)], [((i + 1) * B + j) % n_src for j in range(B)]
        if B == 1:
            target.fuse_depth_frame_host(*pre, host[ids[0]], Ks_all[ids[0]], Ts_all[ids[0]], spec.max_depth, stats_host=stats_host,
                                         next_depth_mm_host=host[nxt[0]])
        else:
            target.fuse_depth_frames_host(*pre, [host[j] for j in ids], Ks_all[ids], Ts_all[ids], spec.max_depth,
                                          stats_host=stats_host, next_depths_mm_host=[host[j] for j in nxt])
        torch.cuda.current_stream().synchronize()

    for i in range(max(args.warmup, 5)):
        one(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    per_mark = max(1, 100 // B)                      # steps per rate sample (~100 frames)
    n_marks = max(1, int(args.sustained) // (per_mark * B))
    n = n_marks * per_mark * B                       # frames actually run
    sampler.mark()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(n_marks + 1)]
    marks[0].record()
    for i in range(n_marks * per_mark):
        one(i)
        if (i + 1) % per_mark == 0:
            marks[(i + 1) // per_mark].record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    vol.check_status()
    seg = [marks[j].elapsed_time(marks[j + 1]) for j in range(n_marks)]
    total_ms = float(np.sum(seg))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t[0])
    n_vox = len(vol)
    if rank == 0:
        print(json.dumps({
            "metric": "sustained_fusion_frames_per_sec", "value": n / (total_ms * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": n // B, "frames": n, "frames_per_step": B, "warmup": max(args.warmup, 5),
            "higher_is_better": True, "data": "synthetic",
            "dtype": "f16" if config.mlp_mode_name() == "tc16" else "f32",
            "config": {"workload": WORKLOAD_DESC + f"; {n_src} distinct frames cycled, host buffers (H2D + stats D2H + host sync every step of {B} frames)",
                       "parallelism": "1 GPU" if world == 1 else f"tile shard over {world} GPUs ({args.exchange} exchange every {args.exchange_every} frames)"},
            "frames_per_sec_per_100_frames": [per_mark * B / (m * 1e-3) for m in seg],
            "clocks": clocks, "map_voxels_rank0": n_vox,
            "e2e": {"value": n / (total_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": B * (spec.height * spec.width * 2 + 100),
                    "d2h_bytes_per_step": 32}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mlp", default=None, choices=[None, "fp32", "tc16"])
    ap.add_argument("--frame-batch", type=int, default=7,
                    help="frames per step = per bnv_fuse_frames call (1: one bnv_fuse_frame call per frame)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-optim", action="store_true", help="skip the optimisation-iteration extra")
    ap.add_argument("--ref-rows", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded-vs-unsharded map comparison")
    ap.add_argument("--paced-fps", type=float, default=0.0,
                    help="opt-in: BASELINE configs[4] paced ARKit-shape stream (e.g. 60), prints latency percentiles instead")
    ap.add_argument("--sustained", type=int, default=0,
                    help="opt-in: BASELINE configs[2] long stream (e.g. 1000 frames) through the host-buffer call")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"],
                    help="tile shard boundary exchange: one NCCL all-gather per epoch (default) or sender-routed stores "
                         "into the peers' inboxes over NVLink (csrc/bnv_p2p.cu)")
    ap.add_argument("--brick-log2", type=int, default=5, help="tile shard: owner bricks of 2^b voxels per side")
    ap.add_argument("--exchange-every", type=int, default=16,
                    help="tile shard: frames per boundary-exchange epoch (halo copies are only read by decode; reads flush)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if os.environ.get("BNV_WATCHDOG"):         # diagnosing hangs: dump every thread's stack and exit
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["BNV_WATCHDOG"]), exit=True)
    if args.impl == "reference":
        run_reference(args)
    elif args.paced_fps > 0:
        run_paced(args)
    elif args.sustained > 0:
        run_sustained(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
