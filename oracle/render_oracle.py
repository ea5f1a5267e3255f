"""TEST INFRASTRUCTURE ONLY -- numpy (float32) restatement of the ray-sampling and SDF-loss arithmetic of the reference's
global optimisation step (/root/reference/src/utils/render_utils.py: lift :405-423, get_camera_params :426-458,
stratified_sampling :76-95, hierarchical_sampling :190-233, render_with_rays :461-507, compute_sdf_loss :510-557,
calculate_loss :559-594).  Pinned by tests/test_render_oracle.py against tests/golden/golden_loss.npz, minted from the
reference's own code + torch autograd (tests/golden/make_golden_loss.py); the stratified random draws are inputs.
"""
import numpy as np

F32 = np.float32


def linspace01(steps):
    """torch.linspace(0, 1, steps) in float32: start + i * step below the midpoint, end - (steps - 1 - i) * step above"""
    step = F32(1.0) / F32(steps - 1)
    i = np.arange(steps)
    lo = (F32(0.0) + step * i.astype(F32)).astype(F32)
    hi = (F32(1.0) - step * (steps - 1 - i).astype(F32)).astype(F32)
    return np.where(i < steps // 2, lo, hi).astype(F32)


def camera_params(uv, T_wc, K):
    """get_camera_params: ray directions [n,3] (unit) and camera centre [3] from pixel coordinates uv [n,2]"""
    uv, T, K = np.asarray(uv, F32), np.asarray(T_wc, F32), np.asarray(K, F32)
    fx, fy, cx, cy, sk = K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1]
    x, y = uv[:, 0], uv[:, 1]
    z = (x * F32(0.0) + F32(1.0)).astype(F32)
    xl = ((((x - cx) + cy * sk / fy) - sk * y / fy).astype(F32) / fx * z).astype(F32)
    yl = ((y - cy).astype(F32) / fy * z).astype(F32)
    cam = np.stack([xl, yl, z, np.ones_like(z)], -1).astype(F32)                       # [n,4]
    world = np.zeros((len(x), 3), F32)
    for r in range(3):                                                                  # bmm, K = 4, sequential order
        acc = (T[r, 0] * cam[:, 0]).astype(F32)
        for k in range(1, 4):
            acc = (acc + (T[r, k] * cam[:, k]).astype(F32)).astype(F32)
        world[:, r] = acc
    cam_loc = T[:3, 3].copy()
    d = (world - cam_loc[None]).astype(F32)
    nrm = np.sqrt((d * d).sum(-1, dtype=F32)).astype(F32)
    return (d / np.maximum(nrm, F32(1e-12))[:, None]).astype(F32), cam_loc


def _depth(p, cam_loc):
    d = (p - cam_loc).astype(F32)
    return np.sqrt(((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(F32) + d[..., 2] * d[..., 2]).astype(F32)).astype(F32)


def _stratified(n_samples, distances, t_rand):
    base = (linspace01(n_samples)[None, :] * distances[:, None]).astype(F32)            # [n,S]
    mids = (F32(0.5) * (base[:, 1:] + base[:, :-1]).astype(F32)).astype(F32)
    upper = np.concatenate([mids, base[:, -1:]], -1)
    lower = np.concatenate([base[:, :1], mids], -1)
    return (lower + ((upper - lower).astype(F32) * t_rand).astype(F32)).astype(F32)


def sample_rays(uv, gt_pts, T_wc, K, t_fine, t_coarse, truncated_dist):
    """render_with_rays up to the decode: points on rays [n,S,3] (S = fine + coarse, sorted by distance like the
    reference), their distances [n,S], the rays' gt depth [n], directions and camera centre"""
    dirs, cam_loc = camera_params(uv, T_wc, K)
    gt = np.asarray(gt_pts, F32)
    gt_depth = _depth(gt, cam_loc[None])
    off = F32(truncated_dist)
    half = (np.zeros_like(gt_depth) + off).astype(F32)
    neg = np.where((gt_depth - off).astype(F32) < 0, gt_depth, half).astype(F32)
    start = (gt - (neg[:, None] * dirs).astype(F32)).astype(F32)
    start_depth = _depth(start, cam_loc[None])
    distances = (np.zeros_like(gt_depth) + F32(truncated_dist * 2)).astype(F32)
    fine = (_stratified(t_fine.shape[-1], distances, np.asarray(t_fine, F32)) + start_depth[:, None]).astype(F32)
    coarse = _stratified(t_coarse.shape[-1], gt_depth, np.asarray(t_coarse, F32))
    dists = np.sort(np.concatenate([fine, coarse], -1), -1)
    pts = (cam_loc[None, None] + (dists[:, :, None] * dirs[:, None, :]).astype(F32)).astype(F32)
    return pts, dists, gt_depth, dirs, cam_loc


def sdf_loss(pts, pred_sdf, gt_pts, cam_loc, nbr_pts, nbr_mask, ray_mask, truncated_dist):
    """compute_sdf_loss + calculate_loss: (loss, d loss / d pred_sdf [n,S])"""
    pts, pred = np.asarray(pts, F32), np.asarray(pred_sdf, F32)
    td = F32(truncated_dist)
    gt_depth = _depth(np.asarray(gt_pts, F32), cam_loc[None])[:, None]
    depth = _depth(pts, cam_loc[None, None])
    gt_sdf = np.clip((gt_depth - depth).astype(F32), -td, td)
    valid = gt_sdf > F32(max(-truncated_dist * 0.5, -0.05))
    d = (np.asarray(nbr_pts, F32)[:, None, :, :] - pts[:, :, None, :]).astype(F32)      # [n,S,Nn,3]
    dist = np.sqrt(((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(F32) + d[..., 2] * d[..., 2]).astype(F32)).astype(F32)
    dist = np.where(np.asarray(nbr_mask)[:, None, :] != 0, dist, F32(10000.0))
    nearest = dist.min(-1)
    sign = np.where(gt_sdf > 0, F32(1.0), F32(-1.0))
    target = np.clip((nearest * sign).astype(F32), -td, td)
    l1 = np.abs((pred - target).astype(F32)) * valid
    nvp = F32(np.asarray(ray_mask, F32).sum(dtype=F32) + F32(1e-4))
    m = np.asarray(ray_mask, F32)[:, None]
    loss = F32((l1 * m).sum(dtype=np.float64)) / nvp
    grad = (np.sign((pred - target).astype(F32)) * valid * m / nvp).astype(F32)
    return float(loss), grad, target, valid
