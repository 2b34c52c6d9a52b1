"""CPU oracle (TEST INFRASTRUCTURE, not product code) for the LQE sampler.

Restates /root/reference/src/models/detrpose/transformer.py:274-288:

    v     = grid_sample(feat, 2*poses - 1, bilinear, zeros, align_corners=False).permute(0, 2, 3, 1)  (:278-280)
    top   = v.topk(k, dim=-1)[0]                                                                      (:282)
    stat  = cat([top, top.mean(-1, keepdim=True)], -1)                                                (:284)
    score = scores + MLP(stat.reshape(B, L, -1))                                                      (:285-288)

with ATen's bilinear sampler formulas (GridSampler.h:27-35 unnormalise, :205 per-corner bounds) shared
with oracle/msda_numpy.py, and the analytic backward of the statistics.  Pinned against the real
reference class (tests/golden/lqe_*.npz from tests/golden/make_golden_blocks.py).
Only tests/, __graft_entry__.smoke() and bench-side baselines may import this module.
"""
from __future__ import annotations

import numpy as np

from .msda_numpy import pixel_coords

__all__ = ["sample_all_channels", "lqe_statistics", "lqe_statistics_backward", "mlp_forward", "lqe_forward"]


def _corners(poses, hf, wf, coord_dtype):
    x = pixel_coords(poses[..., 0], wf, coord_dtype).astype(np.float64)
    y = pixel_coords(poses[..., 1], hf, coord_dtype).astype(np.float64)
    x = np.clip(x, -2.0, wf + 1.0)
    y = np.clip(y, -2.0, hf + 1.0)
    x0, y0 = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
    return x, y, x0, y0


def sample_all_channels(feat, poses, dtype=np.float64, coord_dtype=np.float32):
    """feat (B, C, Hf, Wf), poses (B, P, 2) -> bilinear samples (B, P, C), zero padding per corner."""
    feat = np.asarray(feat, dtype)
    poses = np.asarray(poses)
    b, c, hf, wf = feat.shape
    x, y, x0, y0 = _corners(poses, hf, wf, coord_dtype)
    out = np.zeros(poses.shape[:2] + (c,), dtype)
    bi = np.arange(b)[:, None]
    for dy in (0, 1):
        for dx in (0, 1):
            xi, yi = x0 + dx, y0 + dy
            w = ((x - x0) if dx else (x0 + 1 - x)) * ((y - y0) if dy else (y0 + 1 - y))
            ok = (xi >= 0) & (xi < wf) & (yi >= 0) & (yi < hf)
            v = feat[bi, :, np.clip(yi, 0, hf - 1), np.clip(xi, 0, wf - 1)]          # (B, P, C)
            out += np.where(ok, w, 0.0).astype(dtype)[..., None] * v
    return out


def lqe_statistics(feat, poses, k, dtype=np.float64, coord_dtype=np.float32):
    """Returns (stat (B, P, k+1), idx (B, P, k)): top-k over channels (descending) and their mean."""
    v = sample_all_channels(feat, poses, dtype, coord_dtype)
    idx = np.argsort(-v, axis=-1, kind="stable")[..., :k]
    top = np.take_along_axis(v, idx, axis=-1)
    return np.concatenate([top, top.mean(-1, keepdims=True)], axis=-1), idx


def lqe_statistics_backward(feat, poses, k, grad_stat, dtype=np.float64, coord_dtype=np.float32):
    """Gradients of sum(stat * grad_stat) w.r.t. feat (B, C, Hf, Wf) and poses (B, P, 2)."""
    feat = np.asarray(feat, dtype)
    poses = np.asarray(poses)
    gs = np.asarray(grad_stat, dtype)
    b, c, hf, wf = feat.shape
    _, idx = lqe_statistics(feat, poses, k, dtype, coord_dtype)
    g = gs[..., :k] + gs[..., k:] / k                                                 # (B, P, k) onto the kept channels
    x, y, x0, y0 = _corners(poses, hf, wf, coord_dtype)
    g_feat = np.zeros_like(feat)
    g_pose = np.zeros(poses.shape, dtype)
    bi = np.broadcast_to(np.arange(b)[:, None, None], idx.shape)
    for dy in (0, 1):
        for dx in (0, 1):
            xi, yi = x0 + dx, y0 + dy
            wx = (x - x0) if dx else (x0 + 1 - x)
            wy = (y - y0) if dy else (y0 + 1 - y)
            ok = (xi >= 0) & (xi < wf) & (yi >= 0) & (yi < hf)
            xc, yc = np.clip(xi, 0, wf - 1), np.clip(yi, 0, hf - 1)
            w = np.where(ok, wx * wy, 0.0)[..., None]
            np.add.at(g_feat, (bi, idx, np.broadcast_to(yc[..., None], idx.shape),
                               np.broadcast_to(xc[..., None], idx.shape)), w * g)
            v = feat[bi, idx, yc[..., None], xc[..., None]] * np.where(ok, 1.0, 0.0)[..., None]   # (B, P, k)
            g_pose[..., 0] += ((1 if dx else -1) * wy * (v * g).sum(-1)) * wf
            g_pose[..., 1] += ((1 if dy else -1) * wx * (v * g).sum(-1)) * hf
    return g_feat, g_pose


def mlp_forward(x, weights, biases, dtype=np.float64):
    """The reference MLP (utils.py:75-87): Linear + ReLU ... Linear."""
    x = np.asarray(x, dtype)
    for i, (w, bias) in enumerate(zip(weights, biases)):
        x = x @ np.asarray(w, dtype).T + np.asarray(bias, dtype)
        if i < len(weights) - 1:
            x = np.maximum(x, 0)
    return x


def lqe_forward(scores, pred_poses, feat, k, weights, biases, num_body_points, dtype=np.float64):
    """LQE.forward (:274-288): scores (B, L, 1) + MLP(stat)."""
    pred_poses = np.asarray(pred_poses)
    b, l = pred_poses.shape[:2]
    stat, _ = lqe_statistics(feat, pred_poses.reshape(b, l * num_body_points, 2), k, dtype)
    return np.asarray(scores, dtype) + mlp_forward(stat.reshape(b, l, -1), weights, biases, dtype)
