"""CPU oracle for the multi-scale deformable attention sampling core (numpy).

TEST INFRASTRUCTURE ONLY.  Nothing under ``detrpose_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or the
timed CPU baseline, never as the product path.

What is restated
----------------
The reference hot path is pure PyTorch and its arithmetic lives in a
third-party dependency that is NOT vendored in /root/reference:

* glue: ``ms_deform_attn_core_pytorch`` --
  /root/reference/src/models/detrpose/ms_deform_attn.py:145-193
  (``2*loc-1`` :161, per-level ``F.grid_sample(bilinear, zeros,
  align_corners=False)`` :178, ``cat`` :184, ``* attention_weights`` and
  ``sum(-1)`` :192, ``transpose`` :193);
* sampler: ATen ``grid_sampler_2d`` / ``grid_sampler_2d_backward`` of
  PyTorch (unpinned by the reference; 2.11.0+cu128 in this image).  Published
  formulas: torch/include/ATen/native/GridSampler.h:27-35 (unnormalise,
  ``((g + 1) * size - 1) / 2``), :205 (``within_bounds_2d`` -- each corner is
  dropped individually when it is outside the map), :238-243 (``safe_add_2d``
  -- the backward never writes an out-of-range corner) and the bilinear
  corner weights ``nw = (x1-x)(y1-y), ne = (x-x0)(y1-y), sw = (x1-x)(y-y0),
  se = (x-x0)(y-y0)`` with ``x0 = floor(x), x1 = x0+1``.

Parity pinning: the reference ships no tests or golden vectors,
so this restatement is pinned against *outputs of the reference itself*,
generated in the build container by ``tests/golden/make_golden.py`` (which
imports the reference file by path) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function here against them.

Layouts follow the reference exactly:

* ``value``: list of L arrays ``(N*H, Dh, H_l*W_l)``;
* ``spatial_shapes``: list of ``(H_l, W_l)``;
* ``sampling_locations``: ``(N, Lq, H, L, P, 2)``, last axis ``(x, y)``
  normalised to [0, 1] (may fall outside);
* ``attention_weights``: ``(N, Lq, H, L, P)``;
* output: ``(N, Lq, H*Dh)``.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "level_start_index",
    "pixel_coords",
    "sample_indices",
    "msda_forward",
    "msda_backward",
]


def level_start_index(spatial_shapes):
    """Exclusive running sum of H_l*W_l.

    The reference never materialises this (it splits ``memory`` with
    ``split_sizes`` -- transformer.py:1172-1173,1286); the flattened pyramid
    offset of level l is by construction the sum of the earlier extents.
    """
    starts, acc = [], 0
    for h, w in spatial_shapes:
        starts.append(acc)
        acc += int(h) * int(w)
    return np.asarray(starts, dtype=np.int64)


def pixel_coords(loc, size, dtype=np.float32):
    """Normalised location -> continuous pixel coordinate, op by op.

    Follows ms_deform_attn.py:161 (``g = 2*loc - 1``) then
    GridSampler.h:33-35 (``((g + 1) * size - 1) / 2``) with one rounding per
    arithmetic op in ``dtype`` (no fused multiply-add), which is what the
    elementwise torch ops of the reference do.
    """
    t = np.dtype(dtype).type
    loc = np.asarray(loc, dtype=dtype)
    g = t(2) * loc - t(1)
    u = (g + t(1)) * t(size)
    return (u - t(1)) / t(2)


def sample_indices(sampling_locations, spatial_shapes, dtype=np.float32):
    """Integer top-left corner ``(y0, x0)`` of every sample, ``int32``.

    Returns an array ``(N, Lq, H, L, P, 2)`` with last axis ``(y0, x0)``; these
    are the ``floor`` results ATen's bilinear path takes (``ix_nw``/``iy_nw``),
    clipped to ``[-2, size + 1]``: anything further out has no valid corner and
    the clip keeps the integer conversion defined for absurd locations.
    """
    loc = np.asarray(sampling_locations, dtype=dtype)
    out = np.empty(loc.shape, dtype=np.int32)
    for l, (h, w) in enumerate(spatial_shapes):
        x = pixel_coords(loc[:, :, :, l, :, 0], w, dtype)
        y = pixel_coords(loc[:, :, :, l, :, 1], h, dtype)
        out[:, :, :, l, :, 0] = np.clip(np.floor(y), -2, h + 1).astype(np.int32)
        out[:, :, :, l, :, 1] = np.clip(np.floor(x), -2, w + 1).astype(np.int32)
    return out


def _corners(loc_l, h, w, dtype, coord_dtype=None):
    """Per-sample corner indices, validity and bilinear weights for one level.

    loc_l: (..., 2).  Returns ``(idx[4], valid[4], wgt[4], x, y, x0, y0)`` in the
    ATen corner order nw, ne, sw, se.  ``coord_dtype`` (default ``dtype``) is the
    precision of the pixel-coordinate chain: pass ``np.float32`` with
    ``dtype=np.float64`` to get a high-precision arbiter that takes the *same
    floor decisions* as an fp32 run (the gradient w.r.t. the location is
    discontinuous where a coordinate is integral).
    """
    t = np.dtype(dtype).type
    cd = dtype if coord_dtype is None else coord_dtype
    x = pixel_coords(np.asarray(loc_l[..., 0], dtype=cd), w, cd).astype(dtype)
    y = pixel_coords(np.asarray(loc_l[..., 1], dtype=cd), h, cd).astype(dtype)
    x0f = np.floor(x)
    y0f = np.floor(y)
    x1f = x0f + t(1)
    y1f = y0f + t(1)
    wgt = [
        (x1f - x) * (y1f - y),   # nw
        (x - x0f) * (y1f - y),   # ne
        (x1f - x) * (y - y0f),   # sw
        (x - x0f) * (y - y0f),   # se
    ]
    # clip before the int cast so that absurd locations cannot overflow
    x0 = np.clip(x0f, -2, w + 1).astype(np.int64)
    y0 = np.clip(y0f, -2, h + 1).astype(np.int64)
    x1 = x0 + 1
    y1 = y0 + 1
    idx, valid = [], []
    for yy, xx in ((y0, x0), (y0, x1), (y1, x0), (y1, x1)):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)  # GridSampler.h:205
        valid.append(ok)
        idx.append(np.where(ok, yy * w + xx, 0))
    return idx, valid, wgt, x, y, x0f, y0f


def msda_forward(value, spatial_shapes, sampling_locations, attention_weights,
                 dtype=np.float32, coord_dtype=None):
    """``ms_deform_attn_core_pytorch`` (ms_deform_attn.py:145-193), flags off.

    out[n, q, h*Dh + c] = sum_{l,p} A[n,q,h,l,p] * bilinear_zero_pad(value_l[n*H+h, c], loc)
    """
    loc = np.asarray(sampling_locations, dtype=dtype if coord_dtype is None else coord_dtype)
    att = np.asarray(attention_weights, dtype=dtype)
    N, Lq, H, L, P, _ = loc.shape
    Dh = value[0].shape[1]
    out = np.zeros((N * H, Dh, Lq), dtype=dtype)
    # (N, Lq, H, ...) -> (N*H, Lq, ...) exactly as :162 / :186 do
    loc_nh = loc.transpose(0, 2, 1, 3, 4, 5).reshape(N * H, Lq, L, P, 2)
    att_nh = att.transpose(0, 2, 1, 3, 4).reshape(N * H, Lq, L, P)
    for l, (h, w) in enumerate(spatial_shapes):
        v = np.asarray(value[l], dtype=dtype)               # (N*H, Dh, h*w)
        idx, valid, wgt, *_ = _corners(loc_nh[:, :, l], h, w, dtype, coord_dtype)
        sampled = np.zeros((N * H, Dh, Lq, P), dtype=dtype)
        for k in range(4):
            flat = idx[k].reshape(N * H, 1, Lq * P)
            g = np.take_along_axis(v, np.broadcast_to(flat, (N * H, Dh, Lq * P)), axis=2)
            g = g.reshape(N * H, Dh, Lq, P)
            wk = np.where(valid[k], wgt[k], 0).astype(dtype)
            sampled += g * wk[:, None]
        out += (sampled * att_nh[:, None, :, l]).sum(-1)
    return out.reshape(N, H * Dh, Lq).transpose(0, 2, 1).copy()


def msda_backward(value, spatial_shapes, sampling_locations, attention_weights,
                  grad_output, dtype=np.float64, coord_dtype=None):
    """Analytic gradients of :func:`msda_forward`.

    Follows ATen ``grid_sampler_2d_backward`` (corner weights' derivatives,
    ``gix_mult = W/2`` from GridSampler.h:51-52) chained through
    ``g = 2*loc - 1`` (ms_deform_attn.py:161, factor 2) and the
    ``* attention_weights`` / ``sum`` of :192.

    Returns ``(grad_value_list, grad_sampling_locations, grad_attention_weights)``
    with the shapes of the corresponding inputs.
    """
    loc = np.asarray(sampling_locations, dtype=dtype if coord_dtype is None else coord_dtype)
    att = np.asarray(attention_weights, dtype=dtype)
    go = np.asarray(grad_output, dtype=dtype)
    N, Lq, H, L, P, _ = loc.shape
    Dh = value[0].shape[1]
    go_nh = go.reshape(N, Lq, H, Dh).transpose(0, 2, 3, 1).reshape(N * H, Dh, Lq)
    loc_nh = loc.transpose(0, 2, 1, 3, 4, 5).reshape(N * H, Lq, L, P, 2)
    att_nh = att.transpose(0, 2, 1, 3, 4).reshape(N * H, Lq, L, P)
    g_val = []
    g_loc = np.zeros(loc_nh.shape, dtype=dtype)
    g_att = np.zeros(att_nh.shape, dtype=dtype)
    rows = np.arange(N * H)[:, None, None]
    chans = np.arange(Dh)[None, :, None]
    for l, (h, w) in enumerate(spatial_shapes):
        v = np.asarray(value[l], dtype=dtype)
        idx, valid, wgt, x, y, x0f, y0f = _corners(loc_nh[:, :, l], h, w, dtype, coord_dtype)
        gv = np.zeros_like(v)
        dots = []
        for k in range(4):
            flat = idx[k].reshape(N * H, 1, Lq * P)
            c = np.take_along_axis(v, np.broadcast_to(flat, (N * H, Dh, Lq * P)), axis=2)
            c = c.reshape(N * H, Dh, Lq, P) * valid[k][:, None]
            dots.append((c * go_nh[:, :, :, None]).sum(1))          # (N*H, Lq, P)
            # d out / d value: A * w_k * grad_out, dropped corners never written
            contrib = (np.where(valid[k], wgt[k], 0) * att_nh[:, :, l])[:, None] * go_nh[:, :, :, None]
            np.add.at(gv, (rows, chans, idx[k].reshape(N * H, 1, Lq * P)),
                      contrib.reshape(N * H, Dh, Lq * P))
        d_nw, d_ne, d_sw, d_se = dots
        g_att[:, :, l] = sum(np.where(valid[k], wgt[k], 0) * dots[k] for k in range(4))
        tx = x - x0f
        ty = y - y0f
        one = np.dtype(dtype).type(1)
        dx = (d_ne - d_nw) * (one - ty) + (d_se - d_sw) * ty
        dy = (d_sw - d_nw) * (one - tx) + (d_se - d_ne) * tx
        g_loc[:, :, l, :, 0] = att_nh[:, :, l] * dx * w
        g_loc[:, :, l, :, 1] = att_nh[:, :, l] * dy * h
        g_val.append(gv)
    g_loc = g_loc.reshape(N, H, Lq, L, P, 2).transpose(0, 2, 1, 3, 4, 5).copy()
    g_att = g_att.reshape(N, H, Lq, L, P).transpose(0, 2, 1, 3, 4).copy()
    return g_val, g_loc, g_att
