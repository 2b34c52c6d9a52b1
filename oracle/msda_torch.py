"""CPU oracle, PyTorch-op form: the reference's own op sequence, restated.

TEST INFRASTRUCTURE ONLY (see ``oracle/msda_numpy.py`` for the rule: nothing
under ``detrpose_b200/`` imports ``oracle``).  This file exists because the
reference itself (/root/reference) does not travel to the GPU box, while the
north star asks for "the reference's CPU PyTorch path timed on the box's own
host cores".  The functions below issue the same ATen ops in the same order
as the reference, so timing them on CPU *is* timing the reference's CPU path,
and running them on ``cuda`` gives the device-side implementation the new
kernels have to beat.

Restates:

* ``core``  -- ``ms_deform_attn_core_pytorch`` with all optional flags off,
  /root/reference/src/models/detrpose/ms_deform_attn.py:145-193;
* ``locations_and_weights`` -- the pre-core part of ``MSDeformAttn.forward``,
  ms_deform_attn.py:385-393 (two Linears + softmax over L*P) and :412-416
  (reference-point transpose, ``offsets / [W_l, H_l]``, add);
* ``make_value_list`` -- the caller's value construction,
  /root/reference/src/models/detrpose/transformer.py:1285-1286.

Pinned against the real reference by ``tests/golden/make_golden.py`` +
``tests/test_oracle_golden.py`` (bit-identical on CPU: same ops, same order).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = ["core", "core_fwd_bwd", "locations_and_weights", "make_value_list"]


def make_value_list(memory: torch.Tensor, n_heads: int, spatial_shapes):
    """``memory (N, S, C)`` -> list of L views ``(N*H, Dh, H_l*W_l)``.

    transformer.py:1285-1286.  For N == 1 the result is a zero-copy
    channel-innermost view; for N > 1 ``flatten(0, 1)`` materialises a
    spatial-innermost buffer and the list holds strided views into it.
    """
    extents = [int(h) * int(w) for h, w in spatial_shapes]
    heads = memory.unflatten(2, (n_heads, -1))            # (N, S, H, Dh)
    heads = heads.permute(0, 2, 3, 1).flatten(0, 1)       # (N*H, Dh, S)
    return list(heads.split(extents, dim=-1))


def core(value, spatial_shapes, sampling_locations, attention_weights):
    """ms_deform_attn.py:145-193 with modulation / region / energy off."""
    n_batch, len_q, n_heads, n_levels, n_points, _ = sampling_locations.shape
    d_head = value[0].shape[1]
    grids = (2 * sampling_locations - 1).transpose(1, 2).flatten(0, 1)   # :161-162
    per_level = []
    for lvl, (h, w) in enumerate(spatial_shapes):                        # :165
        fmap = value[lvl].unflatten(2, (int(h), int(w)))                 # :166
        per_level.append(F.grid_sample(fmap, grids[:, :, lvl], mode="bilinear",
                                       padding_mode="zeros", align_corners=False))  # :178
    stacked = torch.cat(per_level, dim=-1)                               # :184
    weights = attention_weights.transpose(1, 2).reshape(
        n_batch * n_heads, 1, len_q, n_levels * n_points)                # :186
    summed = (stacked * weights).sum(-1)                                 # :192
    return summed.view(n_batch, n_heads * d_head, len_q).transpose(1, 2)  # :192-193


def core_fwd_bwd(value, spatial_shapes, sampling_locations, attention_weights, grad_output):
    """Forward + autograd backward; returns ``(out, [grad_value_l], grad_loc, grad_attn)``."""
    value = [v.detach().requires_grad_(True) for v in value]
    loc = sampling_locations.detach().requires_grad_(True)
    att = attention_weights.detach().requires_grad_(True)
    out = core(value, spatial_shapes, loc, att)
    grads = torch.autograd.grad(out, [*value, loc, att], grad_output)
    return out.detach(), list(grads[:len(value)]), grads[-2], grads[-1]


def locations_and_weights(query, reference_points, spatial_shapes, off_w, off_b, att_w, att_b,
                          n_heads, n_levels, n_points):
    """Pre-core math of ``MSDeformAttn.forward`` (2-D reference points).

    ms_deform_attn.py:385,390 (offset Linear + view), :392-393 (attention
    Linear, softmax over L*P), :412 (``transpose(2, 3).flatten(1, 2)``),
    :414-416 (int64 normaliser ``[W_l, H_l]``, divide, add).
    """
    n_batch, len_q, _ = query.shape
    offsets = F.linear(query, off_w, off_b).view(n_batch, len_q, n_heads, n_levels, n_points, 2)
    logits = F.linear(query, att_w, att_b).view(n_batch, len_q, n_heads, n_levels * n_points)
    weights = F.softmax(logits, -1).view(n_batch, len_q, n_heads, n_levels, n_points)
    ref = torch.transpose(reference_points, 2, 3).flatten(1, 2)           # (N, Lq, 1|L, 2)
    normalizer = torch.tensor(spatial_shapes, device=query.device).flip([1])
    normalizer = normalizer.reshape(1, 1, 1, n_levels, 1, 2)
    locations = ref[:, :, None, :, None, :] + offsets / normalizer
    return locations, weights
