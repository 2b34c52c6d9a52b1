"""Drop-in ``MSDeformAttn`` module backed by the sm_100a kernels.

Mirrors the reference module's interface
(/root/reference/src/models/detrpose/ms_deform_attn.py:196-513): identical
constructor keywords and defaults (:197-203), identical parameter names and
shapes (``sampling_offsets.{weight,bias}``, ``attention_weights.{weight,bias}``,
:245-246) so reference ``state_dict``s load unchanged, identical initialisation
(:293-315) and the same ``forward(query, reference_points, value,
input_spatial_shapes)`` (:359) taking the list of strided per-level value views.

The fork-only optional branches (modulation, region sampling, global context,
grouped offsets, grid attention, energy) are outside the kernel contract: this
class refuses them at construction.  To keep a reference model that uses them
working, patch the reference module instead (``detrpose_b200.patch.install``),
which routes those configurations to the reference's own code explicitly.
"""
from __future__ import annotations

import math

import torch
from torch import nn
import torch.nn.functional as F

from . import functional as MF

__all__ = ["MSDeformAttn"]

_UNSUPPORTED = ("use_modulation", "use_region_sampling", "use_global_context", "use_grouped_offsets",
                "use_grid_attention", "is_energy")


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, use_4D_normalizer=False,
                 use_modulation=False, use_region_sampling=False, region_kernel_size=1,
                 use_global_context=False, use_grouped_offsets=False, num_groups=1,
                 use_grid_attention=False, grid_num_points=16, use_grid_offsets=False,
                 use_grid_fusion=True, is_energy=False):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError('d_model must be divisible by n_heads, but got {} and {}'.format(d_model, n_heads))
        flags = dict(use_modulation=use_modulation, use_region_sampling=use_region_sampling,
                     use_global_context=use_global_context, use_grouped_offsets=use_grouped_offsets,
                     use_grid_attention=use_grid_attention, is_energy=is_energy)
        enabled = [k for k in _UNSUPPORTED if flags[k]]
        if enabled:
            raise NotImplementedError(
                f"detrpose_b200.MSDeformAttn implements the baseline sampling core only; {enabled} are "
                "experimental branches of the reference. Use detrpose_b200.patch.install(reference_module) "
                "to keep the reference class and accelerate just its core.")
        if (d_model // n_heads) % 8 != 0:
            raise ValueError(f"head dim {d_model // n_heads} must be a multiple of 8 for the sm_100a kernels")

        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points
        self.is_energy = False
        self.num_groups = 1
        self.use_4D_normalizer = use_4D_normalizer
        self.region_kernel_size = int(region_kernel_size)
        self.fuse_prologue = True          # fused softmax/location kernel (fp32 inputs, 2-D reference points)

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self._reset_parameters()

    def _reset_parameters(self):
        """Reference init (ms_deform_attn.py:293-315): zero weights; the offset bias points each
        head along its own direction, the same for every level and point, and is zero unless
        n_points is a multiple of 4; uniform attention."""
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            angle = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
            direction = torch.stack([angle.cos(), angle.sin()], -1)
            direction = direction / direction.abs().max(-1, keepdim=True)[0]
            bias = direction.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
            self.sampling_offsets.bias = nn.Parameter(bias.reshape(-1))
            if self.n_points % 4 != 0:
                self.sampling_offsets.bias.zero_()
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()

    def forward(self, query, reference_points, value, input_spatial_shapes):
        """
        query: (N, Len_q, C)
        reference_points: (N, nq, n_levels|1, K, 2 or 4) -- transposed/flattened to (N, Len_q, n_levels|1, 2|4) as the
                          reference does (:412)
        value: list of per-level tensors (N * n_heads, d_per_head, H_l * W_l) with any strides
               (or one (N, S, C) tensor: the zero-copy path)
        input_spatial_shapes: list of (H_l, W_l)
        """
        N, Len_q, _ = query.shape
        H, L, P = self.n_heads, self.n_levels, self.n_points
        shapes = MF._shapes_tuple(input_spatial_shapes)

        offsets = self.sampling_offsets(query)
        logits = self.attention_weights(query)
        ref = torch.transpose(reference_points, 2, 3).flatten(1, 2)
        last = ref.shape[-1]
        if last not in (2, 4):
            raise ValueError('Last dim of reference_points must be 2 or 4, but get {} instead.'.format(last))

        if last == 2 and self.fuse_prologue and offsets.is_cuda:
            # softmax + locations + sampling in one launch per direction (row f1), in fp32: locations are
            # bit-identical to the reference's ops for fp32 inputs; under autocast (bf16 / fp16 Linear outputs)
            # the arithmetic is done in fp32 instead of the reduced precision, which only moves the locations
            # closer to the fp32 result
            return MF.ms_deform_attn_fused(value, shapes, offsets, logits, ref, n_heads=H, n_levels=L, n_points=P)
        else:
            offsets = offsets.view(N, Len_q, H, L, P, 2)
            weights = F.softmax(logits.view(N, Len_q, H, L * P), -1).view(N, Len_q, H, L, P)
            if last == 2:
                normalizer = MF.level_normalizer(shapes, query.device)
                locations = ref[:, :, None, :, None, :] + offsets / normalizer.reshape(1, 1, 1, L, 1, 2)
            elif self.use_4D_normalizer:
                normalizer = MF.level_normalizer(shapes, query.device)
                locations = ref[:, :, None, :, None, :2] \
                    + offsets / normalizer[None, None, None, :, None, :] * ref[:, :, None, :, None, 2:] * 0.5
            else:
                locations = ref[:, :, None, :, None, :2] + offsets / P * ref[:, :, None, :, None, 2:] * 0.5
        return MF.ms_deform_attn_core(value, shapes, locations, weights, n_heads=H)
