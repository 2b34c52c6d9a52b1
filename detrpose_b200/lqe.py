"""Drop-in ``LQE``: the keypoint-quality head's sampler.

Mirrors /root/reference/src/models/detrpose/transformer.py:263-288 -- same constructor
``LQE(topk, hidden_dim, num_layers, num_body_points)``, the same parameters
(``reg_conf.layers.{i}.{weight,bias}``, last layer zero-initialised) and
``forward(scores, pred_poses, feat)``.  The sampled ``(B, C, L, 17)`` tensor, its permute and the
sort-based top-k of the reference are replaced by one kernel (``msda_b200_lqe_forward``) that emits the
``(B, L, 17 * (k + 1))`` statistics directly; the small MLP stays a library GEMM.  Differentiable
w.r.t. ``feat`` and ``pred_poses`` (``msda_b200_lqe_backward``).  No CPU path.
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F

from . import _lib
from . import functional as MF
from .functional import _code, _stream_ptr, _require_cuda, stats

__all__ = ["LQE", "lqe_statistics", "install_lqe", "uninstall_lqe"]

_SUPPORTED_C = (128, 256, 384, 512)


class _LQESample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, poses, k, coord_mode):
        _require_cuda(feat, "feat")
        _require_cuda(poses, "pred_poses")
        b, c, hf, wf = feat.shape
        if c not in _SUPPORTED_C or not 1 <= k <= 8:
            raise ValueError(f"LQE sampler supports C in {_SUPPORTED_C} and 1 <= topk <= 8, got C={c}, topk={k}")
        if feat.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"feat must be fp32 or bf16, got {feat.dtype}")
        p = poses.shape[1]
        pts = poses.detach().float().contiguous()
        stat = torch.empty((b, p, k + 1), dtype=torch.float32, device=feat.device)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        idx = torch.empty((b, p, k), dtype=torch.int32, device=feat.device) if need_grad else None
        lib = _lib.load()
        with torch.cuda.device(feat.device):
            rc = lib.msda_b200_lqe_forward(feat.data_ptr(), _code(feat.dtype), _lib.i64_array(feat.stride()),
                                           pts.data_ptr(), stat.data_ptr(), idx.data_ptr() if need_grad else None,
                                           b, c, hf, wf, p, k, coord_mode, _stream_ptr(feat.device))
        _lib.check(rc, "msda_b200_lqe_forward")
        stats["lqe_forward_launches"] = stats.get("lqe_forward_launches", 0) + 1
        if need_grad:
            ctx.save_for_backward(feat, pts, idx)
            ctx.k, ctx.coord_mode, ctx.poses_dtype = k, coord_mode, poses.dtype
        return stat

    @staticmethod
    def backward(ctx, grad_stat):
        feat, pts, idx = ctx.saved_tensors
        b, c, hf, wf = feat.shape
        p = pts.shape[1]
        need_feat, need_poses = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gs = grad_stat.float().contiguous()
        g_feat = None
        if need_feat:
            # fp32, same strides as feat; sparse scalar atomics land in a zero-filled buffer
            g_feat = torch.empty_strided(feat.size(), feat.stride(), dtype=torch.float32, device=feat.device).zero_()
        g_poses = torch.empty_like(pts) if need_poses else None
        lib = _lib.load()
        with torch.cuda.device(feat.device):
            rc = lib.msda_b200_lqe_backward(feat.data_ptr(), _code(feat.dtype), _lib.i64_array(feat.stride()),
                                            pts.data_ptr(), idx.data_ptr(), gs.data_ptr(),
                                            g_feat.data_ptr() if need_feat else None,
                                            g_poses.data_ptr() if need_poses else None,
                                            b, c, hf, wf, p, ctx.k, ctx.coord_mode, _stream_ptr(feat.device))
        _lib.check(rc, "msda_b200_lqe_backward")
        stats["lqe_backward_launches"] = stats.get("lqe_backward_launches", 0) + 1
        if need_feat and g_feat.dtype != feat.dtype:
            g_feat = g_feat.to(feat.dtype)
        if need_poses and g_poses.dtype != ctx.poses_dtype:
            g_poses = g_poses.to(ctx.poses_dtype)
        return g_feat, g_poses, None, None


def _dense(feat: torch.Tensor) -> bool:
    return feat.is_contiguous() or feat.is_contiguous(memory_format=torch.channels_last)


def lqe_statistics(feat: torch.Tensor, poses: torch.Tensor, topk: int, coord_mode: int = None) -> torch.Tensor:
    """``cat([v.topk(k)[0], v.topk(k)[0].mean(-1)])`` of ``v = grid_sample(feat, 2*poses-1)`` over channels
    (transformer.py:278-284).  ``feat`` (B, C, Hf, Wf) fp32/bf16, NCHW or channels-last, read in place;
    ``poses`` (B, P, 2) normalised (x, y).  Returns fp32 (B, P, topk + 1)."""
    if coord_mode is None:
        coord_mode = MF.get_default_coord_mode()
    if not _dense(feat):
        feat = feat.contiguous()
    return _LQESample.apply(feat, poses, int(topk), int(coord_mode))


def _lqe_forward(module, scores, pred_poses, feat):
    b, l = pred_poses.shape[:2]
    nb = module.num_body_points
    if feat.dtype not in (torch.float32, torch.bfloat16):
        feat = feat.float()                       # fp16 autocast: grid_sample runs in fp32 there as well
    stat = lqe_statistics(feat, pred_poses.reshape(b, l * nb, 2), module.k)
    quality_score = module.reg_conf(stat.view(b, l, nb * (module.k + 1)).to(scores.dtype))
    return scores + quality_score


class _MLP(nn.Module):
    """Same structure and parameter names as the reference's MLP (utils.py:75-87)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(i, o) for i, o in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = layer(x) if i == self.num_layers - 1 else F.relu(layer(x))
        return x


class LQE(nn.Module):
    def __init__(self, topk, hidden_dim, num_layers, num_body_points):
        super().__init__()
        self.k = topk
        self.hidden_dim = hidden_dim
        self.reg_conf = _MLP(num_body_points * (topk + 1), hidden_dim, 1, num_layers)
        # reference init (transformer.py:269-270): the head starts as "no correction"
        nn.init.constant_(self.reg_conf.layers[-1].weight.data, 0)
        nn.init.constant_(self.reg_conf.layers[-1].bias.data, 0)
        self.num_body_points = num_body_points

    def forward(self, scores, pred_poses, feat):
        return _lqe_forward(self, scores, pred_poses, feat)


_ORIGINAL_FORWARD = "_detrpose_b200_original_forward"


def install_lqe(reference_transformer_module) -> None:
    """Patch the reference's ``LQE.forward`` in place (parameters and ``state_dict`` untouched): CUDA inputs
    of a supported width run the fused sampler, anything else the reference's own forward."""
    cls = reference_transformer_module.LQE
    if hasattr(cls, _ORIGINAL_FORWARD):
        return
    original = cls.forward

    def forward(self, scores, pred_poses, feat):
        if feat.is_cuda and feat.dim() == 4 and feat.shape[1] in _SUPPORTED_C and 1 <= self.k <= 8:
            return _lqe_forward(self, scores, pred_poses, feat)
        return original(self, scores, pred_poses, feat)

    setattr(cls, _ORIGINAL_FORWARD, original)
    cls.forward = forward


def uninstall_lqe(reference_transformer_module) -> None:
    cls = reference_transformer_module.LQE
    original = getattr(cls, _ORIGINAL_FORWARD, None)
    if original is not None:
        cls.forward = original
        delattr(cls, _ORIGINAL_FORWARD)
