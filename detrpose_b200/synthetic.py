"""Synthetic inputs and byte accounting for the sampling core.

Shapes follow the reference's configs (Appendix A of SURVEY.md):
``configs/detrpose/include/detrpose_hgnetv2.py:29-83`` (L: d_model 256, 8 heads,
3 levels, 4 points, 6 decoder layers, 60 queries x 18 slots = Len_q 1080),
``detrpose_hgnetv2_n.py:42-58`` (N: d_model 128, 2 levels, 6 points),
``_s.py:42-47`` (S: 3 decoder layers), ``_x.py:42-47`` (X: d_model 384).
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch

__all__ = ["WORKLOADS", "make_inputs", "algorithmic_bytes", "pyramid_size"]

_P3 = ((80, 80), (40, 40), (20, 20))

WORKLOADS: Dict[str, dict] = {
    # name: heads, head dim, pyramid at 640x640, points, Len_q at inference, decoder layers
    "detrpose_n": dict(H=8, Dh=16, shapes=((40, 40), (20, 20)), P=6, Lq=1080, layers=3),
    "detrpose_s": dict(H=8, Dh=32, shapes=_P3, P=4, Lq=1080, layers=3),
    "detrpose_m": dict(H=8, Dh=32, shapes=_P3, P=4, Lq=1080, layers=4),
    "detrpose_l": dict(H=8, Dh=32, shapes=_P3, P=4, Lq=1080, layers=6),
    "detrpose_x": dict(H=8, Dh=48, shapes=_P3, P=4, Lq=1080, layers=6),
    # BASELINE.json configs[4]: standalone 4-level sweep
    "sweep4": dict(H=8, Dh=32, shapes=((80, 80), (40, 40), (20, 20), (10, 10)), P=4, Lq=900, layers=1),
}


def pyramid_size(shapes: Sequence[Tuple[int, int]]) -> int:
    return sum(int(h) * int(w) for h, w in shapes)


def make_inputs(N: int, Lq: int, H: int, Dh: int, shapes, P: int, *, seed: int = 0, device="cpu",
                value_dtype: torch.dtype = torch.float32, grad_dtype: torch.dtype = None,
                degenerate: bool = False, offset_px_std: float = 2.0,
                clip: Tuple[float, float] = (-0.1, 1.1)) -> dict:
    """Seeded inputs for one call of the core.

    ``memory`` ~ N(0,1) ``(N, S, H*Dh)`` (the reference's encoder output layout,
    transformer.py:1158-1177); one reference point per query ~ U(0,1) shared by all
    heads/levels/points (as DETRPose passes 2-D keypoint references); pixel offsets
    ~ N(0, offset_px_std^2) divided by (W_l, H_l); locations clipped to ``clip`` so a
    few percent of the corners fall outside the map; attention = softmax(N(0,1)) over
    L*P; ``grad_out`` ~ N(0,1).  ``degenerate=True`` repeats one offset for all P
    points and uses uniform attention -- what the reference's zero-initialised
    Linears produce (ms_deform_attn.py:293-315), the worst case for backward
    contention.
    """
    L = len(shapes)
    S = pyramid_size(shapes)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    f32 = dict(dtype=torch.float32, device=device, generator=gen)
    memory = torch.randn((N, S, H * Dh), **f32).to(value_dtype)
    ref = torch.rand((N, Lq, 1, 1, 1, 2), **f32)
    off_px = torch.randn((N, Lq, H, L, 1 if degenerate else P, 2), **f32) * offset_px_std
    if degenerate:
        off_px = off_px.expand(N, Lq, H, L, P, 2)
    norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32, device=device).view(1, 1, 1, L, 1, 2)
    loc = (ref + off_px / norm).clamp_(clip[0], clip[1]).contiguous()
    if degenerate:
        attn = torch.full((N, Lq, H, L, P), 1.0 / (L * P), dtype=torch.float32, device=device)
    else:
        attn = torch.softmax(torch.randn((N, Lq, H, L * P), **f32), -1).view(N, Lq, H, L, P).contiguous()
    grad_out = torch.randn((N, Lq, H * Dh), **f32).to(grad_dtype or value_dtype)
    return dict(memory=memory, locations=loc, attention=attn, grad_out=grad_out,
                shapes=tuple((int(h), int(w)) for h, w in shapes), N=N, Lq=Lq, H=H, Dh=Dh, L=L, P=P, S=S)


def algorithmic_bytes(N: int, Lq: int, H: int, Dh: int, shapes, P: int, *, e_v: int, e_o: int,
                      e_l: int = 4, e_g: int = 4) -> Tuple[int, int]:
    """Algorithmic bytes of one forward and one backward call:

    B_fwd = e_v*N*S*C + e_l*3*N*Lq*H*L*P + e_o*N*Lq*C
    B_bwd = (e_v + e_g)*N*S*C + e_l*6*N*Lq*H*L*P + e_o*N*Lq*C
    """
    L, S, C = len(shapes), pyramid_size(shapes), H * Dh
    samples = N * Lq * H * L * P
    b_fwd = e_v * N * S * C + e_l * 3 * samples + e_o * N * Lq * C
    b_bwd = (e_v + e_g) * N * S * C + e_l * 6 * samples + e_o * N * Lq * C
    return b_fwd, b_bwd
