"""Randomised stress of the kernels against the reference's op sequence on the same GPU (tools/sweep.py's
restatement): odd shapes, 1..6 levels, 1..8 points, head dims 8..64, long query sets (several backward
chunks), fp32 and bf16 value, both hand-over layouts.  Prints one line per failure and a summary."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import detrpose_b200 as dp                                    # noqa: E402
from detrpose_b200 import synthetic, _lib                     # noqa: E402
from sweep import reference_ops                               # noqa: E402


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=150)
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--fused-trials", type=int, default=100)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    dev = "cuda:0"
    fails = 0
    for t in range(args.trials):
        N = int(rng.integers(1, 5)); H = int(rng.choice([1, 2, 4, 8])); Dh = int(rng.choice([8, 16, 24, 32, 48, 64]))
        L = int(rng.integers(1, 7)); P = int(rng.choice([1, 2, 3, 4, 5, 6, 8]))
        Lq = int(rng.choice([1, 5, 64, 333, 1080, 1584, 2500]))
        shapes = tuple((int(rng.integers(1, 90)), int(rng.integers(1, 90))) for _ in range(L))
        bf16 = bool(rng.integers(0, 2)); as_list = bool(rng.integers(0, 2))
        inp = synthetic.make_inputs(N, Lq, H, Dh, shapes, P, seed=7000 + t, device=dev, clip=(-0.4, 1.4),
                                    offset_px_std=float(rng.choice([0.5, 2.0, 8.0])))
        mem32 = inp["memory"]
        mem = (mem32.bfloat16() if bf16 else mem32).requires_grad_(True)
        memr = mem.detach().float().requires_grad_(True)
        loc = inp["locations"].requires_grad_(True); att = inp["attention"].requires_grad_(True)
        locr = inp["locations"].clone().requires_grad_(True); attr = inp["attention"].clone().requires_grad_(True)
        S = mem.shape[1]
        sizes = [h * w for h, w in shapes]

        def vlist(m):
            return list(m.view(N, S, H, Dh).permute(0, 2, 3, 1).flatten(0, 1).split(sizes, dim=-1))
        tag = f"trial {t}: N={N} H={H} Dh={Dh} L={L} P={P} Lq={Lq} bf16={bf16} list={as_list} shapes={shapes}"
        try:
            out = dp.ms_deform_attn_core(vlist(mem) if as_list else mem, shapes, loc, att, n_heads=H)
            go = inp["grad_out"].to(out.dtype)                         # the reference gets the same (rounded) values
            g = torch.autograd.grad(out, [mem, loc, att], go)
        except Exception as exc:                                       # noqa: BLE001
            print("EXC", tag, repr(exc)[:200]); fails += 1; continue
        ref = reference_ops(vlist(memr), shapes, locr, attr)
        gr = torch.autograd.grad(ref, [memr, locr, attr], go.float())
        ia, _ = dp.sample_indices(loc.detach(), shapes, coord_mode=_lib.COORD_UNFUSED)
        ib, _ = dp.sample_indices(loc.detach(), shapes, coord_mode=_lib.COORD_FMA)
        keep = (ia == ib).all(-1, keepdim=True).float()
        tol_v = 2.0 ** -8 if bf16 else 1e-5
        errs = {"out": (rel(out, ref), tol_v), "grad_value": (rel(g[0], gr[0]), tol_v),
                "grad_loc": (rel(g[1] * keep, gr[1] * keep), 1e-5), "grad_attn": (rel(g[2], gr[2]), 1e-5)}
        bad = {k: v for k, (v, tol) in errs.items() if not v <= tol}
        if bad:
            print("FAIL", tag, bad); fails += 1
    print(f"stress: {args.trials - fails} / {args.trials} trials ok")

    # ---- the same with the prologue fused in (row f1): from offsets / logits / reference points ----
    ffails = 0
    for t in range(args.fused_trials):
        N = int(rng.integers(1, 4)); H = int(rng.choice([1, 2, 4, 8])); Dh = int(rng.choice([8, 16, 24, 32, 48, 64]))
        L = int(rng.integers(1, 6)); P = int(rng.choice([1, 2, 3, 4, 5, 6, 8]))
        Lq = int(rng.choice([1, 5, 64, 333, 1080, 1584, 2500]))
        shapes = tuple((int(rng.integers(1, 90)), int(rng.integers(1, 90))) for _ in range(L))
        bf16 = bool(rng.integers(0, 2)); ref_levels = int(rng.choice([1, L]))
        g = torch.Generator(device=dev).manual_seed(9000 + t)
        S = sum(h * w for h, w in shapes)
        mem32 = torch.randn(N, S, H * Dh, device=dev, generator=g)
        mem = (mem32.bfloat16() if bf16 else mem32).requires_grad_(True)
        memr = mem.detach().float().requires_grad_(True)
        std = float(rng.choice([0.5, 2.0, 8.0]))
        off = (std * torch.randn(N, Lq, H * L * P * 2, device=dev, generator=g)).requires_grad_(True)
        lg = (3.0 * torch.randn(N, Lq, H * L * P, device=dev, generator=g)).requires_grad_(True)
        rp = (torch.rand(N, Lq, ref_levels, 2, device=dev, generator=g) * 1.6 - 0.3).requires_grad_(True)
        offr, lgr, rpr = (x.detach().clone().requires_grad_(True) for x in (off, lg, rp))
        sizes = [h * w for h, w in shapes]
        tag = f"fused trial {t}: N={N} H={H} Dh={Dh} L={L} P={P} Lq={Lq} bf16={bf16} ref_levels={ref_levels} shapes={shapes}"
        try:
            out = dp.functional.ms_deform_attn_fused(mem, shapes, off, lg, rp, n_heads=H, n_levels=L, n_points=P)
            go = torch.randn(out.shape, device=dev, generator=g).to(out.dtype)
            gm, goff, glg, grp = torch.autograd.grad(out, [mem, off, lg, rp], go)
        except Exception as exc:                                       # noqa: BLE001
            print("EXC", tag, repr(exc)[:200]); ffails += 1; continue
        w = torch.softmax(lgr.view(N, Lq, H, L * P), -1).view(N, Lq, H, L, P)
        norm = torch.tensor([[wd, h] for h, wd in shapes], dtype=torch.float32, device=dev).view(1, 1, 1, L, 1, 2)
        locr = rpr[:, :, None, :, None, :] + offr.view(N, Lq, H, L, P, 2) / norm
        ref = reference_ops(list(memr.view(N, S, H, Dh).permute(0, 2, 3, 1).flatten(0, 1).split(sizes, dim=-1)),
                            shapes, locr, w)
        gr = torch.autograd.grad(ref, [memr, offr, lgr, rpr], go.float())
        ia, _ = dp.sample_indices(locr.detach(), shapes, coord_mode=_lib.COORD_UNFUSED)
        ib, _ = dp.sample_indices(locr.detach(), shapes, coord_mode=_lib.COORD_FMA)
        keep = (ia == ib).all(-1, keepdim=True).float().reshape(N, Lq, -1, 1).expand(-1, -1, -1, 2).reshape(N, Lq, -1)
        tol_v = 2.0 ** -8 if bf16 else 1e-5
        kr = (ia == ib).all(-1).float()                                 # (N, Lq, H, L, P)
        keep_ref = kr.amin(dim=(2, 4)) if ref_levels == L else kr.amin(dim=(2, 3, 4)).unsqueeze(-1)
        errs = {"out": (rel(out, ref), tol_v), "grad_value": (rel(gm, gr[0]), tol_v),
                "grad_offsets": (rel(goff * keep, gr[1] * keep), 1e-5), "grad_logits": (rel(glg, gr[2]), 1e-5),
                "grad_ref": (rel(grp * keep_ref.unsqueeze(-1), gr[3] * keep_ref.unsqueeze(-1)), 2e-5)}
        bad = {k: v for k, (v, tol) in errs.items() if not v <= tol}
        if bad:
            print("FAIL", tag, bad); ffails += 1
    print(f"stress (fused prologue): {args.fused_trials - ffails} / {args.fused_trials} trials ok")
    sys.exit(1 if (fails or ffails) else 0)


if __name__ == "__main__":
    main()
