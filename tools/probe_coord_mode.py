"""GPU probe: which rounding chain does ATen's CUDA grid_sampler use for the pixel coordinate?

ATen computes ``((g + 1) * size - 1) / 2`` (GridSampler.cuh:30).  nvcc may contract the
multiply-subtract into one FMA; the CPU build rounds each op.  The forward output is
continuous across a floor flip, so the probe reads the *gradient w.r.t. the grid* on a
quadratic ramp image (slope 2k-1 left of pixel k, 2k+1 right of it) at the samples where the
two chains disagree, and counts which chain torch agrees with.  Prints one JSON line.
"""
import json
import sys
import os

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import detrpose_b200 as dp                      # noqa: E402
from detrpose_b200 import _lib                  # noqa: E402


def main():
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(1)
    report = {}
    for W in (80, 40, 20, 100, 75, 50, 25, 15):
        n = 40_000_000
        lx = torch.rand(n, device=dev, generator=g) * 1.0
        loc = torch.stack([lx, torch.full_like(lx, 0.25)], -1).view(1, n // 16, 1, 1, 16, 2).contiguous()
        ia, _ = dp.sample_indices(loc, [(2, W)], coord_mode=_lib.COORD_UNFUSED)
        ib, _ = dp.sample_indices(loc, [(2, W)], coord_mode=_lib.COORD_FMA)
        amb = (ia[..., 1] != ib[..., 1]).view(-1)
        k = int(amb.sum().item())
        entry = {"samples": n, "ambiguous": k, "torch_eq_unfused": 0, "torch_eq_fma": 0, "neither": 0}
        if k:
            lsel = loc.view(-1, 2)[amb]
            xa = ia.view(-1, 2)[amb][:, 1]
            xb = ib.view(-1, 2)[amb][:, 1]
            ramp = (torch.arange(W, device=dev, dtype=torch.float32) ** 2).view(1, 1, 1, W).expand(1, 1, 2, W).contiguous()
            grid = (2 * lsel - 1).view(1, 1, k, 2).requires_grad_(True)
            out = F.grid_sample(ramp, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
            (gg,) = torch.autograd.grad(out.sum(), grid)
            slope = gg.view(k, 2)[:, 0] / (W / 2)            # d out / d x_pixel = f(x0+1) - f(x0) = 2*x0 + 1
            x0_torch = torch.round((slope - 1) / 2).int()
            inside = (xa >= 0) & (xb >= 0) & (xa < W - 1) & (xb < W - 1)
            entry["torch_eq_unfused"] = int(((x0_torch == xa) & inside).sum().item())
            entry["torch_eq_fma"] = int(((x0_torch == xb) & inside).sum().item())
            entry["neither"] = int(((x0_torch != xa) & (x0_torch != xb) & inside).sum().item())
            entry["inside"] = int(inside.sum().item())
        report[str(W)] = entry
    print(json.dumps({"probe": "aten_cuda_grid_sampler_coord_chain", "by_width": report}))


if __name__ == "__main__":
    main()
