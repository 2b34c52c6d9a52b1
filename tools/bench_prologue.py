#!/usr/bin/env python
"""Row f1: what fusing the prologue into the sampler buys (GPU box).

From the two Linear outputs (offsets, logits) and the reference points to the output and back to their
gradients, DETRPose-S shape, through the public autograd API:
  reference  -- the reference's elementwise ops (softmax, divide, add: ms_deform_attn.py:392-393, :412-416) + its
                core (oracle/msda_torch.core), all PyTorch on the device;
  two_step   -- this package's prologue kernel + core kernels (prologue backward = elementwise torch ops);
  fused      -- ms_deform_attn_fused: one launch forward, one launch + the softmax-backward pass backward.
Prints one JSON line per batch size.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import detrpose_b200 as dp                                           # noqa: E402
from detrpose_b200 import functional as MF, synthetic               # noqa: E402
from oracle import msda_torch as otorch                              # noqa: E402  (reference arm only)


def timeit(fn, steps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    dev = "cuda"
    w = synthetic.WORKLOADS["detrpose_s"]
    H, L, P, Dh, Lq = w["H"], len(w["shapes"]), w["P"], w["Dh"], w["Lq"]
    for N, vdt in ((64, torch.bfloat16), (64, torch.float32), (16, torch.bfloat16)):
        g = torch.Generator(device=dev).manual_seed(0)
        S = synthetic.pyramid_size(w["shapes"])
        memory = torch.randn(N, S, H * Dh, device=dev, generator=g).to(vdt).requires_grad_(True)
        offsets = (2.0 * torch.randn(N, Lq, H * L * P * 2, device=dev, generator=g)).requires_grad_(True)
        logits = torch.randn(N, Lq, H * L * P, device=dev, generator=g).requires_grad_(True)
        ref = torch.rand(N, Lq, 1, 2, device=dev, generator=g)
        go = torch.randn(N, Lq, H * Dh, device=dev, generator=g).to(vdt)
        norm = torch.tensor([[wd, h] for h, wd in w["shapes"]], dtype=torch.float32, device=dev).view(1, 1, 1, L, 1, 2)

        def reference():
            weights = torch.softmax(logits.view(N, Lq, H, L * P), -1).view(N, Lq, H, L, P)
            loc = ref[:, :, None, :, None, :] + offsets.view(N, Lq, H, L, P, 2) / norm
            out = otorch.core(otorch.make_value_list(memory.float(), H, w["shapes"]), w["shapes"], loc, weights)
            torch.autograd.grad(out, [memory, offsets, logits], go.float())

        def two_step():
            loc, att = dp.locations_and_weights(offsets, logits, ref, w["shapes"], H, L, P)
            out = dp.ms_deform_attn_core(memory, w["shapes"], loc, att)
            torch.autograd.grad(out, [memory, offsets, logits], go)

        def fused():
            out = MF.ms_deform_attn_fused(memory, w["shapes"], offsets, logits, ref, n_heads=H, n_levels=L, n_points=P)
            torch.autograd.grad(out, [memory, offsets, logits], go)

        rec = {"batch": N, "value_dtype": str(vdt).split(".")[-1], "Lq": Lq}
        rec["two_step_ms"] = round(timeit(two_step), 4)
        rec["fused_ms"] = round(timeit(fused), 4)
        rec["reference_ops_ms"] = round(timeit(reference, steps=5, warm=2), 3)
        rec["fused_vs_two_step"] = round(rec["two_step_ms"] / rec["fused_ms"], 3)
        rec["fused_vs_reference_ops"] = round(rec["reference_ops_ms"] / rec["fused_ms"], 2)
        print(json.dumps(rec), flush=True)
        del memory, offsets, logits
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
