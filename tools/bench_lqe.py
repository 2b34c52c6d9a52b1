"""LQE sampler at the DETRPose-S inference shape: batch x 60 queries x 17 keypoints
on the 80x80, 256-channel map.  Times the fused kernels (C ABI, back-to-back launches between two events)
against the reference's op sequence on this GPU (grid_sample + permute + topk + mean + cat,
transformer.py:278-284) and reports the sector traffic rate: the kernel reads 2 rows x 32-byte sectors
per (keypoint, channel) from an NCHW map, which is what bounds it (L2 -> SM sector rate), not HBM.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detrpose_b200.lqe import lqe_statistics                          # noqa: E402


def _time_stream(fn, warmup=10, reps=100):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def reference_ops(feat, poses, k):
    b, p, _ = poses.shape
    v = F.grid_sample(feat, (2 * poses - 1).view(b, p, 1, 2), mode="bilinear", padding_mode="zeros",
                      align_corners=False).permute(0, 2, 3, 1)
    top = v.topk(k, dim=-1)[0]
    return torch.cat([top, top.mean(dim=-1, keepdim=True)], dim=-1).view(b, p, k + 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--queries", type=int, default=60)
    ap.add_argument("--C", type=int, default=256)
    ap.add_argument("--map", type=int, default=80)
    args = ap.parse_args()
    dev, k = "cuda:0", 4
    B, P, C, hw = args.batch, args.queries * 17, args.C, args.map
    out = {"B": B, "points_per_image": P, "C": C, "map": [hw, hw], "topk": k}
    torch.manual_seed(0)
    poses = (torch.rand(B, P, 2, device=dev) * 1.1 - 0.05).requires_grad_(True)
    gs = torch.randn(B, P, k + 1, device=dev)
    for name, dt, cl in (("fp32_nchw", torch.float32, False), ("bf16_nchw", torch.bfloat16, False),
                         ("fp32_channels_last", torch.float32, True)):
        feat = torch.randn(B, C, hw, hw, device=dev, dtype=dt)
        if cl:
            feat = feat.contiguous(memory_format=torch.channels_last)
        feat.requires_grad_(True)
        # grid_sample wants one dtype for map and grid (and runs in fp32 under autocast): the torch arm of the
        # bf16 row samples an fp32 copy of the map made outside the timed region
        feat_ref = feat if dt == torch.float32 else feat.detach().float().requires_grad_(True)
        with torch.no_grad():
            t_f = _time_stream(lambda: lqe_statistics(feat, poses, k))
            t_fr = _time_stream(lambda: reference_ops(feat_ref, poses, k), 5, 20)
        st = lqe_statistics(feat, poses, k)
        ref = reference_ops(feat_ref, poses, k)
        t_b = _time_stream(lambda: torch.autograd.grad(st, [feat, poses], gs, retain_graph=True), 5, 50)
        t_br = _time_stream(lambda: torch.autograd.grad(ref, [feat_ref, poses], gs, retain_graph=True), 5, 20)
        err = float((st.detach() - ref.detach().float()).abs().max() / ref.detach().float().abs().max())
        sector_bytes = B * P * C * 2 * 32
        out[name] = {"fwd_ms": round(t_f, 4), "bwd_ms": round(t_b, 4), "torch_ops_fwd_ms": round(t_fr, 4),
                     "torch_ops_bwd_ms": round(t_br, 4), "speedup_fwd": round(t_fr / t_f, 2),
                     "speedup_fwd_bwd": round((t_fr + t_br) / (t_f + t_b), 2),
                     "fwd_sector_GBps": round(sector_bytes / t_f / 1e6, 1), "rel_err_vs_torch_ops": err}
        del feat, feat_ref, st, ref
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
